// shade.cu -- SURVEY.md 8(f)-1: the shading epilogue of kernel Raytracing,
// Assets/_Shaders/Raytracing/Raytracing.compute:178-184, as a separate streaming kernel over the hit
// records: barycentric UV / normal interpolation, the reference's scalar `lightDir` quirk (:181 declares
// it `float`, so only .x = 1/sqrt(3) survives and dot() splats it), max(0.4, .), a bilinear-clamp
// texture fetch, RGBA16F output with alpha = hit (RaytracingMeshDrawer.cs:56, ImageComposer.shader:49).
//
// HBM-bound: 16 B hit + 80 B of the hit triangle (uv + normals) + 4 texels read, 8 B written per pixel.
// The sampler arithmetic the hardware would do in fixed point is DEFINED by the oracle (texel centres at
// (i+0.5)/size, clamp addressing, fp32 weights, lerp = a + (b-a)*t, x then y) and spelled here with the
// same round-to-nearest operations.

#include <cuda_fp16.h>

#include "usrt_internal.cuh"

namespace usrt {

namespace {

__device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float lerp(float a, float b, float t) { return add(a, mul(sub(b, a), t)); }

__device__ __forceinline__ int clamp_texel(float f, int hi) {
    if (!(f >= 0.0f)) return 0;
    if (f > (float)hi) return hi;
    return (int)f;
}

__global__ void __launch_bounds__(256) k_shade(const float4* __restrict__ hits, uint64_t count,
                                               const float4* __restrict__ tris, const float4* __restrict__ tex, int tw, int th,
                                               uint2* __restrict__ out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const float4 h = __ldg(hits + i);                                  // {distance, triangleIndex, u, v}
    const float4* t = tris + (size_t)__float_as_uint(h.y) * 8;          // :178 (triangle 0 for a miss)
    const float4 uv_ab = __ldg(t + 3), uv_c = __ldg(t + 4);            // a_uv, b_uv | c_uv, pad
    const float4 na = __ldg(t + 5), nb = __ldg(t + 6), nc = __ldg(t + 7);
    const float bu = h.z, bv = h.w;
    const float bw = sub(sub(1.0f, bu), bv);
    // :179-180  w*a + u*b + v*c, left to right
    const float u = add(add(mul(bw, uv_ab.x), mul(bu, uv_ab.z)), mul(bv, uv_c.x));
    const float v = add(add(mul(bw, uv_ab.y), mul(bu, uv_ab.w)), mul(bv, uv_c.y));
    const float nx = add(add(mul(bw, na.x), mul(bu, nb.x)), mul(bv, nc.x));
    const float ny = add(add(mul(bw, na.y), mul(bu, nb.y)), mul(bv, nc.y));
    const float nz = add(add(mul(bw, na.z), mul(bu, nb.z)), mul(bv, nc.z));
    // :181 scalar lightDir = normalize(float3(1,1,1)).x
    const float light = __fdiv_rn(1.0f, __fsqrt_rn(add(add(mul(1.0f, 1.0f), mul(1.0f, 1.0f)), mul(1.0f, 1.0f))));
    const float shade = fmaxf(0.4f, add(add(mul(light, nx), mul(light, ny)), mul(light, nz)));   // :183

    // SampleLevel(linearClampSampler, uv, 0)
    const float x = sub(mul(u, (float)tw), 0.5f), y = sub(mul(v, (float)th), 0.5f);
    const float x0f = floorf(x), y0f = floorf(y);
    const float fx = sub(x, x0f), fy = sub(y, y0f);
    const int x0 = clamp_texel(x0f, tw - 1), x1 = clamp_texel(add(x0f, 1.0f), tw - 1);
    const int y0 = clamp_texel(y0f, th - 1), y1 = clamp_texel(add(y0f, 1.0f), th - 1);
    const float4 c00 = __ldg(tex + (size_t)y0 * tw + x0), c10 = __ldg(tex + (size_t)y0 * tw + x1);
    const float4 c01 = __ldg(tex + (size_t)y1 * tw + x0), c11 = __ldg(tex + (size_t)y1 * tw + x1);
    const float r = lerp(lerp(c00.x, c10.x, fx), lerp(c01.x, c11.x, fx), fy);
    const float g = lerp(lerp(c00.y, c10.y, fx), lerp(c01.y, c11.y, fx), fy);
    const float b = lerp(lerp(c00.z, c10.z, fx), lerp(c01.z, c11.z, fx), fy);

    const float alpha = (__float_as_uint(h.x) != 0x4EFF0000u) ? 1.0f : 0.0f;   // :184 distance != MAX_FLOAT
    const __half2 rg = __halves2half2(__float2half_rn(mul(r, shade)), __float2half_rn(mul(g, shade)));
    const __half2 ba = __halves2half2(__float2half_rn(mul(b, shade)), __float2half_rn(alpha));
    out[i] = make_uint2(*reinterpret_cast<const uint32_t*>(&rg), *reinterpret_cast<const uint32_t*>(&ba));
}

}  // namespace

cudaError_t launch_shade(const usrt_raycast_result* hits, uint64_t count, const usrt_triangle* tris, const float4* tex,
                         int tw, int th, void* out_rgba16f, cudaStream_t stream) {
    if (count == 0) return cudaSuccess;
    k_shade<<<(uint32_t)((count + 255) / 256), 256, 0, stream>>>(reinterpret_cast<const float4*>(hits), count,
                                                                reinterpret_cast<const float4*>(tris), tex, tw, th,
                                                                static_cast<uint2*>(out_rgba16f));
    return cudaGetLastError();
}

}  // namespace usrt
