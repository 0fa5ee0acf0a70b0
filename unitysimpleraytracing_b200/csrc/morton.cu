// morton.cu -- K1: per-triangle padded AABB, centroid and 30-bit Morton code.
//
// Replaces the CPU loop of Assets/_Scripts/MeshBufferContainer.cs:123-146 (GetCentroidAndAABB :52-71,
// NormalizeCentroid :73-83, Morton3D :41-50, ExpandBits :32-39). One thread per triangle.
// HBM-bound: 48 B read (three 16-B vertex slots of the 128-B Triangle) + 32 B AABB + 4 B key + 4 B
// value = 88 algorithmic bytes per triangle. The vertex loads carry the .L2::64B fetch-size qualifier (ldg_vertex):
// DRAM then moves the first 64 bytes of a record instead of its whole 128-byte line (ncu at 1M triangles: 134.2 -> 67.1 MB
// read, 27.4 -> 21.7 us; at 16M triangles 0.40 -> 0.29 ms). All fp32 arithmetic is spelled with round-to-nearest
// intrinsics (no FMA contraction, IEEE division) so the keys are bit-identical to the oracle.

#include "usrt_internal.cuh"

#include <algorithm>
#include <cmath>

namespace usrt {

__device__ __forceinline__ uint32_t expand_bits(uint32_t v) {   // MeshBufferContainer.cs:32-39
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}

__device__ __forceinline__ float sel_min(float a, float b) { return a < b ? a : b; }
__device__ __forceinline__ float sel_max(float a, float b) { return a > b ? a : b; }

__device__ __forceinline__ uint32_t quantise(float x) {          // MeshBufferContainer.cs:43,46
    x = sel_min(sel_max(__fmul_rn(x, 1024.0f), 0.0f), 1023.0f);
    return (uint32_t)x;                                           // truncation, like C# (uint)float
}

// 64-bit variant (the reference's sorter is generic over uint / ulong keys but ships no 64-bit Morton
// function, so this is defined by analogy and restated in the oracle): 21 bits per axis, the classic 64-bit spread.
__device__ __forceinline__ uint64_t expand_bits64(uint64_t v) {
    v &= 0x1FFFFFull;
    v = (v | (v << 32)) & 0x001F00000000FFFFull;
    v = (v | (v << 16)) & 0x001F0000FF0000FFull;
    v = (v | (v << 8)) & 0x100F00F00F00F00Full;
    v = (v | (v << 4)) & 0x10C30C30C30C30C3ull;
    v = (v | (v << 2)) & 0x1249249249249249ull;
    return v;
}
__device__ __forceinline__ uint32_t quantise21(float x) {
    x = sel_min(sel_max(__fmul_rn(x, 2097152.0f), 0.0f), 2097151.0f);
    return (uint32_t)x;
}

template <bool kWide>
__global__ void __launch_bounds__(256) k_morton(VertexSource src, uint32_t n, WorldBox whole,
                                                uint32_t* __restrict__ keys, uint64_t* __restrict__ keys64,
                                                uint32_t* __restrict__ values, float4* __restrict__ aabbs) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4* t = src.base + (size_t)i * src.stride;   // 128-B Triangle = 8 x float4 (or a 48-B record); a, b, c are slots 0..2
    const float4 a = ldg_vertex(t + 0), b = ldg_vertex(t + 1), c = ldg_vertex(t + 2);

    // GetCentroidAndAABB (:52-71)
    const float mnx = __fsub_rn(sel_min(sel_min(a.x, b.x), c.x), 0.001f);
    const float mny = __fsub_rn(sel_min(sel_min(a.y, b.y), c.y), 0.001f);
    const float mnz = __fsub_rn(sel_min(sel_min(a.z, b.z), c.z), 0.001f);
    const float mxx = __fadd_rn(sel_max(sel_max(a.x, b.x), c.x), 0.001f);
    const float mxy = __fadd_rn(sel_max(sel_max(a.y, b.y), c.y), 0.001f);
    const float mxz = __fadd_rn(sel_max(sel_max(a.z, b.z), c.z), 0.001f);
    float cx = __fmul_rn(__fadd_rn(mnx, mxx), 0.5f);
    float cy = __fmul_rn(__fadd_rn(mny, mxy), 0.5f);
    float cz = __fmul_rn(__fadd_rn(mnz, mxz), 0.5f);

    // NormalizeCentroid (:73-83): subtract, then a true division by (max - min), per axis (the reference's
    // Whole box is the cube +-125, MeshBufferContainer.cs:9-15; a fitted scene box is per axis)
    cx = __fdiv_rn(__fsub_rn(cx, whole.min[0]), __fsub_rn(whole.max[0], whole.min[0]));
    cy = __fdiv_rn(__fsub_rn(cy, whole.min[1]), __fsub_rn(whole.max[1], whole.min[1]));
    cz = __fdiv_rn(__fsub_rn(cz, whole.min[2]), __fsub_rn(whole.max[2], whole.min[2]));

    // Morton3D (:41-50)
    if (kWide) {
        keys64[i] = expand_bits64(quantise21(cx)) * 4 + expand_bits64(quantise21(cy)) * 2 + expand_bits64(quantise21(cz));
    } else {
        keys[i] = expand_bits(quantise(cx)) * 4 + expand_bits(quantise(cy)) * 2 + expand_bits(quantise(cz));
    }
    values[i] = i;                                                 // :132
    aabbs[(size_t)i * 2 + 0] = make_float4(mnx, mny, mnz, 0.0f);   // pads are C# default(0)
    aabbs[(size_t)i * 2 + 1] = make_float4(mxx, mxy, mxz, 0.0f);
}

// The "find the scene AABB at runtime" TODO of MeshBufferContainer.cs:7: per-axis min / max over all vertices.
// out[0..2] = min, out[3..5] = max, pre-set to +inf / -inf. Floats are reduced with the signed-magnitude trick
// (non-negative values order like ints, negative ones like reversed unsigned ints).
__device__ __forceinline__ void atomic_min_float(float* addr, float v) {
    if (v >= 0.0f) atomicMin(reinterpret_cast<int*>(addr), __float_as_int(v));
    else atomicMax(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}
__device__ __forceinline__ void atomic_max_float(float* addr, float v) {
    if (v >= 0.0f) atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
    else atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

__global__ void __launch_bounds__(256) k_scene_box(VertexSource src, uint32_t n, float* __restrict__ out) {
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float4* t = src.base + (size_t)i * src.stride;
        const float4 a = ldg_vertex(t + 0), b = ldg_vertex(t + 1), c = ldg_vertex(t + 2);
        mn[0] = sel_min(mn[0], sel_min(sel_min(a.x, b.x), c.x)); mx[0] = sel_max(mx[0], sel_max(sel_max(a.x, b.x), c.x));
        mn[1] = sel_min(mn[1], sel_min(sel_min(a.y, b.y), c.y)); mx[1] = sel_max(mx[1], sel_max(sel_max(a.y, b.y), c.y));
        mn[2] = sel_min(mn[2], sel_min(sel_min(a.z, b.z), c.z)); mx[2] = sel_max(mx[2], sel_max(sel_max(a.z, b.z), c.z));
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            mn[k] = sel_min(mn[k], __shfl_xor_sync(0xFFFFFFFFu, mn[k], off));
            mx[k] = sel_max(mx[k], __shfl_xor_sync(0xFFFFFFFFu, mx[k], off));
        }
    }
    if ((threadIdx.x & 31u) == 0) {
#pragma unroll
        for (int k = 0; k < 3; ++k) { atomic_min_float(out + k, mn[k]); atomic_max_float(out + 3 + k, mx[k]); }
    }
}

cudaError_t launch_scene_box(VertexSource src, uint32_t n, float* out6, cudaStream_t stream) {
    static const float init[6] = {INFINITY, INFINITY, INFINITY, -INFINITY, -INFINITY, -INFINITY};
    cudaError_t e = cudaMemcpyAsync(out6, init, sizeof(init), cudaMemcpyHostToDevice, stream);
    if (e != cudaSuccess || n == 0) return e;
    const uint32_t grid = std::min<uint32_t>((n + 255) / 256, 148 * 8);
    k_scene_box<<<grid, 256, 0, stream>>>(src, n, out6);
    return cudaGetLastError();
}

cudaError_t launch_morton(VertexSource src, uint32_t n, const WorldBox& whole, uint32_t* keys, uint32_t* values, usrt_aabb* aabbs,
                          cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    k_morton<false><<<(n + 255) / 256, 256, 0, stream>>>(src, n, whole, keys, nullptr, values, reinterpret_cast<float4*>(aabbs));
    return cudaGetLastError();
}

cudaError_t launch_morton64(VertexSource src, uint32_t n, const WorldBox& whole, uint64_t* keys, uint32_t* values,
                            usrt_aabb* aabbs, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    k_morton<true><<<(n + 255) / 256, 256, 0, stream>>>(src, n, whole, nullptr, keys, values, reinterpret_cast<float4*>(aabbs));
    return cudaGetLastError();
}

}  // namespace usrt
