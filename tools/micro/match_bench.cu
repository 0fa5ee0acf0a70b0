// microbenchmark: cost of computing the same-digit peer mask of a warp, three ways
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t peers_ballot(uint32_t d) {
    uint32_t peers = 0xFFFFFFFFu;
#pragma unroll
    for (int b = 0; b < 8; ++b) {
        const bool bit = (d >> b) & 1u;
        const uint32_t ballot = __ballot_sync(0xFFFFFFFFu, bit);
        peers &= bit ? ballot : ~ballot;
    }
    return peers;
}
template <int MODE>
__global__ void __launch_bounds__(256) k(uint32_t* out, int iters, uint32_t seed) {
    __shared__ uint32_t s_mask[8][256];
    uint32_t x = seed * 2654435761u + threadIdx.x * 40503u + blockIdx.x * 977u;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (MODE == 2) { for (int i = threadIdx.x; i < 8 * 256; i += 256) (&s_mask[0][0])[i] = 0; __syncthreads(); }
    uint32_t acc = 0;
    for (int i = 0; i < iters; ++i) {
        x = x * 1664525u + 1013904223u;
        const uint32_t d = (x >> 13) & 255u;
        uint32_t m;
        if (MODE == 0) m = peers_ballot(d);
        else if (MODE == 1) m = __match_any_sync(0xFFFFFFFFu, d);
        else {
            atomicOr(&s_mask[warp][d], 1u << lane);
            __syncwarp();
            m = s_mask[warp][d];
            __syncwarp();
            s_mask[warp][d] = 0;
            __syncwarp();
        }
        acc += __popc(m) + (__ffs(m) << 8);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <int MODE> void run(const char* name, uint32_t* out) {
    const int iters = 4096, grid = 148 * 8;
    k<MODE><<<grid, 256>>>(out, 16, 1);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    k<MODE><<<grid, 256>>>(out, iters, 2);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    const double warp_ops = (double)grid * 8 * iters;
    // cycles per warp-op per SM at 1.9 GHz, 148 SMs
    printf("%-12s %.3f ms  %.2f ns per warp-op per SM  (~%.1f SM-cycles @1.9GHz)  err=%s\n", name, ms,
           ms * 1e6 / (warp_ops / 148), ms * 1e6 / (warp_ops / 148) * 1.9, cudaGetErrorString(cudaGetLastError()));
}
int main() {
    uint32_t* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
    run<0>("ballot x8", out); run<1>("match.any", out); run<2>("smem atomOr", out);
    run<0>("ballot x8", out); run<1>("match.any", out); run<2>("smem atomOr", out);
    return 0;
}
