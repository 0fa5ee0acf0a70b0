// microbenchmark (developer tool): ranking round candidates, part 3.
//   D  16-lane groups, word = mask16<<16 | count16: atomicAdd(bit+1), warp sync, read back, leader stores count
//   E  same word, but NO read-back: old = atomicAdd(word, bit+1) already holds {lanes whose add landed earlier
//      this round, running count}. If the hardware serialises same-address lanes in ascending lane order, old.count IS
//      the stable rank; each lane verifies that locally (no higher lane's bit in old.mask) and a violation anywhere
//      sends the warp to the slow path. Each lane then removes its own bit (atomicSub, no return value).
//   E64 full 32-lane warps with a 64-bit word {mask32 | count32} and 64-bit shared atomics.
// Reports SM-cycles per warp round and how many lanes saw an order violation.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/micro/rank3_bench.cu -o tools/micro/rank3_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
constexpr int kWarps = 16;
__device__ __forceinline__ uint32_t lanemask_lt() { uint32_t m; asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m)); return m; }
__device__ unsigned long long g_viol[3];

template <int MODE>
__global__ void __launch_bounds__(kWarps * 32, 2) k(uint32_t* out, int rounds, uint32_t seed, int skew_bits) {
    __shared__ uint32_t s_cnt[MODE == 2 ? 1 : kWarps * 512];
    __shared__ unsigned long long s_w64[MODE == 2 ? kWarps * 256 : 1];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, lt = lanemask_lt();
    for (int i = threadIdx.x; i < kWarps * 512; i += blockDim.x) { if (MODE != 2) s_cnt[i] = 0; else if (i < kWarps * 256) s_w64[i] = 0; }
    __syncthreads();
    uint32_t x = seed * 2654435761u + threadIdx.x * 40503u + blockIdx.x * 977u;
    uint32_t acc = 0, viol = 0;
    for (int r = 0; r < rounds; ++r) {
        x = x * 1664525u + 1013904223u;
        const uint32_t d = ((x >> 13) & 255u) >> skew_bits << skew_bits;
        uint32_t rank;
        if (MODE == 0) {
            uint32_t* cnt = s_cnt + (warp * 2 + (lane >> 4)) * 256;
            const uint32_t bit = 0x10000u << (lane & 15u);
            atomicAdd(cnt + d, bit + 1u);
            __syncwarp();
            const uint32_t now = cnt[d];
            __syncwarp();
            const uint32_t m = now >> 16, c = now & 0xFFFFu;
            const uint32_t before = __popc(m & (lt >> (lane & 16u)) & 0xFFFFu);
            if (before == 0) cnt[d] = c;
            __syncwarp();
            rank = c - __popc(m) + before;
        } else if (MODE == 1) {
            uint32_t* cnt = s_cnt + (warp * 2 + (lane >> 4)) * 256;
            const uint32_t bit = 0x10000u << (lane & 15u);
            const uint32_t old = atomicAdd(cnt + d, bit + 1u);
            __syncwarp();
            atomicSub(cnt + d, bit);
            __syncwarp();
            viol += ((old >> 16) & ~((bit >> 16) - 1u)) != 0u;       // a higher lane of my group landed before me
            rank = old & 0xFFFFu;
        } else {
            unsigned long long* w = s_w64 + warp * 256;
            const unsigned long long bit = 1ull << (32 + lane);
            const unsigned long long old = atomicAdd(w + d, bit + 1ull);
            __syncwarp();
            atomicAdd(w + d, 0ull - bit);
            __syncwarp();
            viol += ((uint32_t)(old >> 32) & ~lt) != 0u;
            rank = (uint32_t)old;
        }
        acc = acc * 31u + rank;
    }
    if (viol) atomicAdd(&g_viol[MODE], (unsigned long long)viol);
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <int MODE> float run(uint32_t* out, int rounds, int skew) {
    const int grid = 148 * 2;
    k<MODE><<<grid, kWarps * 32>>>(out, 8, 1, skew);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    k<MODE><<<grid, kWarps * 32>>>(out, rounds, 2, skew);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    return ms;
}
int main() {
    const int rounds = 4096, n = 148 * 2 * kWarps * 32;
    uint32_t *o0, *o1, *o2; cudaMalloc(&o0, n * 4); cudaMalloc(&o1, n * 4); cudaMalloc(&o2, n * 4);
    uint32_t* h0 = new uint32_t[n]; uint32_t* h1 = new uint32_t[n];
    for (int skew = 0; skew <= 8; skew += 2) {
        unsigned long long z[3] = {0, 0, 0}; cudaMemcpyToSymbol(g_viol, z, sizeof(z));
        const float d = run<0>(o0, rounds, skew), e = run<1>(o1, rounds, skew), e64 = run<2>(o2, rounds, skew);
        cudaMemcpy(h0, o0, n * 4, cudaMemcpyDeviceToHost); cudaMemcpy(h1, o1, n * 4, cudaMemcpyDeviceToHost);
        int bad = 0; for (int i = 0; i < n; ++i) bad += h0[i] != h1[i];
        cudaMemcpyFromSymbol(z, g_viol, sizeof(z));
        const double wr = (double)2 * kWarps * rounds;
        auto cyc = [&](float ms) { return ms * 1e-3 * 1.965e9 / wr; };
        printf("distinct 2^%d: D %.3f ms (%.2f cyc/round) | E add+sub %.3f ms (%.2f) rank mismatches vs D %d, order violations %llu | E64 %.3f ms (%.2f) violations %llu | %s\n",
               8 - skew, d, cyc(d), e, cyc(e), bad, z[1], e64, cyc(e64), z[2], cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
