// microbenchmark (developer tool, not part of the library): cost of one warp "ranking round" of the radix pass --
// every lane has an 8-bit digit and needs (a) its rank among the lanes of the warp holding the same digit and
// (b) the running count of that digit over the warp's earlier rounds -- two shared-memory layouts:
//   A  one 64-bit {mask, count} word per (warp, digit): atomicOr on the mask half, 64-bit read by every lane,
//      64-bit write-back {0, count + popc} by the leader                         (what k_onesweep does today)
//   B  separate 32-bit arrays: atomicOr on mask[warp][digit], 32-bit read by every lane, the LEADER alone does
//      atomicAdd(count[warp][digit], popc) and clears the mask; peers get the old count by shuffle
// Both produce the same ranks (checked). Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a rank_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int kWarps = 16;   // 512-thread CTA as in the big tile

template <int MODE>
__global__ void __launch_bounds__(kWarps * 32) k(uint32_t* out, int rounds, uint32_t seed, int skew_bits) {
    __shared__ unsigned long long s_word[MODE == 0 ? kWarps : 1][256];
    __shared__ uint32_t s_mask[MODE == 1 ? kWarps : 1][256];
    __shared__ uint32_t s_cnt[MODE == 1 ? kWarps : 1][256];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < kWarps * 256; i += blockDim.x) {
        if (MODE == 0) (&s_word[0][0])[i] = 0ull;
        else { (&s_mask[0][0])[i] = 0u; (&s_cnt[0][0])[i] = 0u; }
    }
    __syncthreads();
    uint32_t x = seed * 2654435761u + threadIdx.x * 40503u + blockIdx.x * 977u;
    uint32_t acc = 0;
    for (int r = 0; r < rounds; ++r) {
        x = x * 1664525u + 1013904223u;
        const uint32_t d = ((x >> 13) & 255u) >> skew_bits << skew_bits;      // skew_bits > 0: fewer distinct digits
        uint32_t rank;
        if (MODE == 0) {
            uint32_t* half = reinterpret_cast<uint32_t*>(&s_word[warp][d]);   // [0] = mask, [1] = count
            atomicOr(half, 1u << lane);
            __syncwarp();
            const unsigned long long w = s_word[warp][d];
            const uint32_t m = (uint32_t)w, c = (uint32_t)(w >> 32);
            __syncwarp();
            if ((m & ((1u << lane) - 1u)) == 0) s_word[warp][d] = (unsigned long long)(c + __popc(m)) << 32;
            __syncwarp();
            rank = c + __popc(m & ((1u << lane) - 1u));
        } else {
            atomicOr(&s_mask[warp][d], 1u << lane);
            __syncwarp();
            const uint32_t m = s_mask[warp][d];
            __syncwarp();
            const int leader = __ffs(m) - 1;
            uint32_t c = 0;
            if ((int)lane == leader) { c = atomicAdd(&s_cnt[warp][d], (uint32_t)__popc(m)); s_mask[warp][d] = 0; }
            c = __shfl_sync(0xFFFFFFFFu, c, leader);
            __syncwarp();
            rank = c + __popc(m & ((1u << lane) - 1u));
        }
        acc = acc * 31u + rank;
    }
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
    uint32_t *o0, *o1; cudaMalloc(&o0, n * 4); cudaMalloc(&o1, n * 4);
    uint32_t* h0 = new uint32_t[n]; uint32_t* h1 = new uint32_t[n];
    for (int skew = 0; skew <= 6; skew += 3) {
        const float a = run<0>(o0, rounds, skew), b = run<1>(o1, rounds, skew);
        cudaMemcpy(h0, o0, n * 4, cudaMemcpyDeviceToHost); cudaMemcpy(h1, o1, n * 4, cudaMemcpyDeviceToHost);
        int bad = 0; for (int i = 0; i < n; ++i) bad += h0[i] != h1[i];
        const double warp_rounds_per_sm = (double)2 * kWarps * rounds;
        printf("distinct digits 2^%d: A {mask,count} 64-bit %.3f ms (%.1f SM-cycles/round @1.9GHz) | B split + leader atomicAdd %.3f ms (%.1f) | mismatches %d | %s\n",
               8 - skew, a, a * 1e-3 * 1.9e9 / warp_rounds_per_sm, b, b * 1e-3 * 1.9e9 / warp_rounds_per_sm, bad,
               cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
