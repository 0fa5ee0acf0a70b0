// microbenchmark (developer tool, not part of the library): cost of one warp "ranking round" of the radix pass,
// round-2 candidates. Every lane holds an 8-bit digit and needs its stable rank among the keys of its warp with the
// same digit (earlier rounds first, then lower lanes of this round).
//   A  (round 1 kernel) one 64-bit {mask,count} word per (warp,digit), bank-swizzled halves: atomicOr + 64-bit read
//      + leader 64-bit write-back
//   C  32-bit running count per (warp,digit): old = atomicAdd(cnt[d],1); now = cnt[d] after a warp sync. A lane that is
//      alone on its digit this round has now == old+1 and old IS its rank. Digits held by several lanes (1.8 per round
//      for random digits) are fixed up in a warp-uniform loop: shfl the digit of one suspect lane, ballot its peers,
//      rank = now - popc(peers) + popc(peers & lanemask_lt).
//   D  16-lane groups, word = {mask16 << 16 | count16}: old/now as in C but the mask comes back with the count, the
//      leader stores {0, count} (third op, 32-bit)
// All three must produce identical ranks (checked against A). Build:
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/micro/rank2_bench.cu -o tools/micro/rank2_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int kWarps = 16;

__device__ __forceinline__ uint32_t lanemask_lt() { uint32_t m; asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m)); return m; }

template <int MODE>
__global__ void __launch_bounds__(kWarps * 32, 2) k(uint32_t* out, int rounds, uint32_t seed, int skew_bits) {
    __shared__ uint2 s_word[MODE == 0 ? kWarps * 256 : 1];
    __shared__ uint32_t s_cnt[MODE == 1 ? kWarps * 256 : (MODE == 2 ? kWarps * 512 : 1)];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, lt = lanemask_lt();
    for (int i = threadIdx.x; i < kWarps * 256; i += blockDim.x) {
        if (MODE == 0) s_word[i] = make_uint2(0, 0);
        if (MODE == 1) s_cnt[i] = 0;
        if (MODE == 2) { s_cnt[i] = 0; s_cnt[i + kWarps * 256] = 0; }
    }
    __syncthreads();
    uint32_t x = seed * 2654435761u + threadIdx.x * 40503u + blockIdx.x * 977u;
    uint32_t acc = 0;
    for (int r = 0; r < rounds; ++r) {
        x = x * 1664525u + 1013904223u;
        const uint32_t d = ((x >> 13) & 255u) >> skew_bits << skew_bits;
        uint32_t rank;
        if (MODE == 0) {
            uint2* my = s_word + warp * 256;
            uint32_t* words = reinterpret_cast<uint32_t*>(my);
            const uint32_t sel = (d >> 4) & 1u;
            atomicOr(words + 2u * d + sel, 1u << lane);
            __syncwarp();
            const uint2 w = my[d];
            __syncwarp();
            const uint32_t mask = sel ? w.y : w.x, prior = sel ? w.x : w.y;
            const uint32_t before = __popc(mask & lt);
            if (before == 0) {
                const uint32_t now = prior + __popc(mask);
                my[d] = sel ? make_uint2(now, 0u) : make_uint2(0u, now);
            }
            __syncwarp();
            rank = prior + before;
        } else if (MODE == 1) {
            uint32_t* cnt = s_cnt + warp * 256;
            uint32_t old = atomicAdd(cnt + d, 1u);
            __syncwarp();
            const uint32_t now = cnt[d];
            __syncwarp();
            uint32_t coll = __ballot_sync(0xFFFFFFFFu, now - old != 1u);
            while (coll) {
                const int src = __ffs(coll) - 1;
                const uint32_t dc = __shfl_sync(0xFFFFFFFFu, d, src);
                const uint32_t m = __ballot_sync(0xFFFFFFFFu, d == dc);
                if (d == dc) old = now - __popc(m) + __popc(m & lt);
                coll &= ~m;
            }
            rank = old;
        } else {
            // 16-lane groups: group g = lane >> 4 has its own table; word = mask16 << 16 | count16
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
    uint32_t *o0, *o1, *o2; cudaMalloc(&o0, n * 4); cudaMalloc(&o1, n * 4); cudaMalloc(&o2, n * 4);
    uint32_t* h0 = new uint32_t[n]; uint32_t* h1 = new uint32_t[n]; uint32_t* h2 = new uint32_t[n];
    for (int skew = 0; skew <= 8; skew += 2) {
        const float a = run<0>(o0, rounds, skew), c = run<1>(o1, rounds, skew), d = run<2>(o2, rounds, skew);
        cudaMemcpy(h0, o0, n * 4, cudaMemcpyDeviceToHost); cudaMemcpy(h1, o1, n * 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(h2, o2, n * 4, cudaMemcpyDeviceToHost);
        int bad1 = 0; for (int i = 0; i < n; ++i) bad1 += h0[i] != h1[i];
        const double wr = (double)2 * kWarps * rounds;      // warp rounds per SM
        auto cyc = [&](float ms) { return ms * 1e-3 * 1.965e9 / wr; };
        printf("distinct digits 2^%d: A %.3f ms (%.2f SM-cyc/round) | C atomicAdd+fixup %.3f ms (%.2f) mismatches vs A %d | D 16-lane %.3f ms (%.2f) [ranks differ by design: half-warp groups] | %s\n",
               8 - skew, a, cyc(a), c, cyc(c), bad1, d, cyc(d), cudaGetErrorString(cudaGetLastError()));
    }
    (void)h2;
    return 0;
}
