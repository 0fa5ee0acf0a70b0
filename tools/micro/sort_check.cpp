// developer tool: correctness of usrt_sort_pairs_host against std::stable_sort on a few sizes (both tile shapes,
// ragged tails, heavy duplicates) + device time of a 2^26-pair sort, without Python (fast to run on a GPU box).
//   g++ -O2 -Iinclude tools/micro/sort_check.cpp -o tools/micro/sort_check -Lunitysimpleraytracing_b200 -lusrt_b200 -Wl,-rpath,'$ORIGIN/../../unitysimpleraytracing_b200'
#include <algorithm>
#include <cstdio>
#include <cstdint>
#include <numeric>
#include <random>
#include <vector>
#include "usrt.h"

int main() {
    usrt_context* ctx = nullptr;
    if (usrt_create(0, 2, &ctx) != 0) { printf("create failed\n"); return 2; }
    std::mt19937_64 rng(12345);
    int bad = 0;
    for (uint64_t n : {1ull, 2ull, 33ull, 1000ull, (1ull << 18) - 1, 1ull << 18, (1ull << 18) + 1, 1ull << 20, (1ull << 22) + 77}) {
        for (int kind = 0; kind < 3; ++kind) {
            std::vector<uint32_t> k(n), v(n);
            for (uint64_t i = 0; i < n; ++i) {
                const uint64_t r = rng();
                k[i] = kind == 0 ? (uint32_t)r : kind == 1 ? (uint32_t)(r % 37) << 22 : (uint32_t)(r >> 34);   // random / few distinct / 30-bit
                v[i] = (uint32_t)i;
            }
            std::vector<uint32_t> idx(n); std::iota(idx.begin(), idx.end(), 0u);
            std::stable_sort(idx.begin(), idx.end(), [&](uint32_t a, uint32_t b) { return k[a] < k[b]; });
            std::vector<uint32_t> kk = k, vv = v;
            if (usrt_sort_pairs_host(ctx, kk.data(), vv.data(), n) != 0) { printf("sort failed: %s\n", usrt_last_error(ctx)); return 2; }
            uint64_t wrong = 0;
            for (uint64_t i = 0; i < n; ++i) wrong += (kk[i] != k[idx[i]]) || (vv[i] != idx[i]);
            if (wrong) { printf("MISMATCH n=%llu kind=%d: %llu wrong\n", (unsigned long long)n, kind, (unsigned long long)wrong); ++bad; }
        }
    }
    printf("correctness: %s\n", bad ? "FAILED" : "all sizes ok");
    const uint64_t n = 1ull << 26;
    std::vector<uint32_t> k(n), v(n);
    for (uint64_t i = 0; i < n; ++i) { k[i] = (uint32_t)rng(); v[i] = (uint32_t)i; }
    usrt_enable_stage_timing(ctx, 1);
    float best = 1e9f, ms[6];
    for (int it = 0; it < 3; ++it) {
        std::vector<uint32_t> kk = k, vv = v;
        usrt_sort_pairs_host(ctx, kk.data(), vv.data(), n);
        usrt_last_sort_ms(ctx, ms);
        best = std::min(best, ms[5]);
        if (it == 2) {
            uint64_t unsorted = 0; for (uint64_t i = 1; i < n; ++i) unsorted += kk[i - 1] > kk[i];
            printf("2^26 pairs: total %.3f ms (hist %.3f, passes %.3f %.3f %.3f %.3f), unsorted pairs %llu\n", best, ms[0], ms[1], ms[2], ms[3], ms[4],
                   (unsigned long long)unsorted);
        }
    }
    usrt_destroy(ctx);
    return bad ? 1 : 0;
}
