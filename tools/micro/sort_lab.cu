// developer tool: compiles csrc/radix_sort.cu straight into a small driver so that kernel variants (-D switches,
// tile shapes) can be built side by side and timed in one GPU call, each checked against std::stable_sort.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false --expt-relaxed-constexpr \
//        [-DUSRT_BIG_BLOCK=512 -DUSRT_BIG_IPT=16 -DUSRT_BIG_CTAS=2 ...] tools/micro/sort_lab.cu -o tools/micro/lab_x
#include "../../unitysimpleraytracing_b200/csrc/radix_sort.cu"
#include <algorithm>
#include <cstdio>
#include <numeric>
#include <random>
#include <vector>
#ifndef LAB_NAME
#define LAB_NAME "default"
#endif
int main(int argc, char** argv) {
    using namespace usrt;
    const int lg = argc > 1 ? atoi(argv[1]) : 26;
    cudaStream_t st; cudaStreamCreate(&st);
    SortScratch sc;
    std::mt19937_64 rng(99);
    // correctness: ragged size, three key distributions
    int bad = 0;
    for (uint64_t n : {(1ull << 18) + 5, (1ull << 22) + 77}) {
        for (int kind = 0; kind < 3; ++kind) {
            std::vector<uint32_t> k(n), v(n), idx(n);
            for (uint64_t i = 0; i < n; ++i) { const uint64_t r = rng(); k[i] = kind == 0 ? (uint32_t)r : kind == 1 ? (uint32_t)(r % 37) << 22 : (uint32_t)(r >> 34); v[i] = (uint32_t)i; }
            std::iota(idx.begin(), idx.end(), 0u);
            std::stable_sort(idx.begin(), idx.end(), [&](uint32_t a, uint32_t b) { return k[a] < k[b]; });
            uint32_t *dk, *dv, *ak, *av;
            cudaMalloc(&dk, n * 4); cudaMalloc(&dv, n * 4); cudaMalloc(&ak, n * 4); cudaMalloc(&av, n * 4);
            cudaMemcpy(dk, k.data(), n * 4, cudaMemcpyHostToDevice); cudaMemcpy(dv, v.data(), n * 4, cudaMemcpyHostToDevice);
            uint64_t launches = 0;
            if (sort_pairs(dk, dv, ak, av, n, sc, st, &launches) != cudaSuccess) { printf("launch failed\n"); return 2; }
            cudaStreamSynchronize(st);
            std::vector<uint32_t> kk(n), vv(n);
            cudaMemcpy(kk.data(), dk, n * 4, cudaMemcpyDeviceToHost); cudaMemcpy(vv.data(), dv, n * 4, cudaMemcpyDeviceToHost);
            uint64_t wrong = 0;
            for (uint64_t i = 0; i < n; ++i) wrong += (kk[i] != k[idx[i]]) || (vv[i] != idx[i]);
            if (wrong) { printf("MISMATCH n=%llu kind=%d wrong=%llu\n", (unsigned long long)n, kind, (unsigned long long)wrong); ++bad; }
            cudaFree(dk); cudaFree(dv); cudaFree(ak); cudaFree(av);
        }
    }
    // timing
    const uint64_t n = 1ull << lg;
    std::vector<uint32_t> k(n);
    for (uint64_t i = 0; i < n; ++i) k[i] = (uint32_t)rng();
    uint32_t *k0, *dk, *dv, *ak, *av;
    cudaMalloc(&k0, n * 4); cudaMalloc(&dk, n * 4); cudaMalloc(&dv, n * 4); cudaMalloc(&ak, n * 4); cudaMalloc(&av, n * 4);
    cudaMemcpy(k0, k.data(), n * 4, cudaMemcpyHostToDevice);
    cudaEvent_t ev[6]; for (auto& e : ev) cudaEventCreate(&e);
    std::vector<float> tot, pass;
    for (int it = 0; it < 8; ++it) {
        cudaMemcpyAsync(dk, k0, n * 4, cudaMemcpyDeviceToDevice, st);
        cudaMemsetAsync(dv, 0, n * 4, st);
        uint64_t launches = 0;
        sort_pairs(dk, dv, ak, av, n, sc, st, &launches, ev);
        cudaStreamSynchronize(st);
        if (it >= 3) {
            float t; cudaEventElapsedTime(&t, ev[0], ev[5]); tot.push_back(t);
            for (int p = 0; p < 4; ++p) { cudaEventElapsedTime(&t, ev[1 + p], ev[2 + p]); pass.push_back(t); }
        }
    }
    std::sort(tot.begin(), tot.end()); std::sort(pass.begin(), pass.end());
    const float t = tot[tot.size() / 2], p = pass[pass.size() / 2];
    printf("%-28s %s | 2^%d pairs: sort %.4f ms (%.1f Gpairs/s, %.3f of 6550 GB/s at 68 B/pair) | pass median %.4f ms (%.3f of 6550 at 16 B/pair) | %s\n", LAB_NAME,
           bad ? "WRONG" : "ok", lg, t, n / (t * 1e-3) / 1e9, 68.0 * n / (t * 1e-3) / 1e9 / 6550.1, p, 16.0 * n / (p * 1e-3) / 1e9 / 6550.1,
           cudaGetErrorString(cudaGetLastError()));
    return bad ? 1 : 0;
}
