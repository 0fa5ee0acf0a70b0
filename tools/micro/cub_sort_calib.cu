// calibration point (developer tool, NOT product code, not linked into libusrt_b200.so): device time of
// cub::DeviceRadixSort::SortPairs (CUDA 12.9's CCCL, onesweep, Policy1000) on the same random uint32 pairs and the
// same box as this repo's k_onesweep sort, so "fraction of the HBM roof" has a library reference beside it.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a tools/micro/cub_sort_calib.cu -o tools/micro/cub_sort_calib
#include <cstdio>
#include <cstdint>
#include <vector>
#include <algorithm>
#include <cub/device/device_radix_sort.cuh>

__global__ void fill(uint32_t* k, uint32_t* v, uint64_t n, uint64_t seed) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t x = (i + seed) * 0x9E3779B97F4A7C15ull;
        x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull; x = (x ^ (x >> 27)) * 0x94D049BB133111EBull; x ^= x >> 31;
        k[i] = (uint32_t)x; v[i] = (uint32_t)i;
    }
}

int main() {
    for (int lg : {20, 24, 26, 28}) {
        const uint64_t n = 1ull << lg;
        uint32_t *k0, *v0, *k1, *v1;
        cudaMalloc(&k0, n * 4); cudaMalloc(&v0, n * 4); cudaMalloc(&k1, n * 4); cudaMalloc(&v1, n * 4);
        void* tmp = nullptr; size_t tmp_bytes = 0;
        cub::DoubleBuffer<uint32_t> dk(k0, k1), dv(v0, v1);
        cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, dk, dv, (int64_t)n);
        cudaMalloc(&tmp, tmp_bytes);
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        std::vector<float> ms;
        for (int it = 0; it < 8; ++it) {
            fill<<<148 * 8, 256>>>(k0, v0, n, 77 + it);
            cub::DoubleBuffer<uint32_t> xk(k0, k1), xv(v0, v1);
            cudaEventRecord(a);
            cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, xk, xv, (int64_t)n);
            cudaEventRecord(b); cudaEventSynchronize(b);
            float t; cudaEventElapsedTime(&t, a, b);
            if (it >= 3) ms.push_back(t);
        }
        std::sort(ms.begin(), ms.end());
        const float med = ms[ms.size() / 2];
        printf("cub SortPairs 2^%d uint32 pairs: %.4f ms median = %.1f Gpairs/s = %.0f GB/s at 68 B/pair (%.3f of 6550 GB/s) | %s\n", lg, med,
               n / (med * 1e-3) / 1e9, 68.0 * n / (med * 1e-3) / 1e9, 68.0 * n / (med * 1e-3) / 1e9 / 6550.1, cudaGetErrorString(cudaGetLastError()));
        cudaFree(k0); cudaFree(v0); cudaFree(k1); cudaFree(v1); cudaFree(tmp);
    }
    return 0;
}
