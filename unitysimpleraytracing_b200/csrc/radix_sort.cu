// radix_sort.cu -- K2: stable LSD radix sort of (uint32 key, uint32 value) pairs, 4 passes x 8 bits.
//
// Replaces ComputeBufferSorter.Sort() (Assets/_Scripts/ComputeBufferSorter.cs:100-126) and its five
// HLSL kernels per pass: LocalRadixSort (Assets/_Shaders/Sorting/LocalRadixSort.compute:53-134),
// PreScan / BlockSum / GlobalScan (Scan.compute:15-96) and GlobalRadixSort
// (GlobalRadixSort.compute:20-40). Same key width (32), digit width (8), pass order (bitOffset
// 0, 8, 16, 24) and the same net contract: a STABLE ascending sort by the full key.
//
// B200 design (not a port): the reference moves >= 32 B/pair/pass through a block-sorted
// intermediate and three scan dispatches. Here one upfront kernel reads the keys once and builds
// all four digit histograms (4 B/pair), and each pass is ONE kernel ("onesweep"): a tile of
// BLOCK x IPT pairs is ranked with warp ballots (in place of the HLSL WavePrefixCountBits /
// WavePrefixSum), per-digit tile offsets are chained across tiles by decoupled look-back (in place
// of the three Scan.compute dispatches), and pairs are staged in shared memory in digit order so
// the scatter writes coalesced runs. 16 B/pair/pass => 68 algorithmic bytes per pair.

#include "usrt_internal.cuh"

namespace usrt {

namespace {

constexpr int kBlock = 256;                 // threads per tile CTA (8 warps)
constexpr int kIPT = 16;                    // pairs per thread
constexpr int kTile = kBlock * kIPT;        // 4096 pairs per tile
constexpr int kWarps = kBlock / 32;
constexpr uint32_t kHeaderWords = 64;       // tile counters live in the first 256 B of the status buffer

// look-back status word: flag in the top bits, running count below. 32-bit words hold counts
// < 2^30; sorts of >= 2^30 pairs use 64-bit words.
template <typename T> struct StatusTraits;
template <> struct StatusTraits<uint32_t> {
    static constexpr uint32_t kAggregate = 1u << 30, kPrefix = 2u << 30, kFlagMask = 3u << 30, kValueMask = (1u << 30) - 1;
    __device__ static __forceinline__ uint32_t load(const uint32_t* p) { return ld_relaxed_u32(p); }
    __device__ static __forceinline__ void store(uint32_t* p, uint32_t v) { st_relaxed_u32(p, v); }
};
template <> struct StatusTraits<uint64_t> {
    static constexpr uint64_t kAggregate = 1ull << 32, kPrefix = 2ull << 32, kFlagMask = 3ull << 32, kValueMask = 0xFFFFFFFFull;
    __device__ static __forceinline__ uint64_t load(const uint64_t* p) { return ld_relaxed_u64(p); }
    __device__ static __forceinline__ void store(uint64_t* p, uint64_t v) { st_relaxed_u64(p, v); }
};

// ---- upfront histogram: all four digits from one read of the keys --------------------------------
__global__ void __launch_bounds__(256) k_histogram(const uint32_t* __restrict__ keys, uint64_t n,
                                                   uint32_t* __restrict__ hist /* [4][256], zeroed */) {
    __shared__ uint32_t s_hist[kSortPasses * kRadix];
    for (int i = threadIdx.x; i < kSortPasses * kRadix; i += blockDim.x) s_hist[i] = 0;
    __syncthreads();

    auto count = [&](uint32_t k) {
        atomicAdd(&s_hist[0 * kRadix + (k & 255u)], 1u);
        atomicAdd(&s_hist[1 * kRadix + ((k >> 8) & 255u)], 1u);
        atomicAdd(&s_hist[2 * kRadix + ((k >> 16) & 255u)], 1u);
        atomicAdd(&s_hist[3 * kRadix + (k >> 24)], 1u);
    };

    // scalar head up to 16-byte alignment, 128-bit body, scalar tail
    uint64_t head = ((16u - (uint32_t)(reinterpret_cast<uintptr_t>(keys) & 15u)) & 15u) >> 2;
    if (head > n) head = n;
    const uint64_t nvec = (n - head) >> 2;
    const uint4* __restrict__ vkeys = reinterpret_cast<const uint4*>(keys + head);
    const uint64_t gtid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t gstride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t v = gtid; v < nvec; v += gstride) {
        const uint4 q = __ldg(vkeys + v);
        count(q.x); count(q.y); count(q.z); count(q.w);
    }
    if (gtid < head) count(keys[gtid]);
    const uint64_t tail0 = head + (nvec << 2);
    if (tail0 + gtid < n) count(keys[tail0 + gtid]);
    __syncthreads();
    for (int i = threadIdx.x; i < kSortPasses * kRadix; i += blockDim.x) {
        const uint32_t c = s_hist[i];
        if (c) atomicAdd(&hist[i], c);
    }
}

// exclusive scan of each pass's 256 counts -> first output position of every digit (in place).
// optional copy of the raw counts of one pass (for the multi-GPU bucket split).
__global__ void __launch_bounds__(kSortPasses * kRadix) k_scan_histogram(uint32_t* __restrict__ hist,
                                                                         uint32_t* __restrict__ raw_out, int raw_pass) {
    __shared__ uint32_t s_warp[kSortPasses * kRadix / 32];
    const uint32_t t = threadIdx.x, lane = t & 31u, warp = t >> 5;
    const uint32_t c = hist[t];
    if (raw_out != nullptr && (int)(t >> 8) == raw_pass) raw_out[t & 255u] = c;
    uint32_t incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if (lane >= (uint32_t)o) incl += y;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    uint32_t base = 0;
    const uint32_t first_warp = (t >> 8) * (kRadix / 32);   // 8 warps per pass
    for (uint32_t w = first_warp; w < warp; ++w) base += s_warp[w];
    hist[t] = base + incl - c;
}

// ---- one radix pass: rank + look-back + staged stable scatter ------------------------------------
template <typename StatusT, bool kHasValues>
__global__ void __launch_bounds__(kBlock, 4)
k_onesweep(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in, uint32_t* __restrict__ keys_out,
           uint32_t* __restrict__ vals_out, uint32_t n, int shift, const uint32_t* __restrict__ digit_base /* [256] */,
           uint32_t* __restrict__ tile_counter, StatusT* __restrict__ status /* [tiles][256], zeroed */) {
    using ST = StatusTraits<StatusT>;
    __shared__ uint32_t s_warp_hist[kWarps][kRadix];   // per-warp digit counts, then exclusive warp offsets
    __shared__ uint32_t s_keys[kTile];
    __shared__ uint32_t s_vals[kHasValues ? kTile : 1];
    __shared__ uint32_t s_tile_start[kRadix];          // first tile-local slot of each digit
    __shared__ uint32_t s_global_off[kRadix];          // global position of slot 0 of each digit's run, minus tile_start
    __shared__ uint32_t s_scan[kWarps];
    __shared__ uint32_t s_tile_id;

    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;

    // dynamic tile id: a tile only ever waits on tiles that already started => forward progress
    if (tid == 0) s_tile_id = atomicAdd(tile_counter, 1u);
#pragma unroll
    for (int i = tid; i < kWarps * kRadix; i += kBlock) (&s_warp_hist[0][0])[i] = 0;
    __syncthreads();
    const uint32_t tile = s_tile_id;
    const uint32_t tile_base = tile * (uint32_t)kTile;
    const uint32_t valid = min((uint32_t)kTile, n - tile_base);

    // warp-striped load: consecutive lanes read consecutive keys (one 128-B line per warp request);
    // item order (i, lane) is the original order, which the ranking below preserves.
    const uint32_t warp_first = warp * (32u * kIPT);
    uint32_t key[kIPT];
#pragma unroll
    for (int i = 0; i < kIPT; ++i) {
        const uint32_t idx = warp_first + (uint32_t)i * 32u + lane;
        key[i] = idx < valid ? __ldg(keys_in + tile_base + idx) : 0xFFFFFFFFu;   // tail pads sort last, never stored
    }
    uint32_t val[kHasValues ? kIPT : 1];
    if (kHasValues) {
#pragma unroll
        for (int i = 0; i < kIPT; ++i) {
            const uint32_t idx = warp_first + (uint32_t)i * 32u + lane;
            val[i] = idx < valid ? __ldg(vals_in + tile_base + idx) : 0u;
        }
    }

    // stable rank of every key among the keys of its warp with the same digit, by warp ballots
    uint32_t rank[kIPT];
    const uint32_t lt = lanemask_lt();
    uint32_t* my_hist = s_warp_hist[warp];
#pragma unroll
    for (int i = 0; i < kIPT; ++i) {
        const uint32_t d = (key[i] >> shift) & 255u;
        uint32_t peers = 0xFFFFFFFFu;
#pragma unroll
        for (int b = 0; b < kRadixBits; ++b) {
            const bool bit = (d >> b) & 1u;
            const uint32_t ballot = __ballot_sync(0xFFFFFFFFu, bit);
            peers &= bit ? ballot : ~ballot;
        }
        const uint32_t before = __popc(peers & lt);
        const int leader = __ffs(peers) - 1;
        uint32_t prior = 0;
        if (before == 0) {                       // lowest lane holding this digit
            prior = my_hist[d];
            my_hist[d] = prior + __popc(peers);
        }
        prior = __shfl_sync(0xFFFFFFFFu, prior, leader);
        rank[i] = prior + before;
        __syncwarp();
    }
    __syncthreads();

    // thread d owns digit d: exclusive offsets across warps, tile count, look-back, tile-local start
    {
        const uint32_t d = tid;                  // kBlock == kRadix
        uint32_t count = 0;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) {
            const uint32_t c = s_warp_hist[w][d];
            s_warp_hist[w][d] = count;
            count += c;
        }
        StatusT* my_status = status + (size_t)tile * kRadix + d;
        ST::store(my_status, (tile == 0 ? ST::kPrefix : ST::kAggregate) | (StatusT)count);

        // block exclusive scan of the tile counts over digits
        uint32_t incl = count;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if (lane >= (uint32_t)o) incl += y;
        }
        if (lane == 31) s_scan[warp] = incl;
        __syncthreads();
        uint32_t wbase = 0;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) wbase += (w < (int)warp) ? s_scan[w] : 0u;
        const uint32_t tile_start = wbase + incl - count;
        s_tile_start[d] = tile_start;

        // decoupled look-back over the preceding tiles' counts of this digit
        uint32_t exclusive = 0;
        if (tile > 0) {
            const StatusT* look = my_status - kRadix;
            while (true) {
                StatusT s;
                do { s = ST::load(look); } while ((s & ST::kFlagMask) == 0);
                exclusive += (uint32_t)(s & ST::kValueMask);
                if (s & ST::kPrefix) break;
                look -= kRadix;
            }
            ST::store(my_status, ST::kPrefix | (StatusT)(exclusive + count));
        }
        s_global_off[d] = digit_base[d] + exclusive - tile_start;   // wraps mod 2^32 by design
    }
    __syncthreads();

    // stage the tile in digit order
#pragma unroll
    for (int i = 0; i < kIPT; ++i) {
        const uint32_t d = (key[i] >> shift) & 255u;
        const uint32_t slot = s_tile_start[d] + my_hist[d] + rank[i];
        s_keys[slot] = key[i];
        if (kHasValues) s_vals[slot] = val[i];
    }
    __syncthreads();

    // coalesced runs out: slot p of the tile goes to global_off[digit] + p
#pragma unroll
    for (int i = 0; i < kIPT; ++i) {
        const uint32_t p = tid + (uint32_t)i * kBlock;
        if (p < valid) {
            const uint32_t k = s_keys[p];
            const uint32_t dst = s_global_off[(k >> shift) & 255u] + p;
            keys_out[dst] = k;
            if (kHasValues) vals_out[dst] = s_vals[p];
        }
    }
}

inline uint32_t num_tiles(uint64_t count) { return (uint32_t)((count + kTile - 1) / kTile); }
inline bool wide_status(uint64_t count) { return count >= (1ull << 30); }
inline uint64_t status_words_bytes(uint64_t count) { return (uint64_t)num_tiles(count) * kRadix * (wide_status(count) ? 8 : 4); }

template <typename StatusT>
cudaError_t run_pass(const uint32_t* ki, const uint32_t* vi, uint32_t* ko, uint32_t* vo, uint64_t count, int shift,
                     const uint32_t* digit_base, uint32_t* tile_counter, void* status, cudaStream_t stream) {
    const uint32_t tiles = num_tiles(count);
    if (vi != nullptr)
        k_onesweep<StatusT, true><<<tiles, kBlock, 0, stream>>>(ki, vi, ko, vo, (uint32_t)count, shift, digit_base,
                                                               tile_counter, static_cast<StatusT*>(status));
    else
        k_onesweep<StatusT, false><<<tiles, kBlock, 0, stream>>>(ki, vi, ko, vo, (uint32_t)count, shift, digit_base,
                                                                tile_counter, static_cast<StatusT*>(status));
    return cudaGetLastError();
}

}  // namespace

cudaError_t sort_scratch_reserve(SortScratch& s, uint64_t count, bool need_alt) {
    cudaError_t e;
    if (s.hist == nullptr) {
        if ((e = cudaMalloc(&s.hist, kSortPasses * kRadix * sizeof(uint32_t))) != cudaSuccess) return e;
    }
    const uint64_t need = kHeaderWords * 4 + kSortPasses * status_words_bytes(count);
    if (need > s.status_bytes) {
        if (s.status) cudaFree(s.status);
        s.status = nullptr; s.status_bytes = 0;
        if ((e = cudaMalloc(&s.status, need)) != cudaSuccess) return e;
        s.status_bytes = need;
    }
    if (need_alt && count > s.alt_capacity) {
        if (s.keys_alt) cudaFree(s.keys_alt);
        if (s.vals_alt) cudaFree(s.vals_alt);
        s.keys_alt = s.vals_alt = nullptr; s.alt_capacity = 0;
        if ((e = cudaMalloc(&s.keys_alt, count * 4)) != cudaSuccess) return e;
        if ((e = cudaMalloc(&s.vals_alt, count * 4)) != cudaSuccess) return e;
        s.alt_capacity = count;
    }
    return cudaSuccess;
}

void sort_scratch_free(SortScratch& s) {
    if (s.hist) cudaFree(s.hist);
    if (s.status) cudaFree(s.status);
    if (s.keys_alt) cudaFree(s.keys_alt);
    if (s.vals_alt) cudaFree(s.vals_alt);
    s = SortScratch();
}

cudaError_t sort_pairs(uint32_t* keys, uint32_t* vals, uint32_t* keys_alt, uint32_t* vals_alt, uint64_t count,
                       SortScratch& s, cudaStream_t stream, uint64_t* launches) {
    if (count == 0) return cudaSuccess;
    cudaError_t e;
    if ((e = sort_scratch_reserve(s, count, false)) != cudaSuccess) return e;
    const uint64_t pass_bytes = status_words_bytes(count);
    if ((e = cudaMemsetAsync(s.hist, 0, kSortPasses * kRadix * 4, stream)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(s.status, 0, kHeaderWords * 4 + kSortPasses * pass_bytes, stream)) != cudaSuccess) return e;

    const uint64_t vec_work = (count + 4 * 256 - 1) / (4 * 256);
    const uint32_t hgrid = (uint32_t)std::min<uint64_t>(std::max<uint64_t>(vec_work, 1), (uint64_t)kNumSMs * 8);
    k_histogram<<<hgrid, 256, 0, stream>>>(keys, count, s.hist);
    k_scan_histogram<<<1, kSortPasses * kRadix, 0, stream>>>(s.hist, nullptr, -1);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    if (launches) *launches += 2;

    uint32_t* counters = static_cast<uint32_t*>(s.status);
    char* status0 = static_cast<char*>(s.status) + kHeaderWords * 4;
    const uint32_t* ki = keys; const uint32_t* vi = vals;
    uint32_t* ko = keys_alt; uint32_t* vo = vals_alt;
    for (int pass = 0; pass < kSortPasses; ++pass) {               // bitOffset = 0, 8, 16, 24
        void* st = status0 + (uint64_t)pass * pass_bytes;
        if (wide_status(count))
            e = run_pass<uint64_t>(ki, vi, ko, vo, count, pass * kRadixBits, s.hist + pass * kRadix, counters + pass, st, stream);
        else
            e = run_pass<uint32_t>(ki, vi, ko, vo, count, pass * kRadixBits, s.hist + pass * kRadix, counters + pass, st, stream);
        if (e != cudaSuccess) return e;
        if (launches) *launches += 1;
        const uint32_t* tk = ki; const uint32_t* tv = vi;
        ki = ko; vi = vo;
        ko = const_cast<uint32_t*>(tk); vo = const_cast<uint32_t*>(tv);
    }
    return cudaSuccess;   // even number of passes: the result is back in (keys, vals)
}

cudaError_t partition_pass(const uint32_t* src_keys, const uint32_t* src_vals, uint32_t* dst_keys, uint32_t* dst_vals,
                           uint64_t count, int bit_offset, uint32_t* histogram_out, SortScratch& s,
                           cudaStream_t stream, uint64_t* launches) {
    cudaError_t e;
    if ((e = sort_scratch_reserve(s, std::max<uint64_t>(count, 1), false)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(s.hist, 0, kSortPasses * kRadix * 4, stream)) != cudaSuccess) return e;
    if (count == 0) {
        if (histogram_out) return cudaMemsetAsync(histogram_out, 0, kRadix * 4, stream);
        return cudaSuccess;
    }
    const uint64_t pass_bytes = status_words_bytes(count);
    if ((e = cudaMemsetAsync(s.status, 0, kHeaderWords * 4 + pass_bytes, stream)) != cudaSuccess) return e;
    const int pass = bit_offset / kRadixBits;
    const uint64_t vec_work = (count + 4 * 256 - 1) / (4 * 256);
    const uint32_t hgrid = (uint32_t)std::min<uint64_t>(std::max<uint64_t>(vec_work, 1), (uint64_t)kNumSMs * 8);
    k_histogram<<<hgrid, 256, 0, stream>>>(src_keys, count, s.hist);
    k_scan_histogram<<<1, kSortPasses * kRadix, 0, stream>>>(s.hist, histogram_out, pass);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    uint32_t* counters = static_cast<uint32_t*>(s.status);
    void* st = static_cast<char*>(s.status) + kHeaderWords * 4;
    if (wide_status(count))
        e = run_pass<uint64_t>(src_keys, src_vals, dst_keys, dst_vals, count, bit_offset, s.hist + pass * kRadix, counters, st, stream);
    else
        e = run_pass<uint32_t>(src_keys, src_vals, dst_keys, dst_vals, count, bit_offset, s.hist + pass * kRadix, counters, st, stream);
    if (launches) *launches += 3;
    return e;
}

}  // namespace usrt
