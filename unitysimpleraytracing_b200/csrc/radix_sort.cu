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
// BLOCK x IPT pairs is ranked warp by warp -- peer lanes with the same digit find each other through
// one shared-memory atomicOr per key plus popc on lane masks (in place of the HLSL
// WavePrefixCountBits / WavePrefixSum one-bit splits; a ballot loop and match.any were measured and
// are 3x / 8x slower, tools/micro/match_bench.cu) -- per-digit tile offsets are chained across
// tiles by decoupled look-back (in place of the three Scan.compute dispatches), and pairs are
// staged in shared memory in digit order so the scatter writes coalesced runs.
// 16 B/pair/pass => 68 algorithmic bytes per pair. Measured: the pass kernel is bound by the SM's
// L1/LSU data pipe (bank conflicts of the five data-dependent shared-memory accesses per key), not
// by HBM: profiles/r01_summary.md.

#include <algorithm>
#include <cstdlib>

#include "usrt_internal.cuh"

namespace usrt {

namespace {

// Tile shapes. Large sorts are bound by per-key shared-memory work, so tiles are big (fewer look-back
// steps, longer coalesced runs per digit). Measured alternatives at 2^26 pairs: <256,16,4> 0.374 ms,
// <384,16,3> 0.372, <512,8,3> 0.401, <1024,8,1> 0.475 per pass against 0.360 for <512,16,2>. Very small
// sorts (< 2^18 pairs) are pure latency and use small tiles so that more CTAs run at once.
template <int BLOCK, int IPT, int CTAS> struct TileCfg {
    static constexpr int kBlock = BLOCK;            // threads per tile CTA
    static constexpr int kIPT = IPT;                // pairs per thread
    static constexpr int kTile = BLOCK * IPT;       // pairs per tile
    static constexpr int kWarps = BLOCK / 32;
    static constexpr int kCtasPerSM = CTAS;
    static constexpr int kMatchBytes = kWarps * kRadix * 8;
    static_assert(BLOCK >= kRadix, "one thread per digit in the scan / look-back step");
};
using BigTile = TileCfg<512, 16, 2>;                // 8192 pairs
using SmallTile = TileCfg<256, 8, 6>;               // 2048 pairs
constexpr uint64_t kSmallSortLimit = 1ull << 18;    // below this many pairs use SmallTile (measured: 2^20 is faster with BigTile)
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
// One persistent 1024-thread CTA per SM. Shared-memory counters are laid out [pass][digit][lane]:
// a lane only ever touches its own column, so every warp-wide atomicAdd is bank-conflict-free and
// never hits one address twice (the two things that serialise a plain [pass][digit] histogram).
constexpr int kHistThreads = 1024;
constexpr int kHistSmemBytes = kSortPasses * kRadix * 32 * 4;   // 128 KB

__global__ void __launch_bounds__(kHistThreads, 1) k_histogram(const uint32_t* __restrict__ keys, uint64_t n,
                                                               uint32_t* __restrict__ hist /* [4][256], zeroed */) {
    extern __shared__ uint32_t s_cnt[];                          // [4][256][32]
    const uint32_t tid = threadIdx.x, lane = tid & 31u;
    for (int i = tid; i < kSortPasses * kRadix * 32; i += kHistThreads) s_cnt[i] = 0;
    __syncthreads();

    uint32_t* col = s_cnt + lane;
    auto count = [&](uint32_t k) {
        atomicAdd(col + ((0 * kRadix + (k & 255u)) << 5), 1u);
        atomicAdd(col + ((1 * kRadix + ((k >> 8) & 255u)) << 5), 1u);
        atomicAdd(col + ((2 * kRadix + ((k >> 16) & 255u)) << 5), 1u);
        atomicAdd(col + ((3 * kRadix + (k >> 24)) << 5), 1u);
    };

    // scalar head up to 16-byte alignment, 128-bit body, scalar tail
    uint64_t head = ((16u - (uint32_t)(reinterpret_cast<uintptr_t>(keys) & 15u)) & 15u) >> 2;
    if (head > n) head = n;
    const uint64_t nvec = (n - head) >> 2;
    const uint4* __restrict__ vkeys = reinterpret_cast<const uint4*>(keys + head);
    const uint64_t gtid = (uint64_t)blockIdx.x * kHistThreads + tid;
    const uint64_t gstride = (uint64_t)gridDim.x * kHistThreads;
    uint64_t v = gtid;
    for (; v + 3 * gstride < nvec; v += 4 * gstride) {           // 4 independent 128-bit loads in flight
        const uint4 q0 = __ldg(vkeys + v), q1 = __ldg(vkeys + v + gstride);
        const uint4 q2 = __ldg(vkeys + v + 2 * gstride), q3 = __ldg(vkeys + v + 3 * gstride);
        count(q0.x); count(q0.y); count(q0.z); count(q0.w);
        count(q1.x); count(q1.y); count(q1.z); count(q1.w);
        count(q2.x); count(q2.y); count(q2.z); count(q2.w);
        count(q3.x); count(q3.y); count(q3.z); count(q3.w);
    }
    for (; v < nvec; v += gstride) {
        const uint4 q = __ldg(vkeys + v);
        count(q.x); count(q.y); count(q.z); count(q.w);
    }
    if (gtid < head) count(keys[gtid]);
    const uint64_t tail0 = head + (nvec << 2);
    if (tail0 + gtid < n) count(keys[tail0 + gtid]);
    __syncthreads();

    // thread t folds the 32 lane columns of bin t (rotated start => conflict-free) and publishes it
    uint32_t sum = 0;
#pragma unroll
    for (int l = 0; l < 32; ++l) sum += s_cnt[(tid << 5) + ((l + lane) & 31u)];
    if (sum) atomicAdd(&hist[tid], sum);
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
// Shared memory per CTA (dynamic, 50.6 KB => 4 CTAs per SM):
//   s_match [8 warps][256] x {mask, count}  16 KB   per-warp digit words (see ranking below); after
//                                                   ranking, count is rewritten to the warp's base slot
//   s_pairs [4096] x {key, value}           32 KB   the tile staged in digit order
//   s_global_off[256], s_scan[8], s_tile_id
//
// Ranking ("which lanes of my warp hold my digit, and how many did earlier rounds of this warp
// hold") replaces the HLSL WavePrefixCountBits/WavePrefixSum splits of LocalRadixSort.compute:29-91.
// Each lane ORs its lane bit into word[digit].mask with one shared-memory atomic; after a warp sync
// every lane reads the 64-bit word back: mask = this round's peers, count = peers of all earlier
// rounds. The lowest peer lane then writes {0, count + popc(mask)} -- clearing the mask and
// advancing the running count in one store. Measured on B200 this costs ~7 SM-cycles per warp-round
// against ~25 for an 8-ballot match and ~60 for match.any (tools/micro/match_bench.cu).
template <typename Cfg, bool kHasValues> struct PassSmem {
    static constexpr int kPairBytes = Cfg::kTile * (kHasValues ? 8 : 4);
    static constexpr int kTotal = Cfg::kMatchBytes + kPairBytes + kRadix * 4 + Cfg::kWarps * 4 + 16;
};

// kPeer: the multi-GPU bucket exchange. Instead of one output array, every digit has its own base address
// (key_ptrs[d] / val_ptrs[d], this rank's slice of the receive buffer of the GPU that owns bucket d, mapped
// through CUDA IPC): the stable scatter of the pass IS the all-to-all, written straight over NVLink.
template <typename Cfg, typename StatusT, bool kHasValues, bool kPeer = false>
__global__ void __launch_bounds__(Cfg::kBlock, Cfg::kCtasPerSM)
k_onesweep(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in, uint32_t* __restrict__ keys_out,
           uint32_t* __restrict__ vals_out, uint32_t n, int shift, const uint32_t* __restrict__ digit_base /* [256] */,
           uint32_t* __restrict__ tile_counter, StatusT* __restrict__ status /* [tiles][256], zeroed */,
           const unsigned long long* __restrict__ key_ptrs = nullptr, const unsigned long long* __restrict__ val_ptrs = nullptr) {
    using ST = StatusTraits<StatusT>;
    constexpr int kBlock = Cfg::kBlock, kIPT = Cfg::kIPT, kTile = Cfg::kTile, kWarps = Cfg::kWarps;
    constexpr int kMatchBytes = Cfg::kMatchBytes;
    extern __shared__ __align__(16) unsigned char smem[];
    uint2* s_match = reinterpret_cast<uint2*>(smem);                              // [kWarps][256]
    uint2* s_pairs = reinterpret_cast<uint2*>(smem + kMatchBytes);                // kHasValues
    uint32_t* s_keys = reinterpret_cast<uint32_t*>(smem + kMatchBytes);           // !kHasValues
    uint32_t* s_global_off = reinterpret_cast<uint32_t*>(smem + kMatchBytes + PassSmem<Cfg, kHasValues>::kPairBytes);
    uint32_t* s_scan = s_global_off + kRadix;
    uint32_t* s_tile_id = s_scan + kWarps;

    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;

    // dynamic tile id: a tile only ever waits on tiles that already started => forward progress
    if (tid == 0) *s_tile_id = atomicAdd(tile_counter, 1u);
    {
        uint4* z = reinterpret_cast<uint4*>(smem);
#pragma unroll
        for (int i = 0; i < kMatchBytes / 16 / kBlock; ++i) z[tid + i * kBlock] = make_uint4(0, 0, 0, 0);
    }
    __syncthreads();
    const uint32_t tile = *s_tile_id;
    const uint32_t tile_base = tile * (uint32_t)kTile;
    const uint32_t valid = min((uint32_t)kTile, n - tile_base);

    // warp-striped load: consecutive lanes read consecutive keys (one 128-B line per warp request);
    // item order (round i, lane) is the original order, which the ranking below preserves.
    const uint32_t warp_first = warp * (32u * kIPT);
    uint32_t key[kIPT];
#pragma unroll
    for (int i = 0; i < kIPT; ++i) {
        const uint32_t idx = warp_first + (uint32_t)i * 32u + lane;
        key[i] = idx < valid ? __ldg(keys_in + tile_base + idx) : 0xFFFFFFFFu;   // tail pads sort last, never stored
    }

    // stable rank of every key among the keys of its warp with the same digit (< 512: two per register)
    uint32_t rank2[kIPT / 2];
    const uint32_t lt = lanemask_lt();
    const uint32_t lane_bit = 1u << lane;
    uint2* my_match = s_match + warp * kRadix;
    uint32_t* my_words = reinterpret_cast<uint32_t*>(my_match);
    // Word d is the 8-byte pair my_match[d] = {mask, count}, with the two halves SWAPPED when bit 4 of
    // d is set: the 32-bit mask of digit d then lives in bank (2d + ((d >> 4) & 1)) mod 32, so the
    // atomics (and the count/base lookups, which use the other half) spread over all 32 banks.
#pragma unroll
    for (int i = 0; i < kIPT; ++i) {
        const uint32_t d = (key[i] >> shift) & 255u;
        const uint32_t sel = (d >> 4) & 1u;
        atomicOr(my_words + 2u * d + sel, lane_bit);
        __syncwarp();
        const uint2 w = my_match[d];                     // peers of this round + peers of earlier rounds
        __syncwarp();
        const uint32_t mask = sel ? w.y : w.x, prior = sel ? w.x : w.y;
        const uint32_t before = __popc(mask & lt);
        if (before == 0) {
            const uint32_t now = prior + __popc(mask);
            my_match[d] = sel ? make_uint2(now, 0u) : make_uint2(0u, now);
        }
        const uint32_t r = prior + before;
        if (i & 1) rank2[i >> 1] |= r << 16; else rank2[i >> 1] = r;
        __syncwarp();
    }
    __syncthreads();

    // values are fetched now, so their latency hides behind the scan and the look-back below
    uint32_t val[kHasValues ? kIPT : 1];
    if (kHasValues) {
#pragma unroll
        for (int i = 0; i < kIPT; ++i) {
            const uint32_t idx = warp_first + (uint32_t)i * 32u + lane;
            val[i] = idx < valid ? __ldg(vals_in + tile_base + idx) : 0u;
        }
    }

    // thread d (< 256) owns digit d: tile count, exclusive offsets across warps, look-back
    const uint32_t d = tid;
    uint32_t* s_words = reinterpret_cast<uint32_t*>(s_match);
    const uint32_t cnt_half = 1u - ((d >> 4) & 1u);      // which half of word d holds the count
    uint32_t count = 0, incl = 0, my_tile_start = 0;
    StatusT* my_status = status + (size_t)tile * kRadix + (d & 255u);
    if (d < kRadix) {
#pragma unroll
        for (int w = 0; w < kWarps; ++w) count += s_words[(w * kRadix + d) * 2 + cnt_half];
        // publish this tile's count of digit d as early as possible
        ST::store(my_status, (tile == 0 ? ST::kPrefix : ST::kAggregate) | (StatusT)count);
        incl = count;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if (lane >= (uint32_t)o) incl += y;
        }
        if (lane == 31) s_scan[warp] = incl;
    }
    __syncthreads();
    if (d < kRadix) {
        uint32_t wbase = 0;
#pragma unroll
        for (int w = 0; w < kRadix / 32; ++w) wbase += (w < (int)warp) ? s_scan[w] : 0u;
        const uint32_t tile_start = wbase + incl - count;     // first tile-local slot of digit d
        // in place: word[w][d].count becomes the first tile slot of warp w's keys of digit d
        uint32_t running = tile_start;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) {
            const uint32_t c = s_words[(w * kRadix + d) * 2 + cnt_half];
            s_words[(w * kRadix + d) * 2 + cnt_half] = running;
            running += c;
        }

        my_tile_start = tile_start;
    }
    __syncthreads();

    // stage the tile in digit order
#pragma unroll
    for (int i = 0; i < kIPT; ++i) {
        const uint32_t dg = (key[i] >> shift) & 255u;
        const uint32_t slot = my_words[2u * dg + 1u - ((dg >> 4) & 1u)] + ((rank2[i >> 1] >> ((i & 1) * 16)) & 0xFFFFu);
        if (kHasValues) s_pairs[slot] = make_uint2(key[i], val[i]);
        else s_keys[slot] = key[i];
    }
    // Look-back AFTER staging: the aggregate was published before the scan, so by now the preceding
    // tiles have usually posted their inclusive prefixes and the walk resolves in one round trip.
    if (d < kRadix) {
        const uint32_t tile_start = my_tile_start;
        // decoupled look-back over the preceding tiles' counts of this digit, four tiles per round
        // trip (the loads are independent; only the accumulation is ordered)
        uint32_t exclusive = 0;
        if (tile > 0) {
            int32_t t = (int32_t)tile - 1;
            bool done = false;
            while (!done) {
                StatusT s4[4];
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    s4[k] = (t - k >= 0) ? ST::load(status + (size_t)(t - k) * kRadix + d) : (StatusT)ST::kPrefix;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (!done) {
                        while ((s4[k] & ST::kFlagMask) == 0) s4[k] = ST::load(status + (size_t)(t - k) * kRadix + d);
                        exclusive += (uint32_t)(s4[k] & ST::kValueMask);
                        done = (s4[k] & ST::kPrefix) != 0;
                    }
                }
                t -= 4;
            }
            ST::store(my_status, ST::kPrefix | (StatusT)(exclusive + count));
        }
        s_global_off[d] = (kPeer ? 0u : digit_base[d]) + exclusive - tile_start;   // wraps mod 2^32 by design
    }
    __syncthreads();

    // coalesced runs out: slot p of the tile goes to global_off[digit] + p
#pragma unroll
    for (int i = 0; i < kIPT; ++i) {
        const uint32_t p = tid + (uint32_t)i * kBlock;
        if (p < valid) {
            if (kHasValues) {
                const uint2 kv = s_pairs[p];
                const uint32_t dg = (kv.x >> shift) & 255u;
                const uint32_t dst = s_global_off[dg] + p;
                if (kPeer) {
                    reinterpret_cast<uint32_t*>(__ldg(key_ptrs + dg))[dst] = kv.x;
                    reinterpret_cast<uint32_t*>(__ldg(val_ptrs + dg))[dst] = kv.y;
                } else {
                    keys_out[dst] = kv.x;
                    vals_out[dst] = kv.y;
                }
            } else {
                const uint32_t k = s_keys[p];
                keys_out[s_global_off[(k >> shift) & 255u] + p] = k;
            }
        }
    }
}

inline uint32_t histogram_grid(uint64_t count) {
    cudaFuncSetAttribute(k_histogram, cudaFuncAttributeMaxDynamicSharedMemorySize, kHistSmemBytes);   // per device, cheap
    const uint64_t vec_work = (count + 4 * kHistThreads - 1) / (4 * kHistThreads);
    return (uint32_t)std::min<uint64_t>(std::max<uint64_t>(vec_work, 1), (uint64_t)kNumSMs);
}
inline uint32_t tile_pairs(uint64_t count) { return count < kSmallSortLimit ? SmallTile::kTile : BigTile::kTile; }
inline uint32_t num_tiles(uint64_t count) { return (uint32_t)((count + tile_pairs(count) - 1) / tile_pairs(count)); }
inline bool wide_status(uint64_t count) {
    static const bool forced = getenv("USRT_FORCE_WIDE_STATUS") != nullptr;   // test hook: 64-bit look-back words at any size
    return forced || count >= (1ull << 30);
}
inline uint64_t status_words_bytes(uint64_t count) { return (uint64_t)num_tiles(count) * kRadix * (wide_status(count) ? 8 : 4); }

template <typename Cfg, typename StatusT, bool kHasValues>
cudaError_t launch_pass(const uint32_t* ki, const uint32_t* vi, uint32_t* ko, uint32_t* vo, uint64_t count, int shift,
                        const uint32_t* digit_base, uint32_t* tile_counter, void* status, cudaStream_t stream) {
    constexpr int smem = PassSmem<Cfg, kHasValues>::kTotal;
    cudaFuncSetAttribute(k_onesweep<Cfg, StatusT, kHasValues>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    k_onesweep<Cfg, StatusT, kHasValues><<<num_tiles(count), Cfg::kBlock, smem, stream>>>(
        ki, vi, ko, vo, (uint32_t)count, shift, digit_base, tile_counter, static_cast<StatusT*>(status));
    return cudaGetLastError();
}

template <typename StatusT>
cudaError_t run_pass(const uint32_t* ki, const uint32_t* vi, uint32_t* ko, uint32_t* vo, uint64_t count, int shift,
                     const uint32_t* digit_base, uint32_t* tile_counter, void* status, cudaStream_t stream) {
    const bool small = count < kSmallSortLimit;
    if (vi != nullptr)
        return small ? launch_pass<SmallTile, StatusT, true>(ki, vi, ko, vo, count, shift, digit_base, tile_counter, status, stream)
                     : launch_pass<BigTile, StatusT, true>(ki, vi, ko, vo, count, shift, digit_base, tile_counter, status, stream);
    return small ? launch_pass<SmallTile, StatusT, false>(ki, vi, ko, vo, count, shift, digit_base, tile_counter, status, stream)
                 : launch_pass<BigTile, StatusT, false>(ki, vi, ko, vo, count, shift, digit_base, tile_counter, status, stream);
}

}  // namespace

cudaError_t sort_scratch_reserve(SortScratch& s, uint64_t count, bool need_alt) {
    cudaError_t e;
    if (s.hist == nullptr) {
        if ((e = cudaMalloc(&s.hist, kSortPasses * kRadix * sizeof(uint32_t))) != cudaSuccess) return e;
        ++s.generation;
    }
    const uint64_t need = kHeaderWords * 4 + kSortPasses * status_words_bytes(count);
    if (need > s.status_bytes) {
        if (s.status) cudaFree(s.status);
        s.status = nullptr; s.status_bytes = 0;
        if ((e = cudaMalloc(&s.status, need)) != cudaSuccess) return e;
        s.status_bytes = need;
        ++s.generation;
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
                       SortScratch& s, cudaStream_t stream, uint64_t* launches, cudaEvent_t* events) {
    if (count == 0) return cudaSuccess;
    cudaError_t e;
    if ((e = sort_scratch_reserve(s, count, false)) != cudaSuccess) return e;
    const uint64_t pass_bytes = status_words_bytes(count);
    if (events && (e = cudaEventRecord(events[0], stream)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(s.hist, 0, kSortPasses * kRadix * 4, stream)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(s.status, 0, kHeaderWords * 4 + kSortPasses * pass_bytes, stream)) != cudaSuccess) return e;

    k_histogram<<<histogram_grid(count), kHistThreads, kHistSmemBytes, stream>>>(keys, count, s.hist);
    k_scan_histogram<<<1, kSortPasses * kRadix, 0, stream>>>(s.hist, nullptr, -1);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    if (launches) *launches += 2;
    if (events && (e = cudaEventRecord(events[1], stream)) != cudaSuccess) return e;

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
        if (events && (e = cudaEventRecord(events[2 + pass], stream)) != cudaSuccess) return e;
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
    k_histogram<<<histogram_grid(count), kHistThreads, kHistSmemBytes, stream>>>(src_keys, count, s.hist);
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

// Raw counts of one digit (no scan): the first half of the multi-GPU bucket exchange.
cudaError_t digit_histogram(const uint32_t* keys, uint64_t count, int bit_offset, uint32_t* hist_out, SortScratch& s,
                            cudaStream_t stream, uint64_t* launches) {
    cudaError_t e;
    if ((e = sort_scratch_reserve(s, std::max<uint64_t>(count, 1), false)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(s.hist, 0, kSortPasses * kRadix * 4, stream)) != cudaSuccess) return e;
    if (count) {
        k_histogram<<<histogram_grid(count), kHistThreads, kHistSmemBytes, stream>>>(keys, count, s.hist);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
        if (launches) *launches += 1;
    }
    return cudaMemcpyAsync(hist_out, s.hist + (bit_offset / kRadixBits) * kRadix, kRadix * 4, cudaMemcpyDeviceToDevice, stream);
}

// The stable partition pass with one destination base address per digit (peer memory allowed).
cudaError_t partition_scatter(const uint32_t* src_keys, const uint32_t* src_vals, uint64_t count, int bit_offset,
                              const unsigned long long* key_ptrs, const unsigned long long* val_ptrs, SortScratch& s,
                              cudaStream_t stream, uint64_t* launches) {
    if (count == 0) return cudaSuccess;
    cudaError_t e;
    if ((e = sort_scratch_reserve(s, count, false)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(s.status, 0, kHeaderWords * 4 + status_words_bytes(count), stream)) != cudaSuccess) return e;
    uint32_t* counters = static_cast<uint32_t*>(s.status);
    uint32_t* st = reinterpret_cast<uint32_t*>(static_cast<char*>(s.status) + kHeaderWords * 4);
    if (count < kSmallSortLimit) {
        constexpr int smem = PassSmem<SmallTile, true>::kTotal;
        cudaFuncSetAttribute(k_onesweep<SmallTile, uint32_t, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        k_onesweep<SmallTile, uint32_t, true, true><<<num_tiles(count), SmallTile::kBlock, smem, stream>>>(
            src_keys, src_vals, nullptr, nullptr, (uint32_t)count, bit_offset, nullptr, counters, st, key_ptrs, val_ptrs);
    } else {
        constexpr int smem = PassSmem<BigTile, true>::kTotal;
        cudaFuncSetAttribute(k_onesweep<BigTile, uint32_t, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        k_onesweep<BigTile, uint32_t, true, true><<<num_tiles(count), BigTile::kBlock, smem, stream>>>(
            src_keys, src_vals, nullptr, nullptr, (uint32_t)count, bit_offset, nullptr, counters, st, key_ptrs, val_ptrs);
    }
    if (launches) *launches += 1;
    return cudaGetLastError();
}

}  // namespace usrt
