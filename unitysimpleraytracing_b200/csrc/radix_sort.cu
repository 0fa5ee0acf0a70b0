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
// BLOCK x IPT pairs is ranked in 16-lane groups with one returning shared-memory atomicAdd per key (in
// place of the HLSL WavePrefixCountBits / WavePrefixSum one-bit splits; ballot loops, match.any and
// atomicOr + read-back matching were measured and are 2x-10x slower, tools/micro/*_bench.cu) -- per-digit
// tile offsets are chained across tiles by decoupled look-back (in place of the three Scan.compute
// dispatches), and pairs are staged in shared memory in digit order so the scatter writes coalesced runs.
// 16 B/pair/pass => 68 algorithmic bytes per pair. Measured: the pass kernel is bound by the SM's
// L1/LSU data pipe (one wavefront per cycle; bank conflicts of the data-dependent shared-memory accesses
// per key), not by HBM: profiles/r02_summary.md.

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
    static_assert(BLOCK % kRadix == 0, "whole threads per digit in the scan / look-back step");
};
#ifndef USRT_BIG_BLOCK                                // (tools/micro/sort_lab.cu builds other shapes side by side)
#define USRT_BIG_BLOCK 256
#define USRT_BIG_IPT 24
#define USRT_BIG_CTAS 3
#endif
using BigTile = TileCfg<USRT_BIG_BLOCK, USRT_BIG_IPT, USRT_BIG_CTAS>;   // 8192 pairs
using SmallTile = TileCfg<256, 8, 6>;               // 2048 pairs
constexpr uint64_t kSmallSortLimit = 1ull << 18;    // below this many pairs use SmallTile (measured: 2^20 is faster with BigTile)
constexpr uint64_t kPairsSortLimit = 1ull << 22;    // from this many pairs on, intermediate passes move interleaved records
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

// 64-bit keys (ComputeBufferSorter<ulong, uint>): the same four-digit histogram over ONE 32-bit half of every key,
// launched once per half (the [pass][digit][lane] counters of four digits already fill 128 KB of shared memory).
__global__ void __launch_bounds__(kHistThreads, 1) k_histogram64(const uint2* __restrict__ keys /* {low, high} */, uint64_t n,
                                                                 int half, uint32_t* __restrict__ hist /* [4][256], zeroed */) {
    extern __shared__ uint32_t s_cnt[];                          // [4][256][32]
    const uint32_t tid = threadIdx.x, lane = tid & 31u;
    for (int i = tid; i < kSortPasses * kRadix * 32; i += kHistThreads) s_cnt[i] = 0;
    __syncthreads();
    uint32_t* col = s_cnt + lane;
    const uint64_t gstride = (uint64_t)gridDim.x * kHistThreads;
    for (uint64_t i = (uint64_t)blockIdx.x * kHistThreads + tid; i < n; i += gstride) {
        const uint2 q = __ldg(keys + i);
        const uint32_t k = half ? q.y : q.x;
        atomicAdd(col + ((0 * kRadix + (k & 255u)) << 5), 1u);
        atomicAdd(col + ((1 * kRadix + ((k >> 8) & 255u)) << 5), 1u);
        atomicAdd(col + ((2 * kRadix + ((k >> 16) & 255u)) << 5), 1u);
        atomicAdd(col + ((3 * kRadix + (k >> 24)) << 5), 1u);
    }
    __syncthreads();
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
// Shared memory per CTA (dynamic; BigTile = 256 threads x 24 pairs: 58 KB => 3 CTAs per SM, 80 registers per thread):
//   s_tbl  [groups][256] u32            8 KB   one word per (ranking group, digit); a group is a whole warp (default) or a
//                                              16-lane half (fully verified flavour: twice the tables): the group's running
//                                              count of that digit; after the scan: first tile slot of the group's keys of it
//   s_pairs [tile] x {key, value}      48 KB   the tile staged in digit order
//   s_global_off[256], s_part[H][256], s_scan[warps], s_tile_id[2] (this tile's and the next ticket), s_canary[8]
// The CTAs are persistent (grid = SMs x CTAs per SM): each loops over tile tickets drawn from a global counter.
//
// Ranking ("how many keys of my digit precede mine in the tile") replaces the HLSL WavePrefixCountBits /
// WavePrefixSum one-bit splits of LocalRadixSort.compute:29-91. A 16-lane group owns 16*IPT CONSECUTIVE keys of
// the tile (round i = 16 consecutive keys) and one 256-word table. In round i every lane does ONE shared-memory
// atomic, old = atomicAdd(&tbl[digit], lane_bit << 16 | 1): old.count is the number of keys of that digit the group
// held in earlier rounds PLUS the lanes of this round whose atomic was serialised before mine. B200 serialises the
// lanes of one ATOMS that hit the same address in ascending lane order (measured: tools/micro/rank3_bench.cu, zero
// violations), so old.count is already the stable rank. That ordering is not architecturally promised, so it is
// VERIFIED, not assumed: old.mask holds the lanes that went first, and a lane that finds a HIGHER lane there raises
// a flag; a warp with a flag redoes its ranking with the order-independent method (read the full mask back after a
// warp sync, rank = count - popc(mask) + popc(mask below me)). Each lane then takes its bit out again with a second
// atomic (no return value). Cost per round of 32 keys: 5.6 SM-cycles against 12.4 for round 1's 64-bit
// {mask,count} words (atomicOr + 64-bit read + leader write-back) and 25 / 60 for ballot / match.any matching.
// Order-independent ranking of one warp's keys (the fallback of the trusted flavour below; never taken on B200 unless
// forced): peers by eight ballots, the running count read by everyone and advanced by the lowest peer. Kept out of line
// -- it re-reads the keys and hands the ranks over through shared memory -- so that it costs the fast path no registers.
template <typename KeyT, bool kInPairs>
__device__ __noinline__ void rank_warp_order_independent(const KeyT* __restrict__ tile_keys, uint32_t valid, uint32_t item0, int items,
                                                         int shift, uint32_t* tbl, uint16_t* s_rank) {
    const uint32_t below = lanemask_lt();
#pragma unroll 1
    for (int i = 0; i < items; ++i) {
        const uint32_t idx = item0 + (uint32_t)i * 32u;
        KeyT key = ~(KeyT)0;
        if (idx < valid) {
            if constexpr (kInPairs) key = __ldg(reinterpret_cast<const uint2*>(tile_keys) + idx).x;   // interleaved {key, value}
            else key = __ldg(tile_keys + idx);
        }
        const uint32_t d = (uint32_t)(key >> shift) & 255u;
        uint32_t peers = 0xFFFFFFFFu;
#pragma unroll
        for (int bit = 0; bit < kRadixBits; ++bit) {
            const bool set = (d >> bit) & 1u;
            const uint32_t vote = __ballot_sync(0xFFFFFFFFu, set);
            peers &= set ? vote : ~vote;
        }
        const uint32_t prior = tbl[d];
        __syncwarp();
        const uint32_t before = __popc(peers & below);
        if (before == 0) tbl[d] = prior + __popc(peers);
        __syncwarp();
        s_rank[idx] = (uint16_t)(prior + before);
    }
    __syncwarp();
}

// Ranking flavour (compile time). 1 (default): whole warps rank with ONE returning atomicAdd per key and trust the
// hardware's ascending-lane serialisation of same-address shared atomics, which every CTA re-validates on its own SM
// before it ranks (a canary: 32-way and 8-way conflicts must come back in lane order), falling back to an
// order-independent ballot ranking for its tile if the canary ever fails. 0: 16-lane groups that additionally carry a
// lane mask in the word, so EVERY atomic's order is verified, at the price of a second atomic per key (clearing the
// lane bit), twice the tables and split loads: 0.334 ms per 2^26-pair pass against 0.27 ms (tools/micro/sort_lab.cu).
#ifndef USRT_RANK_TRUST_LANE_ORDER
#define USRT_RANK_TRUST_LANE_ORDER 1
#endif
constexpr bool kFullWarpRank = USRT_RANK_TRUST_LANE_ORDER != 0;
#ifndef USRT_LOOKBACK_WINDOW
#define USRT_LOOKBACK_WINDOW 4
#endif
constexpr int kLookBack = USRT_LOOKBACK_WINDOW;   // predecessor status words in flight per look-back round trip
#ifndef USRT_TICKET_ITER
#define USRT_TICKET_ITER 8
#endif
constexpr int kTicketIter = USRT_TICKET_ITER;     // output-loop iteration at which a CTA draws its next tile ticket
#ifndef USRT_PREFETCH_AHEAD
#define USRT_PREFETCH_AHEAD 148
#endif
constexpr int kPrefetchAhead = USRT_PREFETCH_AHEAD;   // tiles between a CTA's own tile and the one it asks L2 to fetch (0: off)
constexpr int kGroupsPerWarp = kFullWarpRank ? 1 : 2;

template <typename Cfg, bool kHasValues, int kKeyBytes = 4> struct PassSmem {
    static constexpr int kTblBytes = kGroupsPerWarp * Cfg::kWarps * kRadix * 4;
    static constexpr int kPairBytes = Cfg::kTile * (kKeyBytes + (kHasValues ? 4 : 0));
    static constexpr int kH = Cfg::kBlock / kRadix;                      // threads per digit in the scan step
    static constexpr int kTotal = kTblBytes + kPairBytes + kRadix * 4 + kH * kRadix * 4 + Cfg::kWarps * 4 + 16 + 32;
};

// kPeer: the multi-GPU bucket exchange. Instead of one output array, every digit has its own base address
// (key_ptrs[d] / val_ptrs[d], this rank's slice of the receive buffer of the GPU that owns bucket d, mapped
// through CUDA IPC): the stable scatter of the pass IS the all-to-all, written straight over NVLink.
// flags bit 0: test hook, every warp takes the order-independent ranking path.
// KeyT = uint32_t (the reference's ComputeBufferSorter<uint,uint>) or uint64_t (its GetRadix is generic over uint /
// ulong, ComputeBufferSorter.cs:179-191): 64-bit keys run 8 passes and stage keys and values in separate arrays.
// kIO (32-bit keys with values only): how the pairs cross global memory. 0: separate key / value arrays on both sides
// (the ABI's layout). 1 / 2 / 3: the INTERMEDIATE passes of a 4-pass sort move interleaved {key, value} records -- one
// 64-bit load and one 64-bit store per pair instead of two 32-bit ones, i.e. half the global-memory instructions and
// ~2 fewer L1 wavefronts per 32 pairs: 1 = separate in, interleaved out (pass 0); 2 = interleaved both (passes 1, 2);
// 3 = interleaved in, separate out (pass 3). Interleaved arrays travel through keys_in / keys_out (as uint2*).
template <typename Cfg, typename StatusT, bool kHasValues, bool kPeer = false, typename KeyT = uint32_t, int kIO = 0>
__global__ void __launch_bounds__(Cfg::kBlock, Cfg::kCtasPerSM)
k_onesweep(const KeyT* __restrict__ keys_in, const uint32_t* __restrict__ vals_in, KeyT* __restrict__ keys_out,
           uint32_t* __restrict__ vals_out, uint32_t n, int shift, const uint32_t* __restrict__ digit_base /* [256] */,
           uint32_t* __restrict__ tile_counter, StatusT* __restrict__ status /* [tiles][256], zeroed */,
           StatusT* __restrict__ next_status /* the next pass's [tiles][256], zeroed HERE (may be NULL) */, uint32_t flags,
           const unsigned long long* __restrict__ key_ptrs = nullptr, const unsigned long long* __restrict__ val_ptrs = nullptr) {
    using ST = StatusTraits<StatusT>;
    using SM = PassSmem<Cfg, kHasValues, (int)sizeof(KeyT)>;
    constexpr bool kWide = sizeof(KeyT) == 8;
    static_assert(!(kWide && kPeer), "the multi-GPU bucket exchange is built for 32-bit keys");
    constexpr bool kInPairs = kIO == 2 || kIO == 3, kOutPairs = kIO == 1 || kIO == 2;
    static_assert(kIO == 0 || (kHasValues && !kWide && !kPeer), "interleaved records: 32-bit keys with values");
    constexpr int kBlock = Cfg::kBlock, kIPT = Cfg::kIPT, kTile = Cfg::kTile, kWarps = Cfg::kWarps;
    constexpr int kGroups = kGroupsPerWarp * kWarps, kH = SM::kH, kGP = kGroups / kH;      // groups per scan thread
    constexpr int kLanes = 32 / kGroupsPerWarp;                                 // lanes per ranking group
    constexpr int kPack = kFullWarpRank ? 3 : 4, kRankBits = 32 / kPack;        // ranks per register (10 / 8 bits each)
    static_assert(kLanes * kIPT <= (1 << kRankBits), "packed ranks must fit their field");
    static_assert(kGroups % kH == 0, "scan threads split the groups evenly");
    extern __shared__ __align__(16) unsigned char smem[];
    uint32_t* s_tbl = reinterpret_cast<uint32_t*>(smem);                         // [kGroups][256]
    uint2* s_pairs = reinterpret_cast<uint2*>(smem + SM::kTblBytes);             // 32-bit keys with values: {key, value}
    KeyT* s_keys = reinterpret_cast<KeyT*>(smem + SM::kTblBytes);                // otherwise: keys[tile] | values[tile]
    uint32_t* s_vals = reinterpret_cast<uint32_t*>(smem + SM::kTblBytes + kTile * (int)sizeof(KeyT));
    uint32_t* s_global_off = reinterpret_cast<uint32_t*>(smem + SM::kTblBytes + SM::kPairBytes);
    uint32_t* s_part = s_global_off + kRadix;                                    // [kH][256]
    uint32_t* s_scan = s_part + kH * kRadix;                                     // [kWarps]
    uint32_t* s_tile_id = s_scan + kWarps;
    uint32_t* s_canary = s_tile_id + 4;                                          // [8]: lane-order self-test (see below)

    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;

    // Persistent CTAs: each draws tile tickets from a counter until none is left. A tile only ever waits on tiles with
    // lower tickets, which running CTAs hold => forward progress. The NEXT ticket is drawn in the middle of the current
    // tile's output loop, so the counter's L2 round trip no longer sits at the head of a tile with nothing else of the CTA
    // to run (0.266 -> 0.254 ms per 2^26-pair pass). Not earlier: tickets fix the order in which tiles expect each other's
    // aggregates, and a ticket that is held for long before its tile starts makes every successor's look-back wait for it
    // (drawn before the look-back: 0.272 ms, with the next tile's keys requested ahead as well: 0.297 ms).
    const uint32_t total_tiles = (n + (uint32_t)kTile - 1u) / (uint32_t)kTile;
    if (kFullWarpRank && warp == 0) {
        // Canary: the fast ranking below relies on one property of this SM's shared-memory atomic unit -- the lanes of
        // ONE instruction that hit the same address are applied in ascending lane order. Test it here, on this SM,
        // with a 32-way and four 8-way conflicts; if it ever fails this CTA ranks its tile the order-independent way.
        // The test itself must run as ONE instruction per atomic, i.e. with the warp converged: nothing lane-dependent
        // may precede it (every lane clears a word; the ticket draw of thread 0 comes afterwards) -- a warp that reaches
        // the atomics in two groups reports the second group out of order although the hardware did nothing wrong.
        s_canary[lane & 7u] = 0u;
        __syncwarp();
        const uint32_t a = atomicAdd(s_canary + 0, 1u);
        const uint32_t b = atomicAdd(s_canary + 1 + (lane & 3u), 1u);
        __syncwarp();
        if (a != lane || b != (lane >> 2) || (flags & 1u)) s_canary[7] = 1u;
    }
    if (tid == 0) s_tile_id[0] = atomicAdd(tile_counter, 1u);
    // Item map. A ranking group (a whole warp, or a 16-lane half in the fully verified flavour) owns kLanes * IPT
    // CONSECUTIVE keys of the tile, item i of its lane l being key group * kLanes * IPT + kLanes * i + l -- so (round,
    // lane) order inside a group IS memory order, which is what makes the per-group ranks stable. With whole warps a
    // load is one coalesced 128-byte line; with half-warp groups it touches two 64-byte segments.
    const uint32_t gl = lane & (uint32_t)(kLanes - 1);                      // lane within its group
    const uint32_t group = (uint32_t)kGroupsPerWarp * warp + (kGroupsPerWarp == 2 ? (lane >> 4) : 0u);
    const uint32_t item0 = group * (uint32_t)(kLanes * kIPT) + gl;
#pragma unroll 1
    for (uint32_t it = 0;; ++it) {
        {
            uint4* z = reinterpret_cast<uint4*>(smem);
#pragma unroll
            for (int i = 0; i < SM::kTblBytes / 16 / kBlock; ++i) z[tid + i * kBlock] = make_uint4(0, 0, 0, 0);
        }
        __syncthreads();                                            // also: every warp has left the previous tile's output loop
        const uint32_t tile = s_tile_id[it & 1u];
        if (tile >= total_tiles) break;
        const uint32_t tile_base = tile * (uint32_t)kTile;
        const uint32_t valid = min((uint32_t)kTile, n - tile_base);
        // every tile clears its row of the NEXT pass's look-back words (that pass starts after this kernel has finished):
        // one 1 KB store per tile instead of a memset over all passes' status before the sort
        if (next_status != nullptr && tid < kRadix) next_status[(size_t)tile * kRadix + tid] = 0;
        if (kPrefetchAhead > 0) {
            // Ask L2 for a tile a little further down the input (tiles are handed out in order, so some CTA will want it
            // soon): its own loads then hit L2 instead of waiting on DRAM at the start of a CTA, where nothing else of that
            // CTA can run. Any distance from 64 to 300 tiles measures the same: 0.277 -> 0.264 ms per 2^26-pair pass.
            const uint64_t ahead = (uint64_t)tile_base + (uint64_t)kPrefetchAhead * kTile;
            if (ahead + kTile <= n) {
                if constexpr (kInPairs) {
                    const char* p = reinterpret_cast<const char*>(reinterpret_cast<const uint2*>(keys_in) + ahead);
                    for (uint32_t b = tid * 128u; b < (uint32_t)kTile * 8u; b += kBlock * 128u) asm volatile("prefetch.global.L2 [%0];" ::"l"(p + b));
                } else {
                    const char* p = reinterpret_cast<const char*>(keys_in + ahead);
                    for (uint32_t b = tid * 128u; b < (uint32_t)kTile * (uint32_t)sizeof(KeyT); b += kBlock * 128u) asm volatile("prefetch.global.L2 [%0];" ::"l"(p + b));
                    if (kHasValues) {
                        const char* q = reinterpret_cast<const char*>(vals_in + ahead);
                        for (uint32_t b = tid * 128u; b < (uint32_t)kTile * 4u; b += kBlock * 128u) asm volatile("prefetch.global.L2 [%0];" ::"l"(q + b));
                    }
                }
            }
        }

        KeyT key[kIPT];
        uint32_t val[kHasValues ? kIPT : 1];
#pragma unroll
        for (int i = 0; i < kIPT; ++i) {
            const uint32_t idx = item0 + (uint32_t)i * kLanes;
            if constexpr (kInPairs) {
                const uint2 kv = idx < valid ? __ldg(reinterpret_cast<const uint2*>(keys_in) + tile_base + idx) : make_uint2(0xFFFFFFFFu, 0u);
                key[i] = kv.x; val[i] = kv.y;
            } else {
                key[i] = idx < valid ? __ldg(keys_in + tile_base + idx) : ~(KeyT)0;  // tail pads sort last, never stored
            }
        }

        // stable rank of every key among the keys of its group with the same digit, kPack to a register
        uint32_t rankp[(kIPT + kPack - 1) / kPack];
        uint32_t* tbl = s_tbl + group * kRadix;
        auto put_rank = [&](int i, uint32_t r) {
            if (i % kPack == 0) rankp[i / kPack] = r; else rankp[i / kPack] |= r << ((i % kPack) * kRankBits);
        };
        if constexpr (kFullWarpRank) {
            if (s_canary[7] == 0u) {
                // One returning atomic per key: old = keys of this digit in the warp's earlier rounds + the lower lanes of this
                // round (ascending-lane serialisation, validated by the canary above) = the stable rank.
#pragma unroll
                for (int i = 0; i < kIPT; ++i) {
                    put_rank(i, atomicAdd(tbl + ((uint32_t)(key[i] >> shift) & 255u), 1u));
                    __syncwarp();                                   // rounds are ordered
                }
            } else {
                uint16_t* s_rank = reinterpret_cast<uint16_t*>(smem + SM::kTblBytes);       // the staging area is still free
                const KeyT* tile_keys = kInPairs ? reinterpret_cast<const KeyT*>(reinterpret_cast<const uint2*>(keys_in) + tile_base)
                                                 : keys_in + tile_base;
                rank_warp_order_independent<KeyT, kInPairs>(tile_keys, valid, item0, kIPT, shift, tbl, s_rank);
#pragma unroll
                for (int i = 0; i < kIPT; ++i) put_rank(i, s_rank[item0 + (uint32_t)i * kLanes]);
            }
        } else {
            // Fully verified flavour: the word also carries the lane bits of the round, so a lane sees which lanes of its group
            // were applied before it; a higher lane among them sends the warp to the order-independent method.
            const uint32_t add = (0x10000u << gl) | 1u, bit = 0x10000u << gl;
            const uint32_t not_below = 0xFFFFu << gl;              // my own lane and the higher ones of my group
            uint32_t out_of_order = flags & 1u;
#pragma unroll
            for (int i = 0; i < kIPT; ++i) {
                const uint32_t d = (uint32_t)(key[i] >> shift) & 255u;
                const uint32_t old = atomicAdd(tbl + d, add);
                __syncwarp();
                atomicSub(tbl + d, bit);
                __syncwarp();
                out_of_order |= (old >> 16) & not_below;
                put_rank(i, old & 0xFFFFu);
            }
            if (__any_sync(0xFFFFFFFFu, out_of_order != 0u)) {
                // Order-independent ranking (never taken on B200 unless forced): start the warp's two tables over.
                uint32_t* mine = s_tbl + 2u * warp * kRadix;
#pragma unroll
                for (int i = 0; i < 2 * kRadix / 32; ++i) mine[lane + 32 * i] = 0u;
                __syncwarp();
                const uint32_t below = (1u << gl) - 1u;
#pragma unroll
                for (int i = 0; i < kIPT; ++i) {
                    const uint32_t d = (uint32_t)(key[i] >> shift) & 255u;
                    atomicAdd(tbl + d, add);
                    __syncwarp();
                    const uint32_t now = tbl[d];                    // every peer of this round has added itself
                    __syncwarp();
                    const uint32_t peers = now >> 16, total = now & 0xFFFFu;
                    const uint32_t before = __popc(peers & below);
                    if (before == 0) tbl[d] = total;                // the lowest peer clears the round's lane bits
                    __syncwarp();
                    put_rank(i, total - __popc(peers) + before);
                }
            }
        }
        __syncthreads();

        // values are fetched now, so their latency hides behind the scan and the look-back below
        if (kHasValues && !kInPairs) {
#pragma unroll
            for (int i = 0; i < kIPT; ++i) {
                const uint32_t idx = item0 + (uint32_t)i * kLanes;
                val[i] = idx < valid ? __ldg(vals_in + tile_base + idx) : 0u;
            }
        }

        // Scan step: kH threads per digit, each owning kGP consecutive groups. Thread (d, h) sums its groups' counts of
        // digit d, the kH partial sums meet in s_part, every thread then knows the tile's count of d (published for the
        // look-back right away) and the counts of the groups before its own.
        const uint32_t d = tid & 255u, h = tid >> 8;
        uint32_t cnt[kGP];
        uint32_t part = 0;
#pragma unroll
        for (int j = 0; j < kGP; ++j) { cnt[j] = s_tbl[(h * kGP + j) * kRadix + d]; part += cnt[j]; }
        if (kH > 1) {
            s_part[h * kRadix + d] = part;
            __syncthreads();
        }
        uint32_t count = 0, groups_before = 0;
        if (kH > 1) {
#pragma unroll
            for (int k = 0; k < kH; ++k) {
                const uint32_t v = s_part[k * kRadix + d];
                count += v;
                groups_before += (k < (int)h) ? v : 0u;
            }
        } else {
            count = part;
        }
        StatusT* my_status = status + (size_t)tile * kRadix + d;
        if (h == 0) ST::store(my_status, (tile == 0 ? ST::kPrefix : ST::kAggregate) | (StatusT)count);
        // exclusive scan of the 256 digit counts (every h does it for itself: no extra barrier, the values are the same)
        uint32_t incl = count;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if (lane >= (uint32_t)o) incl += y;
        }
        if (lane == 31) s_scan[warp] = incl;
        __syncthreads();
        uint32_t tile_start;                                        // first tile-local slot of digit d
        {
            const uint32_t w0 = h * (kRadix / 32);                  // first warp of my h
            uint32_t wbase = 0;
#pragma unroll
            for (int w = 0; w < kRadix / 32; ++w) wbase += (w0 + w < warp) ? s_scan[w0 + w] : 0u;
            tile_start = wbase + incl - count;
            // in place: the count of (group, d) becomes the first tile slot of that group's keys of digit d
            uint32_t running = tile_start + groups_before;
#pragma unroll
            for (int j = 0; j < kGP; ++j) { s_tbl[(h * kGP + j) * kRadix + d] = running; running += cnt[j]; }
        }
        __syncthreads();

        // stage the tile in digit order
#pragma unroll
        for (int i = 0; i < kIPT; ++i) {
            const uint32_t dg = (uint32_t)(key[i] >> shift) & 255u;
            const uint32_t slot = tbl[dg] + ((rankp[i / kPack] >> ((i % kPack) * kRankBits)) & ((1u << kRankBits) - 1u));
            if constexpr (kHasValues && !kWide) {
                s_pairs[slot] = make_uint2((uint32_t)key[i], val[i]);
            } else {
                s_keys[slot] = key[i];
                if (kHasValues) s_vals[slot] = val[i];
            }
        }
        // Look-back AFTER staging: the aggregate was published before the scan, so by now the preceding
        // tiles have usually posted their inclusive prefixes and the walk resolves in one round trip.
        if (h == 0) {
            // decoupled look-back over the preceding tiles' counts of this digit, kLookBack tiles per round
            // trip (the loads are independent; only the accumulation is ordered)
            uint32_t exclusive = 0;
            if (tile > 0) {
                int32_t t = (int32_t)tile - 1;
                bool done = false;
                while (!done) {
                    StatusT sw[kLookBack];
#pragma unroll
                    for (int k = 0; k < kLookBack; ++k)
                        sw[k] = (t - k >= 0) ? ST::load(status + (size_t)(t - k) * kRadix + d) : (StatusT)ST::kPrefix;
#pragma unroll
                    for (int k = 0; k < kLookBack; ++k) {
                        if (!done) {
                            while ((sw[k] & ST::kFlagMask) == 0) sw[k] = ST::load(status + (size_t)(t - k) * kRadix + d);
                            exclusive += (uint32_t)(sw[k] & ST::kValueMask);
                            done = (sw[k] & ST::kPrefix) != 0;
                        }
                    }
                    t -= kLookBack;
                }
                ST::store(my_status, ST::kPrefix | (StatusT)(exclusive + count));
            }
            s_global_off[d] = (kPeer ? 0u : digit_base[d]) + exclusive - tile_start;   // wraps mod 2^32 by design
        }
        __syncthreads();

        // coalesced runs out: slot p of the tile goes to global_off[digit] + p
#pragma unroll
        for (int i = 0; i < kIPT; ++i) {
            if (i == (kTicketIter < kIPT ? kTicketIter : 0) && tid == 0)
                s_tile_id[(it + 1u) & 1u] = atomicAdd(tile_counter, 1u);            // next ticket; read after the loop-top barrier
            const uint32_t p = tid + (uint32_t)i * kBlock;
            if (p < valid) {
                if constexpr (kHasValues && !kWide) {
                    const uint2 kv = s_pairs[p];
                    const uint32_t dg = (kv.x >> shift) & 255u;
                    const uint32_t dst = s_global_off[dg] + p;
                    if (kPeer) {
                        const unsigned long long kp = __ldg(key_ptrs + dg), vp = __ldg(val_ptrs + dg);
                        if (kp != 0ull) {                          // null plan = a receive buffer would overflow: write nothing
                            reinterpret_cast<uint32_t*>(kp)[dst] = kv.x;
                            reinterpret_cast<uint32_t*>(vp)[dst] = kv.y;
                        }
                    } else if (kOutPairs) {
                        reinterpret_cast<uint2*>(keys_out)[dst] = kv;
                    } else {
                        keys_out[dst] = kv.x;
                        vals_out[dst] = kv.y;
                    }
                } else {
                    const KeyT k = s_keys[p];
                    const uint32_t dst = s_global_off[(uint32_t)(k >> shift) & 255u] + p;
                    keys_out[dst] = k;
                    if (kHasValues) vals_out[dst] = s_vals[p];
                }
            }
        }
    }   // next tile
}

// ---- multi-GPU bucket exchange: the landing plan, computed on the device --------------------------------------------
// One 256-thread block turns the all-gathered top-byte histograms (world x 256 counts) into (a) contiguous bucket
// ranges of ~equal mass, one per rank, (b) for THIS rank as a source, the element offset of each of its 256 runs
// inside its owner's receive buffer (source-rank-major, digits ascending within a source -- the order an all-to-all
// would produce, so the result stays globally stable), turned into absolute key / value addresses, and (c) how many
// pairs every rank receives. Every rank runs the same integer arithmetic on the same input, so the plans agree.
// It replaces a host round trip (histograms to the CPU, numpy plan, pointer table back) in the middle of the sort.
constexpr int kMaxPeerWorld = 16;
__global__ void __launch_bounds__(kRadix) k_peer_scatter_plan(const uint32_t* __restrict__ all_hist, int world, int rank,
                                                              const unsigned long long* __restrict__ peer_base,
                                                              unsigned long long capacity,
                                                              unsigned long long* __restrict__ key_ptrs,
                                                              unsigned long long* __restrict__ val_ptrs,
                                                              unsigned long long* __restrict__ recv_total /* [world] */,
                                                              uint32_t* __restrict__ bounds_out /* [world+1] */) {
    __shared__ unsigned long long s_csum[kRadix + 1];                   // global exclusive prefix over digits
    __shared__ unsigned long long s_src[kMaxPeerWorld][kRadix + 1];     // per-source exclusive prefix over digits
    __shared__ unsigned long long s_warp[kRadix / 32];
    __shared__ uint32_t s_bounds[kMaxPeerWorld + 1];
    const uint32_t d = threadIdx.x, lane = d & 31u, warp = d >> 5;
    auto block_exclusive = [&](unsigned long long v, unsigned long long* out /* [257] */) {
        unsigned long long incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long y = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if (lane >= (uint32_t)o) incl += y;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        unsigned long long base = 0;
        for (uint32_t w = 0; w < warp; ++w) base += s_warp[w];
        out[d + 1] = base + incl;
        if (d == 0) out[0] = 0;
        __syncthreads();
    };
    unsigned long long g = 0;
    for (int s = 0; s < world; ++s) g += all_hist[s * kRadix + d];
    block_exclusive(g, s_csum);
    for (int s = 0; s < world; ++s) block_exclusive(all_hist[s * kRadix + d], s_src[s]);
    if (d == 0) {
        // bucket ranges: the boundary for rank r is the digit boundary whose prefix is closest to total * r / world
        // (ties towards the lower one), kept non-decreasing; compared as exact integers (prefix * world vs total * r)
        const unsigned long long total = s_csum[kRadix];
        s_bounds[0] = 0;
        for (int r = 1; r < world; ++r) {
            const unsigned long long target = total * (unsigned long long)r;
            uint32_t b = 0;
            while (b < (uint32_t)kRadix && s_csum[b] * (unsigned long long)world < target) ++b;
            if (b > 0) {
                const unsigned long long hi = s_csum[b] * (unsigned long long)world, lo = s_csum[b - 1] * (unsigned long long)world;
                const unsigned long long dist_hi = hi >= target ? hi - target : target - hi, dist_lo = target - lo;
                if (dist_lo <= dist_hi) --b;
            }
            s_bounds[r] = min(max(b, s_bounds[r - 1]), (uint32_t)kRadix);
        }
        s_bounds[world] = kRadix;
    }
    __syncthreads();
    int owner = 0;
    while (owner + 1 < world && d >= s_bounds[owner + 1]) ++owner;
    const uint32_t ob = s_bounds[owner], oe = s_bounds[owner + 1];
    unsigned long long before = 0;                                      // pairs the lower-ranked sources send to my owner
    for (int s = 0; s < rank; ++s) before += s_src[s][oe] - s_src[s][ob];
    const unsigned long long offset = before + s_src[rank][d] - s_src[rank][ob];
    unsigned long long t = 0;
    if (d < (uint32_t)world) {
        for (int s = 0; s < world; ++s) t += s_src[s][s_bounds[d + 1]] - s_src[s][s_bounds[d]];
        recv_total[d] = t;
    }
    // a receive buffer that is too small: null addresses make the scatter pass write nothing (the host sees the
    // counts and reports the error) instead of running past a peer's allocation
    const bool overflow = __syncthreads_or(t > capacity) != 0;
    key_ptrs[d] = overflow ? 0ull : peer_base[owner] + 4ull * offset;
    val_ptrs[d] = overflow ? 0ull : peer_base[owner] + 4ull * (capacity + offset);
    if (d <= (uint32_t)world && bounds_out != nullptr) bounds_out[d] = s_bounds[d];
}

// SM count of the current device (148 on B200); persistent grids are sized from it
inline int sm_count() {
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) {
        cudaGetLastError();
        return kNumSMs;
    }
    return sms;
}
inline uint32_t histogram_grid(uint64_t count) {
    cudaFuncSetAttribute(k_histogram, cudaFuncAttributeMaxDynamicSharedMemorySize, kHistSmemBytes);   // per device, cheap
    const uint64_t vec_work = (count + 4 * kHistThreads - 1) / (4 * kHistThreads);
    return (uint32_t)std::min<uint64_t>(std::max<uint64_t>(vec_work, 1), (uint64_t)sm_count());
}
inline uint32_t tile_pairs(uint64_t count) { return count < kSmallSortLimit ? SmallTile::kTile : BigTile::kTile; }
inline uint32_t num_tiles(uint64_t count) { return (uint32_t)((count + tile_pairs(count) - 1) / tile_pairs(count)); }
inline bool wide_status(uint64_t count) {
    static const bool forced = getenv("USRT_FORCE_WIDE_STATUS") != nullptr;   // test hook: 64-bit look-back words at any size
    return forced || count >= (1ull << 30);
}
inline uint32_t pass_flags() {
    static const uint32_t f = getenv("USRT_FORCE_SLOW_RANK") != nullptr ? 1u : 0u;   // test hook: order-independent ranking everywhere
    return f;
}
// persistent CTAs: as many as fit on the GPU at once (or one per tile when there are fewer tiles)
template <typename Cfg> inline uint32_t pass_grid(uint32_t tiles) { return std::min<uint32_t>(tiles, (uint32_t)(sm_count() * Cfg::kCtasPerSM)); }
inline uint64_t status_words_bytes(uint64_t count) { return (uint64_t)num_tiles(count) * kRadix * (wide_status(count) ? 8 : 4); }

template <typename Cfg, typename StatusT, bool kHasValues, int kIO = 0>
cudaError_t launch_pass(const uint32_t* ki, const uint32_t* vi, uint32_t* ko, uint32_t* vo, uint64_t count, int shift,
                        const uint32_t* digit_base, uint32_t* tile_counter, void* status, void* next_status, cudaStream_t stream) {
    constexpr int smem = PassSmem<Cfg, kHasValues>::kTotal;
    auto kern = k_onesweep<Cfg, StatusT, kHasValues, false, uint32_t, kIO>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    kern<<<pass_grid<Cfg>(num_tiles(count)), Cfg::kBlock, smem, stream>>>(ki, vi, ko, vo, (uint32_t)count, shift, digit_base, tile_counter,
                                                         static_cast<StatusT*>(status), static_cast<StatusT*>(next_status), pass_flags(), nullptr, nullptr);
    return cudaGetLastError();
}

// io: 0 separate arrays; 1 / 2 / 3 interleaved records out / both / in (big tiles with values only)
template <typename StatusT>
cudaError_t run_pass(const uint32_t* ki, const uint32_t* vi, uint32_t* ko, uint32_t* vo, uint64_t count, int shift,
                     const uint32_t* digit_base, uint32_t* tile_counter, void* status, void* next_status, cudaStream_t stream, int io = 0) {
    const bool small = count < kSmallSortLimit;
    if (io == 1) return launch_pass<BigTile, StatusT, true, 1>(ki, vi, ko, vo, count, shift, digit_base, tile_counter, status, next_status, stream);
    if (io == 2) return launch_pass<BigTile, StatusT, true, 2>(ki, vi, ko, vo, count, shift, digit_base, tile_counter, status, next_status, stream);
    if (io == 3) return launch_pass<BigTile, StatusT, true, 3>(ki, vi, ko, vo, count, shift, digit_base, tile_counter, status, next_status, stream);
    if (vi != nullptr)
        return small ? launch_pass<SmallTile, StatusT, true>(ki, vi, ko, vo, count, shift, digit_base, tile_counter, status, next_status, stream)
                     : launch_pass<BigTile, StatusT, true>(ki, vi, ko, vo, count, shift, digit_base, tile_counter, status, next_status, stream);
    return small ? launch_pass<SmallTile, StatusT, false>(ki, vi, ko, vo, count, shift, digit_base, tile_counter, status, next_status, stream)
                 : launch_pass<BigTile, StatusT, false>(ki, vi, ko, vo, count, shift, digit_base, tile_counter, status, next_status, stream);
}

}  // namespace

cudaError_t sort_scratch_reserve(SortScratch& s, uint64_t count, bool need_alt) {
    cudaError_t e;
    if (s.hist == nullptr) {
        if ((e = cudaMalloc(&s.hist, 2 * kSortPasses * kRadix * sizeof(uint32_t))) != cudaSuccess) return e;   // 8 digits: 64-bit keys
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
    // two interleaved {key, value} buffers for the intermediate passes of large sorts (best effort: a sort that cannot
    // get them runs all four passes on separate arrays)
    if (count >= kPairsSortLimit && count > s.pairs_capacity && getenv("USRT_SORT_NO_PAIRS") == nullptr) {
        if (s.pairs_x) cudaFree(s.pairs_x);
        if (s.pairs_y) cudaFree(s.pairs_y);
        s.pairs_x = s.pairs_y = nullptr; s.pairs_capacity = 0;
        if (cudaMalloc(&s.pairs_x, count * 8) == cudaSuccess && cudaMalloc(&s.pairs_y, count * 8) == cudaSuccess) {
            s.pairs_capacity = count;
        } else {
            cudaGetLastError();
            if (s.pairs_x) cudaFree(s.pairs_x);
            s.pairs_x = s.pairs_y = nullptr;
        }
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
    if (s.keys64_alt) cudaFree(s.keys64_alt);
    if (s.vals64_alt) cudaFree(s.vals64_alt);
    if (s.pairs_x) cudaFree(s.pairs_x);
    if (s.pairs_y) cudaFree(s.pairs_y);
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
    // tile counters + pass 0's look-back words; every pass clears the words of the pass that follows it
    if ((e = cudaMemsetAsync(s.status, 0, kHeaderWords * 4 + pass_bytes, stream)) != cudaSuccess) return e;

    k_histogram<<<histogram_grid(count), kHistThreads, kHistSmemBytes, stream>>>(keys, count, s.hist);
    k_scan_histogram<<<1, kSortPasses * kRadix, 0, stream>>>(s.hist, nullptr, -1);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    if (launches) *launches += 2;
    if (events && (e = cudaEventRecord(events[1], stream)) != cudaSuccess) return e;

    uint32_t* counters = static_cast<uint32_t*>(s.status);
    char* status0 = static_cast<char*>(s.status) + kHeaderWords * 4;
    const uint32_t* ki = keys; const uint32_t* vi = vals;
    uint32_t* ko = keys_alt; uint32_t* vo = vals_alt;
    // large key/value sorts: (keys, vals) -> X -> Y -> X -> (keys, vals) with X, Y interleaved {key, value} records
    const bool pairs = vals != nullptr && count >= kPairsSortLimit && s.pairs_capacity >= count && count >= kSmallSortLimit;
    for (int pass = 0; pass < kSortPasses; ++pass) {               // bitOffset = 0, 8, 16, 24
        void* st = status0 + (uint64_t)pass * pass_bytes;
        void* next = pass + 1 < kSortPasses ? status0 + (uint64_t)(pass + 1) * pass_bytes : nullptr;
        int io = 0;
        if (pairs) {
            uint32_t* x = reinterpret_cast<uint32_t*>(s.pairs_x); uint32_t* y = reinterpret_cast<uint32_t*>(s.pairs_y);
            io = pass == 0 ? 1 : (pass == kSortPasses - 1 ? 3 : 2);
            ki = pass == 0 ? keys : (pass == 2 ? y : x);           // passes 1 and 3 read X, pass 2 reads Y
            vi = vals;
            ko = pass == kSortPasses - 1 ? keys : (pass == 1 ? y : x);
            vo = vals;
        }
        if (wide_status(count))
            e = run_pass<uint64_t>(ki, vi, ko, vo, count, pass * kRadixBits, s.hist + pass * kRadix, counters + pass, st, next, stream, io);
        else
            e = run_pass<uint32_t>(ki, vi, ko, vo, count, pass * kRadixBits, s.hist + pass * kRadix, counters + pass, st, next, stream, io);
        if (e != cudaSuccess) return e;
        if (launches) *launches += 1;
        if (events && (e = cudaEventRecord(events[2 + pass], stream)) != cudaSuccess) return e;
        if (!pairs) {
            const uint32_t* tk = ki; const uint32_t* tv = vi;
            ki = ko; vi = vo;
            ko = const_cast<uint32_t*>(tk); vo = const_cast<uint32_t*>(tv);
        }
    }
    return cudaSuccess;   // even number of passes: the result is back in (keys, vals)
}

// ---- ComputeBufferSorter<ulong, uint>: 8 passes x 8 bits over 64-bit keys (ComputeBufferSorter.cs:179-191) -----------
namespace {
using Tile64 = TileCfg<512, 8, 2>;                   // 4096 pairs: 64-bit keys take two registers each
inline uint32_t num_tiles64(uint64_t count) { return (uint32_t)((count + Tile64::kTile - 1) / Tile64::kTile); }
inline uint64_t status_bytes64(uint64_t count) { return (uint64_t)num_tiles64(count) * kRadix * (wide_status(count) ? 8 : 4); }

template <typename StatusT, bool kHasValues>
cudaError_t launch_pass64(const uint64_t* ki, const uint32_t* vi, uint64_t* ko, uint32_t* vo, uint64_t count, int shift,
                          const uint32_t* digit_base, uint32_t* tile_counter, void* status, void* next_status, cudaStream_t stream) {
    constexpr int smem = PassSmem<Tile64, kHasValues, 8>::kTotal;
    auto kern = k_onesweep<Tile64, StatusT, kHasValues, false, uint64_t>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    kern<<<pass_grid<Tile64>(num_tiles64(count)), Tile64::kBlock, smem, stream>>>(ki, vi, ko, vo, (uint32_t)count, shift, digit_base, tile_counter,
                                                               static_cast<StatusT*>(status), static_cast<StatusT*>(next_status), pass_flags(), nullptr, nullptr);
    return cudaGetLastError();
}
}  // namespace

cudaError_t sort_scratch_reserve64(SortScratch& s, uint64_t count, bool need_alt) {
    cudaError_t e;
    if ((e = sort_scratch_reserve(s, 1, false)) != cudaSuccess) return e;          // hist
    const uint64_t need = kHeaderWords * 4 + 2 * kSortPasses * status_bytes64(count);
    if (need > s.status_bytes) {
        if (s.status) cudaFree(s.status);
        s.status = nullptr; s.status_bytes = 0;
        if ((e = cudaMalloc(&s.status, need)) != cudaSuccess) return e;
        s.status_bytes = need;
        ++s.generation;
    }
    if (need_alt && count > s.alt64_capacity) {
        if (s.keys64_alt) cudaFree(s.keys64_alt);
        if (s.vals64_alt) cudaFree(s.vals64_alt);
        s.keys64_alt = nullptr; s.vals64_alt = nullptr; s.alt64_capacity = 0;
        if ((e = cudaMalloc(&s.keys64_alt, count * 8)) != cudaSuccess) return e;
        if ((e = cudaMalloc(&s.vals64_alt, count * 4)) != cudaSuccess) return e;
        s.alt64_capacity = count;
    }
    return cudaSuccess;
}

cudaError_t sort_pairs64(uint64_t* keys, uint32_t* vals, uint64_t* keys_alt, uint32_t* vals_alt, uint64_t count, SortScratch& s,
                         cudaStream_t stream, uint64_t* launches) {
    if (count == 0) return cudaSuccess;
    cudaError_t e;
    if ((e = sort_scratch_reserve64(s, count, false)) != cudaSuccess) return e;
    constexpr int kPasses = 2 * kSortPasses;
    const uint64_t pass_bytes = status_bytes64(count);
    if ((e = cudaMemsetAsync(s.hist, 0, kPasses * kRadix * 4, stream)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(s.status, 0, kHeaderWords * 4 + pass_bytes, stream)) != cudaSuccess) return e;
    cudaFuncSetAttribute(k_histogram64, cudaFuncAttributeMaxDynamicSharedMemorySize, kHistSmemBytes);
    const uint32_t grid = (uint32_t)std::min<uint64_t>(std::max<uint64_t>((count + kHistThreads - 1) / kHistThreads, 1), (uint64_t)sm_count());
    for (int half = 0; half < 2; ++half) {
        k_histogram64<<<grid, kHistThreads, kHistSmemBytes, stream>>>(reinterpret_cast<const uint2*>(keys), count, half,
                                                                       s.hist + half * kSortPasses * kRadix);
        k_scan_histogram<<<1, kSortPasses * kRadix, 0, stream>>>(s.hist + half * kSortPasses * kRadix, nullptr, -1);
    }
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    if (launches) *launches += 4;
    uint32_t* counters = static_cast<uint32_t*>(s.status);
    char* status0 = static_cast<char*>(s.status) + kHeaderWords * 4;
    const uint64_t* ki = keys; const uint32_t* vi = vals;
    uint64_t* ko = keys_alt; uint32_t* vo = vals_alt;
    for (int pass = 0; pass < kPasses; ++pass) {                    // bitOffset = 0, 8, ..., 56
        void* st = status0 + (uint64_t)pass * pass_bytes;
        void* next = pass + 1 < kPasses ? status0 + (uint64_t)(pass + 1) * pass_bytes : nullptr;
        const bool wide = wide_status(count);
        if (vi != nullptr)
            e = wide ? launch_pass64<uint64_t, true>(ki, vi, ko, vo, count, pass * kRadixBits, s.hist + pass * kRadix, counters + pass, st, next, stream)
                     : launch_pass64<uint32_t, true>(ki, vi, ko, vo, count, pass * kRadixBits, s.hist + pass * kRadix, counters + pass, st, next, stream);
        else
            e = wide ? launch_pass64<uint64_t, false>(ki, vi, ko, vo, count, pass * kRadixBits, s.hist + pass * kRadix, counters + pass, st, next, stream)
                     : launch_pass64<uint32_t, false>(ki, vi, ko, vo, count, pass * kRadixBits, s.hist + pass * kRadix, counters + pass, st, next, stream);
        if (e != cudaSuccess) return e;
        if (launches) *launches += 1;
        const uint64_t* tk = ki; const uint32_t* tv = vi;
        ki = ko; vi = vo;
        ko = const_cast<uint64_t*>(tk); vo = const_cast<uint32_t*>(tv);
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
        e = run_pass<uint64_t>(src_keys, src_vals, dst_keys, dst_vals, count, bit_offset, s.hist + pass * kRadix, counters, st, nullptr, stream);
    else
        e = run_pass<uint32_t>(src_keys, src_vals, dst_keys, dst_vals, count, bit_offset, s.hist + pass * kRadix, counters, st, nullptr, stream);
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
        k_onesweep<SmallTile, uint32_t, true, true><<<pass_grid<SmallTile>(num_tiles(count)), SmallTile::kBlock, smem, stream>>>(
            src_keys, src_vals, static_cast<uint32_t*>(nullptr), nullptr, (uint32_t)count, bit_offset, nullptr, counters, st, static_cast<uint32_t*>(nullptr), pass_flags(), key_ptrs, val_ptrs);
    } else {
        constexpr int smem = PassSmem<BigTile, true>::kTotal;
        cudaFuncSetAttribute(k_onesweep<BigTile, uint32_t, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        k_onesweep<BigTile, uint32_t, true, true><<<pass_grid<BigTile>(num_tiles(count)), BigTile::kBlock, smem, stream>>>(
            src_keys, src_vals, static_cast<uint32_t*>(nullptr), nullptr, (uint32_t)count, bit_offset, nullptr, counters, st, static_cast<uint32_t*>(nullptr), pass_flags(), key_ptrs, val_ptrs);
    }
    if (launches) *launches += 1;
    return cudaGetLastError();
}

cudaError_t peer_scatter_plan(const uint32_t* all_hist, int world, int rank, const unsigned long long* peer_base,
                              unsigned long long capacity, unsigned long long* key_ptrs, unsigned long long* val_ptrs,
                              unsigned long long* recv_total, uint32_t* bounds_out, cudaStream_t stream, uint64_t* launches) {
    if (world < 1 || world > kMaxPeerWorld || rank < 0 || rank >= world) return cudaErrorInvalidValue;
    k_peer_scatter_plan<<<1, kRadix, 0, stream>>>(all_hist, world, rank, peer_base, capacity, key_ptrs, val_ptrs, recv_total, bounds_out);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

}  // namespace usrt
