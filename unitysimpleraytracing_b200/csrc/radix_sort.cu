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
#define USRT_BIG_BLOCK 512
#define USRT_BIG_IPT 16
#define USRT_BIG_CTAS 2
#endif
using BigTile = TileCfg<USRT_BIG_BLOCK, USRT_BIG_IPT, USRT_BIG_CTAS>;   // 8192 pairs
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
// Shared memory per CTA (dynamic; BigTile 97 KB => 2 CTAs per SM):
//   s_tbl  [2 x warps][256] u32        32 KB   one word per (16-lane group, digit): running count (low 16 bits) and,
//                                              during a ranking round, the lane bits of the group (high 16 bits);
//                                              after the scan: first tile slot of that group's keys of that digit
//   s_pairs [tile] x {key, value}      64 KB   the tile staged in digit order
//   s_global_off[256], s_part[H][256], s_scan[warps], s_tile_id
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
template <typename Cfg, bool kHasValues, int kKeyBytes = 4> struct PassSmem {
    static constexpr int kTblBytes = 2 * Cfg::kWarps * kRadix * 4;
    static constexpr int kPairBytes = Cfg::kTile * (kKeyBytes + (kHasValues ? 4 : 0));
    static constexpr int kH = Cfg::kBlock / kRadix;                      // threads per digit in the scan step
    static constexpr int kTotal = kTblBytes + kPairBytes + kRadix * 4 + kH * kRadix * 4 + Cfg::kWarps * 4 + 16;
};

// kPeer: the multi-GPU bucket exchange. Instead of one output array, every digit has its own base address
// (key_ptrs[d] / val_ptrs[d], this rank's slice of the receive buffer of the GPU that owns bucket d, mapped
// through CUDA IPC): the stable scatter of the pass IS the all-to-all, written straight over NVLink.
// flags bit 0: test hook, every warp takes the order-independent ranking path.
// KeyT = uint32_t (the reference's ComputeBufferSorter<uint,uint>) or uint64_t (its GetRadix is generic over uint /
// ulong, ComputeBufferSorter.cs:179-191): 64-bit keys run 8 passes and stage keys and values in separate arrays.
template <typename Cfg, typename StatusT, bool kHasValues, bool kPeer = false, typename KeyT = uint32_t>
__global__ void __launch_bounds__(Cfg::kBlock, Cfg::kCtasPerSM)
k_onesweep(const KeyT* __restrict__ keys_in, const uint32_t* __restrict__ vals_in, KeyT* __restrict__ keys_out,
           uint32_t* __restrict__ vals_out, uint32_t n, int shift, const uint32_t* __restrict__ digit_base /* [256] */,
           uint32_t* __restrict__ tile_counter, StatusT* __restrict__ status /* [tiles][256], zeroed */, uint32_t flags,
           const unsigned long long* __restrict__ key_ptrs = nullptr, const unsigned long long* __restrict__ val_ptrs = nullptr) {
    using ST = StatusTraits<StatusT>;
    using SM = PassSmem<Cfg, kHasValues, (int)sizeof(KeyT)>;
    constexpr bool kWide = sizeof(KeyT) == 8;
    static_assert(!(kWide && kPeer), "the multi-GPU bucket exchange is built for 32-bit keys");
    constexpr int kBlock = Cfg::kBlock, kIPT = Cfg::kIPT, kTile = Cfg::kTile, kWarps = Cfg::kWarps;
    constexpr int kGroups = 2 * kWarps, kH = SM::kH, kGP = kGroups / kH;      // groups per scan thread
    static_assert(kIPT % 4 == 0 && 16 * kIPT <= 256, "ranks are packed four to a register (< 256 each)");
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

    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;

    // dynamic tile id: a tile only ever waits on tiles that already started => forward progress
    if (tid == 0) *s_tile_id = atomicAdd(tile_counter, 1u);
    {
        uint4* z = reinterpret_cast<uint4*>(smem);
#pragma unroll
        for (int i = 0; i < SM::kTblBytes / 16 / kBlock; ++i) z[tid + i * kBlock] = make_uint4(0, 0, 0, 0);
    }
    __syncthreads();
    const uint32_t tile = *s_tile_id;
    const uint32_t tile_base = tile * (uint32_t)kTile;
    const uint32_t valid = min((uint32_t)kTile, n - tile_base);

    // Group-major item map: 16-lane group G = 2*warp + lane/16 owns the 16*IPT consecutive keys from G*16*IPT,
    // item i of lane l16 is key G*16*IPT + 16*i + l16 -- so (round, lane) order inside a group IS memory order,
    // which is what makes the per-group ranks stable.
    const uint32_t l16 = lane & 15u;
    const uint32_t group = 2u * warp + (lane >> 4);
    const uint32_t item0 = group * (16u * kIPT) + l16;
    KeyT key[kIPT];
#ifndef USRT_LAB_SHFL_LOADS
    // direct form: every load touches two 64-byte segments, one per group (two L1 wavefronts instead of one)
#pragma unroll
    for (int i = 0; i < kIPT; ++i) {
        const uint32_t idx = item0 + (uint32_t)i * 16u;
        key[i] = idx < valid ? __ldg(keys_in + tile_base + idx) : ~(KeyT)0;      // tail pads sort last, never stored
    }
#else
    // (measured slower, 0.348 vs 0.334 ms per pass at 2^26: the shuffles cost more than the extra load wavefronts)
    // Warp-striped 128-byte loads over the warp's 32*IPT keys, then one lane-xor-16 shuffle per pair of loads hands
    // each half-warp the half of the chunk it owns: load i < IPT/2 holds group A's items 2i (lanes 0-15) and 2i+1
    // (lanes 16-31); load i + IPT/2 holds group B's items 2i and 2i+1 the same way.
    const bool upper = (lane & 16u) != 0u;
    const uint32_t warp_first = warp * (32u * kIPT) + lane;
    auto load_items = [&](const uint32_t* __restrict__ src, uint32_t (&item)[kIPT], uint32_t pad) {
#pragma unroll
        for (int i = 0; i < kIPT / 2; ++i) {
            const uint32_t ia = warp_first + (uint32_t)i * 32u, ib = ia + 16u * kIPT;
            const uint32_t a = ia < valid ? __ldg(src + tile_base + ia) : pad;
            const uint32_t b = ib < valid ? __ldg(src + tile_base + ib) : pad;
            const uint32_t got = __shfl_xor_sync(0xFFFFFFFFu, upper ? a : b, 16);
            item[2 * i] = upper ? got : a;
            item[2 * i + 1] = upper ? b : got;
        }
    };
    load_items(keys_in, key, 0xFFFFFFFFu);                   // tail pads sort last, never stored
#endif

    // stable rank of every key among the keys of its group with the same digit (< 256: four per register)
    uint32_t rank4[kIPT / 4];
    uint32_t* tbl = s_tbl + group * kRadix;
    {
        const uint32_t add = (0x10000u << l16) | 1u, bit = 0x10000u << l16;
        const uint32_t not_below = 0xFFFFu << l16;             // my own lane and the higher ones of my group
        uint32_t out_of_order = flags & 1u;
#pragma unroll
        for (int i = 0; i < kIPT; ++i) {
            const uint32_t d = (uint32_t)(key[i] >> shift) & 255u;
            const uint32_t old = atomicAdd(tbl + d, add);
            __syncwarp();
            atomicSub(tbl + d, bit);
#ifndef USRT_LAB_NO_SYNC2
            __syncwarp();
#endif
            out_of_order |= (old >> 16) & not_below;
            const uint32_t r = old & 0xFFFFu;
            if ((i & 3) == 0) rank4[i >> 2] = r; else rank4[i >> 2] |= r << ((i & 3) * 8);
        }
        if (__any_sync(0xFFFFFFFFu, out_of_order != 0u)) {
            // Order-independent ranking (never taken on B200 unless forced): start the warp's two tables over.
            uint32_t* mine = s_tbl + 2u * warp * kRadix;
#pragma unroll
            for (int i = 0; i < 2 * kRadix / 32; ++i) mine[lane + 32 * i] = 0u;
            __syncwarp();
            const uint32_t below = (1u << l16) - 1u;
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
                const uint32_t r = total - __popc(peers) + before;
                if ((i & 3) == 0) rank4[i >> 2] = r; else rank4[i >> 2] |= r << ((i & 3) * 8);
            }
        }
    }
    __syncthreads();

    // values are fetched now, so their latency hides behind the scan and the look-back below
    uint32_t val[kHasValues ? kIPT : 1];
    if (kHasValues) {
#ifndef USRT_LAB_SHFL_LOADS
#pragma unroll
        for (int i = 0; i < kIPT; ++i) {
            const uint32_t idx = item0 + (uint32_t)i * 16u;
            val[i] = idx < valid ? __ldg(vals_in + tile_base + idx) : 0u;
        }
#else
        load_items(vals_in, reinterpret_cast<uint32_t (&)[kIPT]>(val), 0u);
#endif
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
        const uint32_t slot = tbl[dg] + ((rank4[i >> 2] >> ((i & 3) * 8)) & 0xFFu);
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
inline uint32_t pass_flags() {
    static const uint32_t f = getenv("USRT_FORCE_SLOW_RANK") != nullptr ? 1u : 0u;   // test hook: order-independent ranking everywhere
    return f;
}
inline uint64_t status_words_bytes(uint64_t count) { return (uint64_t)num_tiles(count) * kRadix * (wide_status(count) ? 8 : 4); }

template <typename Cfg, typename StatusT, bool kHasValues>
cudaError_t launch_pass(const uint32_t* ki, const uint32_t* vi, uint32_t* ko, uint32_t* vo, uint64_t count, int shift,
                        const uint32_t* digit_base, uint32_t* tile_counter, void* status, cudaStream_t stream) {
    constexpr int smem = PassSmem<Cfg, kHasValues>::kTotal;
    cudaFuncSetAttribute(k_onesweep<Cfg, StatusT, kHasValues>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    k_onesweep<Cfg, StatusT, kHasValues><<<num_tiles(count), Cfg::kBlock, smem, stream>>>(
        ki, vi, ko, vo, (uint32_t)count, shift, digit_base, tile_counter, static_cast<StatusT*>(status), pass_flags());
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

// ---- ComputeBufferSorter<ulong, uint>: 8 passes x 8 bits over 64-bit keys (ComputeBufferSorter.cs:179-191) -----------
namespace {
using Tile64 = TileCfg<512, 8, 2>;                   // 4096 pairs: 64-bit keys take two registers each
inline uint32_t num_tiles64(uint64_t count) { return (uint32_t)((count + Tile64::kTile - 1) / Tile64::kTile); }
inline uint64_t status_bytes64(uint64_t count) { return (uint64_t)num_tiles64(count) * kRadix * (wide_status(count) ? 8 : 4); }

template <typename StatusT, bool kHasValues>
cudaError_t launch_pass64(const uint64_t* ki, const uint32_t* vi, uint64_t* ko, uint32_t* vo, uint64_t count, int shift,
                          const uint32_t* digit_base, uint32_t* tile_counter, void* status, cudaStream_t stream) {
    constexpr int smem = PassSmem<Tile64, kHasValues, 8>::kTotal;
    auto kern = k_onesweep<Tile64, StatusT, kHasValues, false, uint64_t>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    kern<<<num_tiles64(count), Tile64::kBlock, smem, stream>>>(ki, vi, ko, vo, (uint32_t)count, shift, digit_base, tile_counter,
                                                               static_cast<StatusT*>(status), pass_flags(), nullptr, nullptr);
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
    if ((e = cudaMemsetAsync(s.status, 0, kHeaderWords * 4 + kPasses * pass_bytes, stream)) != cudaSuccess) return e;
    cudaFuncSetAttribute(k_histogram64, cudaFuncAttributeMaxDynamicSharedMemorySize, kHistSmemBytes);
    const uint32_t grid = (uint32_t)std::min<uint64_t>(std::max<uint64_t>((count + kHistThreads - 1) / kHistThreads, 1), (uint64_t)kNumSMs);
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
        const bool wide = wide_status(count);
        if (vi != nullptr)
            e = wide ? launch_pass64<uint64_t, true>(ki, vi, ko, vo, count, pass * kRadixBits, s.hist + pass * kRadix, counters + pass, st, stream)
                     : launch_pass64<uint32_t, true>(ki, vi, ko, vo, count, pass * kRadixBits, s.hist + pass * kRadix, counters + pass, st, stream);
        else
            e = wide ? launch_pass64<uint64_t, false>(ki, vi, ko, vo, count, pass * kRadixBits, s.hist + pass * kRadix, counters + pass, st, stream)
                     : launch_pass64<uint32_t, false>(ki, vi, ko, vo, count, pass * kRadixBits, s.hist + pass * kRadix, counters + pass, st, stream);
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
            src_keys, src_vals, static_cast<uint32_t*>(nullptr), nullptr, (uint32_t)count, bit_offset, nullptr, counters, st, pass_flags(), key_ptrs, val_ptrs);
    } else {
        constexpr int smem = PassSmem<BigTile, true>::kTotal;
        cudaFuncSetAttribute(k_onesweep<BigTile, uint32_t, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        k_onesweep<BigTile, uint32_t, true, true><<<num_tiles(count), BigTile::kBlock, smem, stream>>>(
            src_keys, src_vals, static_cast<uint32_t*>(nullptr), nullptr, (uint32_t)count, bit_offset, nullptr, counters, st, pass_flags(), key_ptrs, val_ptrs);
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
