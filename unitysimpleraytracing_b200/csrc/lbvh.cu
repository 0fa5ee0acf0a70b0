// lbvh.cu -- K3 DistributeKeys, K4 TreeConstructor, K5 BVHConstructor (+ packed traversal arrays).
//
// K3 replaces the serial CPU loop of Assets/_Scripts/MeshBufferContainer.cs:154-169 (and its full
//    GPU->CPU->GPU round trip) with one single-pass scan kernel (decoupled look-back).
// K4 replaces kernel TreeConstructor, Assets/_Shaders/BVH/BVH.compute:94-149 (delta :23-33,
//    DetermineRange :35-52, FindSplit :54-92).
// K5 replaces kernel BVHConstructor, BVH.compute:172-220 (MergeAABB :152-170).

#include "usrt_internal.cuh"

namespace usrt {

namespace {

// =================================================================================================
// K3 -- new[0] = 0; new[i] = new[i-1] + max(k[i] - k[i-1], 1), wrapping uint32  (inclusive scan)
// =================================================================================================
constexpr int kScanBlock = 256;
constexpr int kScanVecPerThread = 4;                                   // 4 x uint4 = 16 keys per thread
constexpr int kScanTile = kScanBlock * kScanVecPerThread * 4;          // 4096 keys per tile
constexpr uint64_t kScanAggregate = 1ull << 32, kScanPrefix = 2ull << 32, kScanFlagMask = 3ull << 32;

// Each warp owns 512 consecutive keys as 4 rounds of one uint4 per lane (128-bit coalesced loads);
// within a round a lane holds 4 consecutive keys, lanes are consecutive, rounds are consecutive.
__global__ void __launch_bounds__(kScanBlock) k_distribute_keys(const uint32_t* __restrict__ src,
                                                                uint32_t* __restrict__ dst, uint32_t n,
                                                                uint64_t* __restrict__ status /* [0]=tile counter, [1+tile] */) {
    __shared__ uint32_t s_warp_total[kScanBlock / 32];
    __shared__ uint32_t s_tile_id;
    __shared__ uint32_t s_tile_prefix;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    if (tid == 0) s_tile_id = (uint32_t)atomicAdd(reinterpret_cast<unsigned long long*>(status), 1ull);
    __syncthreads();
    const uint32_t tile = s_tile_id;
    const uint32_t tile_base = tile * (uint32_t)kScanTile;
    const uint32_t warp_base = tile_base + warp * (32u * 4u * kScanVecPerThread);

    uint32_t k[kScanVecPerThread][4];
    uint32_t prev[kScanVecPerThread];
#pragma unroll
    for (int r = 0; r < kScanVecPerThread; ++r) {
        const uint32_t i0 = warp_base + (uint32_t)r * 128u + lane * 4u;
        if (i0 + 3 < n) {
            const uint4 q = __ldg(reinterpret_cast<const uint4*>(src + i0));
            k[r][0] = q.x; k[r][1] = q.y; k[r][2] = q.z; k[r][3] = q.w;
        } else {
#pragma unroll
            for (int c = 0; c < 4; ++c) k[r][c] = (i0 + c < n) ? __ldg(src + i0 + c) : 0u;
        }
        // key just before this lane's first key: previous lane's last, or one scalar load for lane 0
        uint32_t p = __shfl_up_sync(0xFFFFFFFFu, k[r][3], 1);
        if (lane == 0) p = (i0 > 0 && i0 < n) ? __ldg(src + i0 - 1) : 0u;
        prev[r] = p;
    }

    // per-element increments d[i] = max(k[i]-k[i-1], 1) (d[0] = 0, d[i>=n] = 0), local inclusive sums
    uint32_t carry = 0;                 // running total of this warp's earlier rounds
    uint32_t x[kScanVecPerThread][4];
#pragma unroll
    for (int r = 0; r < kScanVecPerThread; ++r) {
        const uint32_t i0 = warp_base + (uint32_t)r * 128u + lane * 4u;
        uint32_t before = prev[r], run = 0;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const uint32_t i = i0 + c;
            uint32_t d = k[r][c] - before;                 // wrapping, like C# unchecked uint
            d = d > 1u ? d : 1u;                           // Math.Max(uint, 1)
            if (i == 0 || i >= n) d = 0;
            before = k[r][c];
            run += d;
            x[r][c] = run;
        }
        uint32_t incl = run;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if (lane >= (uint32_t)o) incl += y;
        }
        const uint32_t excl = incl - run + carry;
#pragma unroll
        for (int c = 0; c < 4; ++c) x[r][c] += excl;
        carry += __shfl_sync(0xFFFFFFFFu, incl, 31);
    }
    if (lane == 0) s_warp_total[warp] = carry;
    __syncthreads();

    uint32_t warp_excl = 0, tile_total = 0;
#pragma unroll
    for (int w = 0; w < kScanBlock / 32; ++w) {
        const uint32_t t = s_warp_total[w];
        if (w < (int)warp) warp_excl += t;
        tile_total += t;
    }

    // decoupled look-back on one 64-bit word per tile (flag << 32 | running sum), 32 predecessors per
    // step: all tiles of a 1M-key scan start together, so a one-tile-at-a-time walk would be ~100 steps
    if (warp == 0) {
        uint64_t* my = status + 1 + tile;
        if (lane == 0) st_relaxed_u64(my, (tile == 0 ? kScanPrefix : kScanAggregate) | tile_total);
        uint32_t exclusive = 0;
        int32_t base = (int32_t)tile - 1;                       // nearest predecessor of this window
        while (base >= 0) {
            const int32_t t = base - (int32_t)lane;
            uint64_t s = kScanPrefix;                           // tiles before the first count as prefix 0
            if (t >= 0) {
                do { s = ld_relaxed_u64(status + 1 + t); } while ((s & kScanFlagMask) == 0);
            }
            const uint32_t has_prefix = __ballot_sync(0xFFFFFFFFu, (s & kScanPrefix) != 0);
            const int first = __ffs(has_prefix) - 1;            // nearest lane holding an inclusive prefix (-1: none)
            uint32_t v = (first < 0 || (int)lane <= first) ? (uint32_t)s : 0u;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
            exclusive += v;
            if (first >= 0) break;
            base -= 32;
        }
        if (lane == 0) {
            if (tile > 0) st_relaxed_u64(my, kScanPrefix | (uint32_t)(exclusive + tile_total));
            s_tile_prefix = exclusive;
        }
    }
    __syncthreads();
    const uint32_t add = s_tile_prefix + warp_excl;

#pragma unroll
    for (int r = 0; r < kScanVecPerThread; ++r) {
        const uint32_t i0 = warp_base + (uint32_t)r * 128u + lane * 4u;
        if (i0 + 3 < n) {
            *reinterpret_cast<uint4*>(dst + i0) = make_uint4(x[r][0] + add, x[r][1] + add, x[r][2] + add, x[r][3] + add);
        } else {
#pragma unroll
            for (int c = 0; c < 4; ++c)
                if (i0 + c < n) dst[i0 + c] = x[r][c] + add;
        }
    }
}

// K3 on 64-bit keys: the same recurrence in wrapping uint64. Not on the headline path, so the scan is
// the plain three-step form: per-tile sums, one block scanning the tile sums, per-tile apply. One key per thread.
constexpr int kDk64Tile = 1024;

__device__ __forceinline__ uint64_t dk64_increment(const uint64_t* __restrict__ src, uint32_t i, uint32_t n) {
    if (i == 0 || i >= n) return 0ull;
    const uint64_t d = __ldg(src + i) - __ldg(src + i - 1);             // wrapping, like C# unchecked ulong
    return d > 1ull ? d : 1ull;                                          // Math.Max(ulong, 1)
}

// inclusive scan over the block's 1024 threads; *total = the block's sum
__device__ __forceinline__ uint64_t block_inclusive_u64(uint64_t v, uint64_t* s_warp /* [32] */, uint64_t* total) {
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint64_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint64_t y = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if (lane >= (uint32_t)o) incl += y;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    uint64_t base = 0, sum = 0;
    for (uint32_t w = 0; w < blockDim.x / 32; ++w) { const uint64_t t = s_warp[w]; if (w < warp) base += t; sum += t; }
    __syncthreads();
    *total = sum;
    return base + incl;
}

__global__ void __launch_bounds__(kDk64Tile) k_dk64_tile_sums(const uint64_t* __restrict__ src, uint32_t n, uint64_t* __restrict__ sums) {
    __shared__ uint64_t s_warp[32];
    uint64_t total;
    block_inclusive_u64(dk64_increment(src, blockIdx.x * kDk64Tile + threadIdx.x, n), s_warp, &total);
    if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(kDk64Tile) k_dk64_scan_sums(uint64_t* __restrict__ sums, uint32_t tiles) {
    __shared__ uint64_t s_warp[32];
    uint64_t carry = 0;
    for (uint32_t base = 0; base < tiles; base += kDk64Tile) {
        const uint32_t i = base + threadIdx.x;
        const uint64_t v = i < tiles ? sums[i] : 0ull;
        uint64_t total;
        const uint64_t incl = block_inclusive_u64(v, s_warp, &total);
        if (i < tiles) sums[i] = carry + incl - v;                       // exclusive
        carry += total;
    }
}

__global__ void __launch_bounds__(kDk64Tile) k_dk64_apply(const uint64_t* __restrict__ src, uint64_t* __restrict__ dst, uint32_t n,
                                                          const uint64_t* __restrict__ sums) {
    __shared__ uint64_t s_warp[32];
    const uint32_t i = blockIdx.x * kDk64Tile + threadIdx.x;
    uint64_t total;
    const uint64_t incl = block_inclusive_u64(dk64_increment(src, i, n), s_warp, &total);
    if (i < n) dst[i] = sums[blockIdx.x] + incl;
}

// =================================================================================================
// K4 -- Karras LBVH topology, one thread per internal node
// =================================================================================================
// Key views. The reference builds the tree on 32-bit keys that DistributeKeys made unique (Keys32). Two opt-in
// variants reuse the same kernel: Keys32Tie skips DistributeKeys and breaks ties between equal Morton
// codes with the sorted position, i.e. works on the 64-bit key (code << 32 | index) as Karras 2012 proposes; Keys64
// reads 64-bit (63-bit Morton) keys made unique by the 64-bit DistributeKeys.
struct Keys32 {
    typedef uint32_t KeyT;
    const uint32_t* __restrict__ codes;
    __device__ __forceinline__ KeyT key(int i) const { return __ldg(codes + i); }
    // BVH.compute:18-21: clz32 = 31 - firstbithigh == __clz, 32 for 0
    static __device__ __forceinline__ int clz(KeyT x) { return __clz((int)x); }
};
struct Keys32Tie {
    typedef uint64_t KeyT;
    const uint32_t* __restrict__ codes;
    __device__ __forceinline__ KeyT key(int i) const { return ((uint64_t)__ldg(codes + i) << 32) | (uint32_t)i; }
    static __device__ __forceinline__ int clz(KeyT x) { return __clzll((long long)x); }
};
struct Keys64 {
    typedef uint64_t KeyT;
    const uint64_t* __restrict__ codes;
    __device__ __forceinline__ KeyT key(int i) const { return __ldg(codes + i); }
    static __device__ __forceinline__ int clz(KeyT x) { return __clzll((long long)x); }
};

template <typename View> struct KeyView {
    View v;
    int n;
    // BVH.compute:23-33
    __device__ __forceinline__ int delta(int x, int y) const {
        if (x >= 0 && x <= n - 1 && y >= 0 && y <= n - 1) return View::clz(v.key(x) ^ v.key(y));
        return -1;
    }
};

// Besides the reference's node arrays, K4 emits one private 32-bit "up link" per node for K5:
//   bits 0..29 parent index | bit 30 = this node is its parent's RIGHT child | bit 31 = the parent's
//   whole leaf range lies inside one kRefitBlock-aligned block of leaves (K5 merges it in shared memory).
constexpr int kRefitBlock = 128;                 // leaves per K5 CTA
constexpr uint32_t kUpRight = 1u << 30, kUpLocal = 1u << 31, kUpParentMask = (1u << 30) - 1;

template <typename View>
__global__ void __launch_bounds__(256) k_construct_tree(View view, uint32_t n,
                                                        usrt_internal_node* __restrict__ internal,
                                                        usrt_leaf_node* __restrict__ leaf,
                                                        uint32_t* __restrict__ up_internal,
                                                        uint32_t* __restrict__ up_leaf) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;                                            // BVH.compute:101
    const KeyView<View> kv{view, (int)n};
    typedef typename View::KeyT KeyT;
    const int idx = (int)i;

    // DetermineRange (BVH.compute:35-52); uint*int products wrap in 32 bits as in HLSL
    const int dl = kv.delta(idx, idx + 1) - kv.delta(idx, idx - 1);
    const int d = (dl > 0) - (dl < 0);
    const int dmin = kv.delta(idx, idx - d);
    uint32_t lmax = 2;
    while (kv.delta(idx, (int)((uint32_t)idx + lmax * (uint32_t)d)) > dmin) lmax *= 2;
    int l = 0;
    for (uint32_t t = lmax / 2; t >= 1; t /= 2)
        if (kv.delta(idx, (int)((uint32_t)idx + ((uint32_t)l + t) * (uint32_t)d)) > dmin) l += (int)t;
    const int j = idx + l * d;
    const int first = min(idx, j), last = max(idx, j);

    // FindSplit (BVH.compute:54-92)
    int split;
    {
        const KeyT first_code = view.key(first), last_code = view.key(last);
        if (first_code == last_code) {
            split = (first + last) >> 1;
        } else {
            const int common = View::clz(first_code ^ last_code);
            split = first;
            int step = last - first;
            do {
                step = (step + 1) >> 1;
                const int cand = split + step;
                if (cand < last) {
                    const int prefix = View::clz(first_code ^ view.key(cand));      // all bits for equal keys
                    if (prefix > common) split = cand;
                }
            } while (step > 1);
        }
    }

    // BVH.compute:111-147. Node i's own five words are two 8-byte stores + index; `parent` belongs
    // to the parent's thread. The root's parent keeps the NullLeaf sentinel (never written, :115).
    const uint32_t left = (uint32_t)split, right = (uint32_t)split + 1;
    const bool left_leaf = split == first, right_leaf = split + 1 == last;
    uint32_t* node = reinterpret_cast<uint32_t*>(internal + i);
    *reinterpret_cast<uint2*>(node + 0) = make_uint2(left, left_leaf ? USRT_LEAF_NODE : USRT_INTERNAL_NODE);
    *reinterpret_cast<uint2*>(node + 2) = make_uint2(right, right_leaf ? USRT_LEAF_NODE : USRT_INTERNAL_NODE);
    node[5] = i;
    if (left_leaf) *reinterpret_cast<uint2*>(leaf + left) = make_uint2(i, left);
    else internal[left].parent = i;
    if (right_leaf) *reinterpret_cast<uint2*>(leaf + right) = make_uint2(i, right);
    else internal[right].parent = i;

    const uint32_t link = i | (((uint32_t)first / kRefitBlock == (uint32_t)last / kRefitBlock) ? kUpLocal : 0u);
    (left_leaf ? up_leaf : up_internal)[left] = link;
    (right_leaf ? up_leaf : up_internal)[right] = link | kUpRight;
    if (i == 0) up_internal[0] = USRT_NULL;                            // the root (always node 0) has no parent
}

// =================================================================================================
// K5 -- bottom-up AABB propagation, one thread per leaf
// =================================================================================================
__device__ __forceinline__ float sel_min(float a, float b) { return a < b ? a : b; }
__device__ __forceinline__ float sel_max(float a, float b) { return a > b ? a : b; }

__device__ __forceinline__ float4 atom_exch_128(float4* p, float4 v) {
    const uint64_t lo = (uint64_t)__float_as_uint(v.x) | ((uint64_t)__float_as_uint(v.y) << 32);
    const uint64_t hi = (uint64_t)__float_as_uint(v.z) | ((uint64_t)__float_as_uint(v.w) << 32);
    uint64_t olo, ohi;
    asm volatile("{\n\t.reg .b128 vin, vout;\n\tmov.b128 vin, {%2, %3};\n\t"
                 "atom.relaxed.gpu.global.exch.b128 vout, [%4], vin;\n\tmov.b128 {%0, %1}, vout;\n\t}"
                 : "=l"(olo), "=l"(ohi) : "l"(lo), "l"(hi), "l"(p) : "memory");
    return make_float4(__uint_as_float((uint32_t)olo), __uint_as_float((uint32_t)(olo >> 32)),
                       __uint_as_float((uint32_t)ohi), __uint_as_float((uint32_t)(ohi >> 32)));
}

// One CTA owns kRefitBlock consecutive leaves. A node whose leaf range lies inside the block is
// "local": its two arrivals meet in shared memory -- each child deposits {box, ref} in its own slot,
// bumps a shared-memory counter, and the second arrival merges -- so ~99 % of the n-1 merges need
// no global atomic, no gpu-scope fence and no global re-read of the sibling box. Nodes spanning
// blocks keep the reference's rule (BVH.compute:184-215: first arrival leaves, second merges) but swap
// their boxes through 128-bit atomic exchanges instead of a counter + fence + reload.
__global__ void __launch_bounds__(kRefitBlock) k_construct_bvh(uint32_t n, const uint32_t* __restrict__ sorted_indices,
                                                               const float4* __restrict__ tri_aabb,
                                                               VertexSource vertices,
                                                               const usrt_internal_node* __restrict__ internal,
                                                               const uint32_t* __restrict__ up_internal,
                                                               const uint32_t* __restrict__ up_leaf,
                                                               float4* bvh, float4* slots, float4* packed_nodes,
                                                               float4* __restrict__ packed_tris) {
    __shared__ float4 s_slot[kRefitBlock][2][2];                       // [local node][side]{min|ref, max}
    __shared__ uint32_t s_count[kRefitBlock];
    __shared__ uint32_t s_up[kRefitBlock];                             // parent links of the block's internal nodes
    const uint32_t block_first = blockIdx.x * kRefitBlock;
    const uint32_t j = block_first + threadIdx.x;
    s_count[threadIdx.x] = 0;
    s_up[threadIdx.x] = (j + 1 < n) ? __ldg(up_internal + j) : USRT_NULL;   // local node ids lie in the block's range
    __syncthreads();
    if (j >= n) return;                                                // BVH.compute:179

    const uint32_t tri = __ldg(sorted_indices + j);
    float4 bmin, bmax;
    {
        // The leaf needs its triangle's vertices anyway (traversal-side copy in leaf order, triangle id in
        // a.w), so its padded box is recomputed from them with K1's exact operations
        // (MeshBufferContainer.cs:52-63) instead of gathering the 32-byte triangleAABB entry as well.
        const float4* t = vertices.base + (size_t)tri * vertices.stride;
        float4 a = ldg_vertex(t + 0), b = ldg_vertex(t + 1), c = ldg_vertex(t + 2);
        bmin = make_float4(__fsub_rn(sel_min(sel_min(a.x, b.x), c.x), 0.001f), __fsub_rn(sel_min(sel_min(a.y, b.y), c.y), 0.001f),
                           __fsub_rn(sel_min(sel_min(a.z, b.z), c.z), 0.001f), 0.0f);
        bmax = make_float4(__fadd_rn(sel_max(sel_max(a.x, b.x), c.x), 0.001f), __fadd_rn(sel_max(sel_max(a.y, b.y), c.y), 0.001f),
                           __fadd_rn(sel_max(sel_max(a.z, b.z), c.z), 0.001f), 0.0f);
        a.w = __uint_as_float(tri); b.w = 0.0f; c.w = 0.0f;
        packed_tris[(size_t)j * 3 + 0] = a;
        packed_tris[(size_t)j * 3 + 1] = b;
        packed_tris[(size_t)j * 3 + 2] = c;
    }

    uint32_t cur_ref = 0x80000000u | j;                                // ref: leaf => bit 31 | sorted position
    uint32_t link = __ldg(up_leaf + j);                                // BVH.compute:181
    while (link != USRT_NULL) {                                        // :182
        const uint32_t parent = link & kUpParentMask;
        const uint32_t side = (link >> 30) & 1u;                       // 0 = we are the left child
        float4 smin, smax;
        uint32_t sib_ref, next_link;
        if (link & kUpLocal) {
            const uint32_t lp = parent - block_first;
            next_link = s_up[lp];                                      // a local node's id is inside the block
            s_slot[lp][side][0] = make_float4(bmin.x, bmin.y, bmin.z, __uint_as_float(cur_ref));
            s_slot[lp][side][1] = bmax;
            __threadfence_block();
            const uint32_t old = atomicAdd(&s_count[lp], 1u);          // :184-189 first arrival leaves
            if (old == 0) break;
            __threadfence_block();
            smin = s_slot[lp][side ^ 1u][0];
            smax = s_slot[lp][side ^ 1u][1];
            sib_ref = __float_as_uint(smin.w);
        } else {
            // Cross-block node: the two arrivals swap boxes through a 32-byte slot with two 128-bit
            // atomic exchanges (ATOMG.EXCH.128). The box travels INSIDE the atomic, so there is no
            // store -> fence -> counter -> sibling-load chain per level, and the ~30 levels that span
            // blocks cost one L2 round trip each. Word 0 = {min.xyz, child ref}, word 1 = {max.xyz,
            // side}; .w == 0xFFFFFFFF marks an empty word. Whoever finds word 0 empty arrived first and
            // leaves (BVH.compute:184-189); the other one merges. It may have overtaken the first
            // arrival on word 1: then it re-exchanges its own max until the sibling's shows up (the first
            // arrival's second exchange is unconditional, so the wait is bounded). The merger finally
            // empties the slot again, which keeps the protocol re-runnable without a memset.
            next_link = __ldg(up_internal + parent);
            float4* slot = slots + (size_t)parent * 2;
            const float4 mine1 = make_float4(bmax.x, bmax.y, bmax.z, __uint_as_float(side));
            smin = atom_exch_128(slot, make_float4(bmin.x, bmin.y, bmin.z, __uint_as_float(cur_ref)));
            smax = atom_exch_128(slot + 1, mine1);
            if (__float_as_uint(smin.w) == USRT_NULL) break;
            while (__float_as_uint(smax.w) != (side ^ 1u)) smax = atom_exch_128(slot + 1, mine1);
            sib_ref = __float_as_uint(smin.w);
            const float4 empty = make_float4(__uint_as_float(USRT_NULL), __uint_as_float(USRT_NULL), __uint_as_float(USRT_NULL),
                                             __uint_as_float(USRT_NULL));
            __stcg(slot, empty);
            __stcg(slot + 1, empty);
        }
        const float4 lmin = side ? smin : bmin, lmax = side ? smax : bmax;
        const float4 rmin = side ? bmin : smin, rmax = side ? bmax : smax;
        const uint32_t lref = side ? sib_ref : cur_ref, rref = side ? cur_ref : sib_ref;

        // MergeAABB (:152-170), pads = 0
        bmin = make_float4(sel_min(lmin.x, rmin.x), sel_min(lmin.y, rmin.y), sel_min(lmin.z, rmin.z), 0.0f);
        bmax = make_float4(sel_max(lmax.x, rmax.x), sel_max(lmax.y, rmax.y), sel_max(lmax.z, rmax.z), 0.0f);
        __stcg(bvh + (size_t)parent * 2, bmin);                        // :215
        __stcg(bvh + (size_t)parent * 2 + 1, bmax);

        // packed traversal node: both child boxes + child refs in one 64-byte record
        float4* pn = packed_nodes + (size_t)parent * 4;
        pn[0] = make_float4(lmin.x, lmin.y, lmin.z, lmax.x);
        pn[1] = make_float4(lmax.y, lmax.z, rmin.x, rmin.y);
        pn[2] = make_float4(rmin.z, rmax.x, rmax.y, rmax.z);
        pn[3] = make_float4(__uint_as_float(lref), __uint_as_float(rref), 0.0f, 0.0f);

        cur_ref = parent;
        link = next_link;                                              // :217
    }
}

__global__ void k_count_corrupted(const usrt_leaf_node* __restrict__ leaf, const usrt_internal_node* __restrict__ internal,
                                  uint32_t n, uint32_t* __restrict__ out2) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    // MeshBufferContainer.cs:181-195
    if (i < n && leaf[i].index == USRT_NULL && leaf[i].parent == USRT_NULL) atomicAdd(out2 + 0, 1u);
    if (i + 1 < n && internal[i].index == USRT_NULL && internal[i].parent == USRT_NULL) atomicAdd(out2 + 1, 1u);
}

// rebuild the traversal-side packed arrays from the seven reference-layout buffers (a BVH
// that was built elsewhere -- loaded from a dump, or produced by the reference's own kernels).
__global__ void __launch_bounds__(256) k_pack_traversal(uint32_t n, const uint32_t* __restrict__ sorted_indices,
                                                        const float4* __restrict__ tri_aabb, const float4* __restrict__ tris,
                                                        const usrt_internal_node* __restrict__ internal,
                                                        const usrt_leaf_node* __restrict__ leaf, const float4* __restrict__ bvh,
                                                        float4* __restrict__ packed_nodes, float4* __restrict__ packed_tris) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {                                                       // leaf slot i -> triangle copy in leaf order
        const uint32_t tri = __ldg(sorted_indices + i);
        const float4* t = tris + (size_t)tri * 8;
        float4 a = ldg_vertex(t + 0), b = ldg_vertex(t + 1), c = ldg_vertex(t + 2);
        a.w = __uint_as_float(tri); b.w = 0.0f; c.w = 0.0f;
        packed_tris[(size_t)i * 3 + 0] = a; packed_tris[(size_t)i * 3 + 1] = b; packed_tris[(size_t)i * 3 + 2] = c;
    }
    if (i + 1 < n) {                                                   // internal node i
        const uint32_t* node = reinterpret_cast<const uint32_t*>(internal + i);
        const uint2 l = __ldg(reinterpret_cast<const uint2*>(node + 0)), r = __ldg(reinterpret_cast<const uint2*>(node + 2));
        float4 cmin[2], cmax[2];
        uint32_t ref[2];
        const uint2 ch[2] = {l, r};
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            if (ch[k].y == USRT_INTERNAL_NODE) {
                cmin[k] = __ldg(bvh + (size_t)ch[k].x * 2); cmax[k] = __ldg(bvh + (size_t)ch[k].x * 2 + 1);
                ref[k] = ch[k].x;
            } else {                                                   // Raytracing.compute:158: leafNodes[c].index -> sorted slot
                const uint32_t slot = __ldg(&leaf[ch[k].x].index);
                const uint32_t tri = __ldg(sorted_indices + slot);
                cmin[k] = __ldg(tri_aabb + (size_t)tri * 2); cmax[k] = __ldg(tri_aabb + (size_t)tri * 2 + 1);
                ref[k] = 0x80000000u | slot;
            }
        }
        float4* pn = packed_nodes + (size_t)i * 4;
        pn[0] = make_float4(cmin[0].x, cmin[0].y, cmin[0].z, cmax[0].x);
        pn[1] = make_float4(cmax[0].y, cmax[0].z, cmin[1].x, cmin[1].y);
        pn[2] = make_float4(cmin[1].z, cmax[1].x, cmax[1].y, cmax[1].z);
        pn[3] = make_float4(__uint_as_float(ref[0]), __uint_as_float(ref[1]), 0.0f, 0.0f);
    }
}


// ---- imported trees (usrt_upload_bvh): validation + the private K4 -> K5 parent links --------------------------
// A dump can be corrupt or hostile: every index the traversal or the refit will follow is checked on the device
// before the context accepts the tree. err[0] counts violations (the import is rejected), err[1] counts leaves whose
// `index` is not their own slot (legal for Raytracing.compute:158, but BVH.compute:199-208 refits by slot, so such a
// tree can be traced but not re-fitted).
constexpr int kMaxImportDepth = 64;                                  // internal nodes on a root-to-leaf path (trace stack :133 holds 64)

__global__ void __launch_bounds__(256) k_import_check_nodes(uint32_t n, const uint32_t* __restrict__ sorted_indices,
                                                            const usrt_internal_node* __restrict__ internal,
                                                            const usrt_leaf_node* __restrict__ leaf,
                                                            uint32_t* __restrict__ ref_internal, uint32_t* __restrict__ ref_leaf,
                                                            uint32_t* __restrict__ err) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        if (__ldg(sorted_indices + i) >= n || leaf[i].index >= n) atomicAdd(err, 1u);
        else if (leaf[i].index != i) atomicAdd(err + 1, 1u);
    }
    if (i + 1 < n) {
        const usrt_internal_node nd = internal[i];
        const uint32_t child[2] = {nd.leftNode, nd.rightNode}, type[2] = {nd.leftNodeType, nd.rightNodeType};
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            if (type[k] == USRT_LEAF_NODE && child[k] < n) atomicAdd(ref_leaf + child[k], 1u);
            else if (type[k] == USRT_INTERNAL_NODE && child[k] != 0 && child[k] < n - 1) atomicAdd(ref_internal + child[k], 1u);
            else atomicAdd(err, 1u);                                   // bad type, index out of range, or the root as a child
        }
    }
}

// every leaf and every internal node but the root is referenced by exactly one parent; the root by none
__global__ void __launch_bounds__(256) k_import_check_refs(uint32_t n, const uint32_t* __restrict__ ref_internal,
                                                           const uint32_t* __restrict__ ref_leaf, uint32_t* __restrict__ err) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && ref_leaf[i] != 1u) atomicAdd(err, 1u);
    if (i + 1 < n && ref_internal[i] != (i == 0 ? 0u : 1u)) atomicAdd(err, 1u);
}

// The up links K4 would have emitted. No node is marked block-local: an imported tree need not have Karras'
// numbering (node id inside its own leaf range), so K5 merges all of its nodes through the global exchange slots.
__global__ void __launch_bounds__(256) k_import_links(uint32_t n, const usrt_internal_node* __restrict__ internal,
                                                      uint32_t* __restrict__ up_internal, uint32_t* __restrict__ up_leaf) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i + 1 >= n) return;
    const usrt_internal_node nd = internal[i];
    (nd.leftNodeType == USRT_LEAF_NODE ? up_leaf : up_internal)[nd.leftNode] = i;
    (nd.rightNodeType == USRT_LEAF_NODE ? up_leaf : up_internal)[nd.rightNode] = i | kUpRight;
    if (i == 0) up_internal[0] = USRT_NULL;
}

// With one parent per node the links form a forest; it is one tree of bounded depth iff every leaf reaches node 0
// within kMaxImportDepth steps (a component cut off from the root always contains a leaf, which then fails here).
__global__ void __launch_bounds__(256) k_import_check_depth(uint32_t n, const uint32_t* __restrict__ up_internal,
                                                            const uint32_t* __restrict__ up_leaf, uint32_t* __restrict__ err) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    uint32_t link = up_leaf[j];
    uint32_t last = USRT_NULL;
    for (int step = 0; step < kMaxImportDepth && link != USRT_NULL; ++step) {
        last = link & kUpParentMask;
        link = up_internal[last];
    }
    if (link != USRT_NULL || last != 0u) atomicAdd(err, 1u);
}

}  // namespace

cudaError_t launch_pack_traversal(uint32_t n, const uint32_t* sorted_indices, const usrt_aabb* tri_aabb,
                                  const usrt_triangle* tris, const usrt_internal_node* internal, const usrt_leaf_node* leaf,
                                  const usrt_aabb* bvh, float4* packed_nodes, float4* packed_tris, cudaStream_t stream) {
    k_pack_traversal<<<(n + 255) / 256, 256, 0, stream>>>(n, sorted_indices, reinterpret_cast<const float4*>(tri_aabb),
                                                          reinterpret_cast<const float4*>(tris), internal, leaf,
                                                          reinterpret_cast<const float4*>(bvh), packed_nodes, packed_tris);
    return cudaGetLastError();
}

cudaError_t launch_import_validate(uint32_t n, const uint32_t* sorted_indices, const usrt_internal_node* internal,
                                   const usrt_leaf_node* leaf, uint32_t* up_internal, uint32_t* up_leaf, uint32_t* err2,
                                   cudaStream_t stream) {
    cudaError_t e;
    if ((e = cudaMemsetAsync(err2, 0, 8, stream)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(up_internal, 0, (size_t)n * 4, stream)) != cudaSuccess) return e;     // reference counters first,
    if ((e = cudaMemsetAsync(up_leaf, 0, (size_t)n * 4, stream)) != cudaSuccess) return e;         // then the links themselves
    const uint32_t grid = (n + 255) / 256;
    k_import_check_nodes<<<grid, 256, 0, stream>>>(n, sorted_indices, internal, leaf, up_internal, up_leaf, err2);
    k_import_check_refs<<<grid, 256, 0, stream>>>(n, up_internal, up_leaf, err2);
    return cudaGetLastError();
}

cudaError_t launch_import_links(uint32_t n, const usrt_internal_node* internal, uint32_t* up_internal, uint32_t* up_leaf,
                                uint32_t* err2, cudaStream_t stream) {
    const uint32_t grid = (n + 255) / 256;
    k_import_links<<<grid, 256, 0, stream>>>(n, internal, up_internal, up_leaf);
    k_import_check_depth<<<grid, 256, 0, stream>>>(n, up_internal, up_leaf, err2);
    return cudaGetLastError();
}

uint64_t distribute_status_bytes(uint32_t n) {
    const uint64_t tiles = ((uint64_t)n + kScanTile - 1) / kScanTile;
    return (tiles + 1) * 8;
}

cudaError_t launch_distribute_keys(const uint32_t* src, uint32_t* dst, uint32_t n, void* scan_status,
                                   cudaStream_t stream, int* launches) {
    if (n == 0) return cudaSuccess;
    cudaError_t e = cudaMemsetAsync(scan_status, 0, distribute_status_bytes(n), stream);
    if (e != cudaSuccess) return e;
    const uint32_t tiles = (uint32_t)(((uint64_t)n + kScanTile - 1) / kScanTile);
    k_distribute_keys<<<tiles, kScanBlock, 0, stream>>>(src, dst, n, static_cast<uint64_t*>(scan_status));
    if (launches) *launches += 1;
    return cudaGetLastError();
}

cudaError_t launch_distribute_keys64(const uint64_t* src, uint64_t* dst, uint32_t n, void* scan_status, cudaStream_t stream,
                                     int* launches) {
    if (n == 0) return cudaSuccess;
    const uint32_t tiles = (n + kDk64Tile - 1) / kDk64Tile;
    uint64_t* sums = static_cast<uint64_t*>(scan_status);               // >= tiles x 8 bytes (distribute_status_bytes64)
    k_dk64_tile_sums<<<tiles, kDk64Tile, 0, stream>>>(src, n, sums);
    k_dk64_scan_sums<<<1, kDk64Tile, 0, stream>>>(sums, tiles);
    k_dk64_apply<<<tiles, kDk64Tile, 0, stream>>>(src, dst, n, sums);
    if (launches) *launches += 3;
    return cudaGetLastError();
}
uint64_t distribute_status_bytes64(uint32_t n) { return ((uint64_t)(n + kDk64Tile - 1) / kDk64Tile) * 8; }

cudaError_t launch_construct_tree(const void* keys, int key_mode, uint32_t n, usrt_internal_node* internal,
                                  usrt_leaf_node* leaf, uint32_t* up_internal, uint32_t* up_leaf, cudaStream_t stream) {
    const uint32_t threads = n - 1, grid = (threads + 255) / 256;
    if (key_mode == 2)
        k_construct_tree<<<grid, 256, 0, stream>>>(Keys64{static_cast<const uint64_t*>(keys)}, n, internal, leaf, up_internal, up_leaf);
    else if (key_mode == 1)
        k_construct_tree<<<grid, 256, 0, stream>>>(Keys32Tie{static_cast<const uint32_t*>(keys)}, n, internal, leaf, up_internal, up_leaf);
    else
        k_construct_tree<<<grid, 256, 0, stream>>>(Keys32{static_cast<const uint32_t*>(keys)}, n, internal, leaf, up_internal, up_leaf);
    return cudaGetLastError();
}

cudaError_t launch_construct_bvh(uint32_t n, const uint32_t* sorted_indices, const usrt_aabb* tri_aabb,
                                 VertexSource vertices, const usrt_internal_node* internal,
                                 const uint32_t* up_internal, const uint32_t* up_leaf, usrt_aabb* bvh, float4* slots,
                                 float4* packed_nodes, float4* packed_tris, cudaStream_t stream) {
    k_construct_bvh<<<(n + kRefitBlock - 1) / kRefitBlock, kRefitBlock, 0, stream>>>(
        n, sorted_indices, reinterpret_cast<const float4*>(tri_aabb), vertices, internal,
        up_internal, up_leaf, reinterpret_cast<float4*>(bvh), slots, packed_nodes, packed_tris);
    return cudaGetLastError();
}

cudaError_t launch_count_corrupted(const usrt_leaf_node* leaf, const usrt_internal_node* internal, uint32_t n,
                                   uint32_t* out2, cudaStream_t stream) {
    cudaError_t e = cudaMemsetAsync(out2, 0, 8, stream);
    if (e != cudaSuccess) return e;
    k_count_corrupted<<<(n + 255) / 256, 256, 0, stream>>>(leaf, internal, n, out2);
    return cudaGetLastError();
}

}  // namespace usrt
