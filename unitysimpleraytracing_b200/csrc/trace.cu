// trace.cu -- K6: stack-based BVH traversal with the reference's slab and Moller-Trumbore tests.
//
// Replaces kernel Raytracing, Assets/_Shaders/Raytracing/Raytracing.compute:105-176 (ray generation
// :108-126, RayBoxIntersection :75-87, RayTriangleIntersection :37-73, CheckTriangle :89-103), up to
// and excluding the shading epilogue (:178-184). Output is the RaycastResult hit record (:30-35).
//
// Visiting semantics are the reference's exactly (strict mode): DFS from node 0, left child then
// right child handled at the parent's visit (leaf => tested now, internal => pushed), so the right
// internal subtree is walked before the left one; a candidate replaces the best hit only if strictly
// closer; no distance culling, no t > 0 test. The reference tests a node's box when it is POPPED;
// here the parent's 64-byte packed record carries both child boxes, so a child is tested before it
// is PUSHED -- the same boxes are tested against the same ray and the surviving visit order is
// identical, but each visit costs one 64-byte record instead of a 32-byte box + 24-byte node +
// 8-byte leaf + 4-byte index chase. Leaf boxes are the triangles' own padded AABBs, i.e. the
// CheckTriangle box test (:91).
//
// fp32 arithmetic is written with round-to-nearest intrinsics in the oracle's canonical order
// (SURVEY.md 8a): no FMA contraction, IEEE division and square root.

#include "usrt_internal.cuh"

namespace usrt {

namespace {

struct Ray { float ox, oy, oz, dx, dy, dz, ix, iy, iz; };

__device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float dot3(float ax, float ay, float az, float bx, float by, float bz) {
    return add(add(mul(ax, bx), mul(ay, by)), mul(az, bz));
}

// One 64-byte packed node = two 256-bit loads (LDG.E.256, new on sm_100): the lanes of a warp sit on ~9
// different nodes on average, and every load instruction pays one L1 wavefront per distinct line, so
// halving the instruction count per node halves the LSU-pipe cost of the walk.
struct Node8 { float4 lo, hi; };
__device__ __forceinline__ Node8 ldg256(const float4* p) {
    Node8 r;
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(r.lo.x), "=f"(r.lo.y), "=f"(r.lo.z), "=f"(r.lo.w), "=f"(r.hi.x), "=f"(r.hi.y), "=f"(r.hi.z), "=f"(r.hi.w)
                 : "l"(p));
    return r;
}

// Constants.cginc:7 -- integer literal 0x7F7FFFFF converted to float
__device__ __forceinline__ float max_float() { return __uint_as_float(0x4EFF0000u); }

// Raytracing.compute:75-87; returns tmin through *entry for the (non-parity) culled mode
__device__ __forceinline__ bool ray_box(float bminx, float bminy, float bminz, float bmaxx, float bmaxy, float bmaxz,
                                        const Ray& r, float* entry) {
    const float t1x = mul(sub(bminx, r.ox), r.ix), t2x = mul(sub(bmaxx, r.ox), r.ix);
    const float t1y = mul(sub(bminy, r.oy), r.iy), t2y = mul(sub(bmaxy, r.oy), r.iy);
    const float t1z = mul(sub(bminz, r.oz), r.iz), t2z = mul(sub(bmaxz, r.oz), r.iz);
    const float tmin = fmaxf(fminf(t1x, t2x), fmaxf(fminf(t1y, t2y), fminf(t1z, t2z)));
    const float tmax = fminf(fmaxf(t1x, t2x), fminf(fmaxf(t1y, t2y), fmaxf(t1z, t2z)));
    *entry = tmin;
    return tmax > tmin && tmax > 0.0f;
}

// Raytracing.compute:37-73 + the strict '<' replace of CheckTriangle :95-99
__device__ __forceinline__ void ray_triangle(const Ray& r, const float4 v0, const float4 v1, const float4 v2,
                                             usrt_raycast_result& best) {
    const float e1x = sub(v1.x, v0.x), e1y = sub(v1.y, v0.y), e1z = sub(v1.z, v0.z);
    const float e2x = sub(v2.x, v0.x), e2y = sub(v2.y, v0.y), e2z = sub(v2.z, v0.z);
    // pvec = cross(dir, e2)
    const float px = sub(mul(r.dy, e2z), mul(r.dz, e2y));
    const float py = sub(mul(r.dz, e2x), mul(r.dx, e2z));
    const float pz = sub(mul(r.dx, e2y), mul(r.dy, e2x));
    const float det = dot3(e1x, e1y, e1z, px, py, pz);
    if (det < 1e-8f && det > -1e-8f) return;
    const float inv_det = __fdiv_rn(1.0f, det);
    const float tx = sub(r.ox, v0.x), ty = sub(r.oy, v0.y), tz = sub(r.oz, v0.z);
    const float u = mul(dot3(tx, ty, tz, px, py, pz), inv_det);
    if (u < 0.0f || u > 1.0f) return;
    // qvec = cross(tvec, e1)
    const float qx = sub(mul(ty, e1z), mul(tz, e1y));
    const float qy = sub(mul(tz, e1x), mul(tx, e1z));
    const float qz = sub(mul(tx, e1y), mul(ty, e1x));
    const float v = mul(dot3(r.dx, r.dy, r.dz, qx, qy, qz), inv_det);
    if (v < 0.0f || add(u, v) > 1.0f) return;
    const float dist = mul(dot3(e2x, e2y, e2z, qx, qy, qz), inv_det);
    if (dist < best.distance) {
        best.distance = dist;
        best.triangleIndex = __float_as_uint(v0.w);
        best.uv[0] = u;
        best.uv[1] = v;
    }
}

// kNearFirst (mode 2, with culling only): of two internal children that are both hit, the one the ray enters first is
// walked first, so the closest hit is found early and more of the far side is culled. Neither culled mode is part of the
// parity contract: ties between equally distant triangles can resolve differently from the reference's visiting order.
template <bool kCulled, bool kNearFirst = false>
__device__ __forceinline__ usrt_raycast_result traverse(const TraceScene& s, const Ray& ray) {
    usrt_raycast_result best;
    best.distance = max_float();                       // Raytracing.compute:129-131
    best.triangleIndex = 0;
    best.uv[0] = 0.0f; best.uv[1] = 0.0f;

    // node 0 is popped and its own box tested first (:135-146)
    {
        const float4 rmin = __ldg(reinterpret_cast<const float4*>(s.bvh));
        const float4 rmax = __ldg(reinterpret_cast<const float4*>(s.bvh) + 1);
        float entry;
        if (!ray_box(rmin.x, rmin.y, rmin.z, rmax.x, rmax.y, rmax.z, ray, &entry)) return best;
    }

    // The reference pushes left then right and pops right first (:148-175). Equivalent walk with the
    // current node in a register: continue straight into the right child when it is an internal hit,
    // deferring the left one on the stack; only the deferred siblings ever touch the stack (<= tree
    // depth <= 33 entries, because DistributeKeys makes the 32-bit keys unique).
    uint32_t stack[64];                                // :133
    int sp = 0;
    uint32_t index = 0;
    while (true) {
        const float4* pn = s.packed_nodes + (size_t)index * 4;
        const Node8 n01 = ldg256(pn), n23 = ldg256(pn + 2);
        const float4 q0 = n01.lo, q1 = n01.hi, q2 = n23.lo, q3 = n23.hi;
        const uint32_t lref = __float_as_uint(q3.x), rref = __float_as_uint(q3.y);

        float lentry, rentry;
        bool lhit = ray_box(q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, ray, &lentry);
        bool rhit = ray_box(q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, ray, &rentry);
        if (kCulled) lhit = lhit && !(lentry > best.distance);
        // left child (:148-160): a leaf is tested now, an internal node is visited later
        if (lhit && (lref & 0x80000000u)) {
            const float4* t = s.packed_tris + (size_t)(lref & 0x7FFFFFFFu) * 3;
            ray_triangle(ray, __ldg(t + 0), __ldg(t + 1), __ldg(t + 2), best);
        }
        if (kCulled) rhit = rhit && !(rentry > best.distance);
        // right child (:162-175)
        if (rhit && (rref & 0x80000000u)) {
            const float4* t = s.packed_tris + (size_t)(rref & 0x7FFFFFFFu) * 3;
            ray_triangle(ray, __ldg(t + 0), __ldg(t + 1), __ldg(t + 2), best);
        }
        const bool lgo = lhit && !(lref & 0x80000000u), rgo = rhit && !(rref & 0x80000000u);
        if (kNearFirst && lgo && rgo) {
            const bool left_first = lentry < rentry;
            stack[sp++] = left_first ? rref : lref;
            index = left_first ? lref : rref;
        } else if (rgo) {
            if (lgo) stack[sp++] = lref;
            index = rref;
        } else if (lgo) {
            index = lref;
        } else {
            if (sp == 0) break;
            index = stack[--sp];
        }
    }
    return best;
}

// Strict mode, warp-cooperative form. The reference never culls, so the walk (which nodes are visited, in
// which order) does not depend on any intersection result: only the ORDER in which a ray's candidate
// triangles are tested matters (strict '<', first visited wins). The ~75-instruction Moller-Trumbore block
// is therefore taken out of the node loop: a lane appends the leaves it reaches to a small per-lane FIFO
// (shared memory, [slot][thread] so it is bank-conflict-free) and the warp runs one "triangle round" --
// every lane pops its oldest pending leaf -- only when some lane's FIFO is nearly full, or at the end. In
// the plain loop the block ran in ~2/3 of all iterations with one or two lanes active; here it runs a
// dozen times per warp with most lanes active. Per-ray test order is unchanged, so results are identical.
constexpr int kLeafFifo = 8;                            // entries per lane; an iteration appends at most 2

__device__ __forceinline__ usrt_raycast_result traverse_strict(const TraceScene& s, const Ray& ray, bool valid,
                                                               uint32_t (*fifo)[128]) {
    usrt_raycast_result best;
    best.distance = max_float();                       // Raytracing.compute:129-131
    best.triangleIndex = 0;
    best.uv[0] = 0.0f; best.uv[1] = 0.0f;
    // slot k of this lane = fifo[k][tid]; addressed as a 32-bit shared-window offset so that a push is
    // {and, shift-add, st.shared} under a predicate instead of a generic-pointer rebuild behind a branch
    const uint32_t fifo_sh = (uint32_t)__cvta_generic_to_shared(&fifo[0][threadIdx.x]);
    constexpr uint32_t kSlotBytes = 128u * 4u;
    constexpr uint32_t kDone = 0xFFFFFFFFu;             // no node left for this lane

    uint32_t index = 0;
    if (valid) {                                        // node 0 is popped and its own box tested first (:135-146)
        const float4 rmin = __ldg(reinterpret_cast<const float4*>(s.bvh));
        const float4 rmax = __ldg(reinterpret_cast<const float4*>(s.bvh) + 1);
        float entry;
        if (!ray_box(rmin.x, rmin.y, rmin.z, rmax.x, rmax.y, rmax.z, ray, &entry)) index = kDone;
    } else {
        index = kDone;
    }
    uint32_t walking = __popc(__ballot_sync(0xFFFFFFFFu, index != kDone));   // warp-uniform: lanes with nodes left
    if (walking == 0) return best;                      // off-frame warp, or every ray misses the root box
    uint32_t stack[64];                                // :133 (deferred left siblings only, <= 33 deep)
    int sp = 0;
    uint32_t head = 0, tail = 0;                        // FIFO of this lane: running counters, slot = counter & 7
    while (true) {
        bool finished = false;
        if (index != kDone) {
            const float4* pn = s.packed_nodes + (size_t)index * 4;
            const Node8 n01 = ldg256(pn), n23 = ldg256(pn + 2);
            const float4 q0 = n01.lo, q1 = n01.hi, q2 = n23.lo, q3 = n23.hi;
            const uint32_t lref = __float_as_uint(q3.x), rref = __float_as_uint(q3.y);
            float lentry, rentry;
            const bool lhit = ray_box(q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, ray, &lentry);
            const bool rhit = ray_box(q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, ray, &rentry);
            // left child then right child (:148-175): leaves are queued in that order (leaf bit kept, cleared at the pop)
            if (lhit && (lref & 0x80000000u)) {
                asm volatile("st.shared.u32 [%0], %1;" ::"r"(fifo_sh + (tail & (kLeafFifo - 1)) * kSlotBytes), "r"(lref) : "memory");
                ++tail;
            }
            if (rhit && (rref & 0x80000000u)) {
                asm volatile("st.shared.u32 [%0], %1;" ::"r"(fifo_sh + (tail & (kLeafFifo - 1)) * kSlotBytes), "r"(rref) : "memory");
                ++tail;
            }
            const bool lgo = lhit && !(lref & 0x80000000u), rgo = rhit && !(rref & 0x80000000u);
            // next node, written as selects (the lanes of a warp disagree here all the time): right child first,
            // the left one deferred; with no child to enter, the youngest deferred sibling; else this lane is done
            if (lgo && rgo) stack[sp++] = lref;
            const bool none = !(lgo || rgo);
            const bool pop = none && sp != 0;
            uint32_t next = rgo ? rref : lref;
            if (pop) next = stack[--sp];
            finished = none && !pop;
            index = finished ? kDone : next;
        }
        // One vote per visit: anything to do besides walking on? (a FIFO that the next visit could overflow, or a
        // lane that has just run out of nodes)
        if (!__any_sync(0xFFFFFFFFu, finished || tail - head > (uint32_t)(kLeafFifo - 2))) continue;
        walking -= __popc(__ballot_sync(0xFFFFFFFFu, finished));
        // triangle rounds: while some lane could overflow on its next visit, or to drain at the end
        while (__any_sync(0xFFFFFFFFu, tail - head > (uint32_t)(kLeafFifo - 2)) ||
               (walking == 0 && __any_sync(0xFFFFFFFFu, tail != head))) {
            if (tail != head) {
                uint32_t leaf;
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(leaf) : "r"(fifo_sh + (head & (kLeafFifo - 1)) * kSlotBytes) : "memory");
                leaf &= 0x7FFFFFFFu;
                ++head;
                const float4* t = s.packed_tris + (size_t)leaf * 3;
                ray_triangle(ray, __ldg(t + 0), __ldg(t + 1), __ldg(t + 2), best);
            }
        }
        if (walking == 0) break;
    }
    return best;
}

// Raytracing.compute:108-126 with the uniforms of RaytracingMeshDrawer.cs:78-81
__device__ __forceinline__ Ray primary_ray(const PrimaryParams& p, uint32_t x, uint32_t y) {
    const float near_plane = p.near_plane;
    const float fov = p.tan_half_fov;
    const float height = mul(mul(2.0f, near_plane), fov);
    const float fw = __int2float_rn(p.width), fh = __int2float_rn(p.height);
    const float width = __fdiv_rn(mul(fw, height), fh);
    const float dx = add(__fdiv_rn(-width, 2.0f), mul(__fdiv_rn(width, fw), add(__uint2float_rn(x), 0.5f)));
    const float dy = add(__fdiv_rn(-height, 2.0f), mul(__fdiv_rn(height, fh), add(__uint2float_rn(y), 0.5f)));
    const float dz = -near_plane;
    const float* m = p.m;
    Ray r;
    // origin = mul(M, (0,0,0,1)), dir = mul(M, (dir,0)): row . vector, left to right, w term included
    r.ox = add(add(add(mul(m[0], 0.0f), mul(m[1], 0.0f)), mul(m[2], 0.0f)), mul(m[3], 1.0f));
    r.oy = add(add(add(mul(m[4], 0.0f), mul(m[5], 0.0f)), mul(m[6], 0.0f)), mul(m[7], 1.0f));
    r.oz = add(add(add(mul(m[8], 0.0f), mul(m[9], 0.0f)), mul(m[10], 0.0f)), mul(m[11], 1.0f));
    const float wx = add(add(add(mul(m[0], dx), mul(m[1], dy)), mul(m[2], dz)), mul(m[3], 0.0f));
    const float wy = add(add(add(mul(m[4], dx), mul(m[5], dy)), mul(m[6], dz)), mul(m[7], 0.0f));
    const float wz = add(add(add(mul(m[8], dx), mul(m[9], dy)), mul(m[10], dz)), mul(m[11], 0.0f));
    const float len = __fsqrt_rn(dot3(wx, wy, wz, wx, wy, wz));
    r.dx = __fdiv_rn(wx, len); r.dy = __fdiv_rn(wy, len); r.dz = __fdiv_rn(wz, len);
    r.ix = __fdiv_rn(1.0f, r.dx); r.iy = __fdiv_rn(1.0f, r.dy); r.iz = __fdiv_rn(1.0f, r.dz);
    return r;
}

__device__ __forceinline__ void store_hit(usrt_raycast_result* out, size_t i, const usrt_raycast_result& h) {
    reinterpret_cast<float4*>(out)[i] = make_float4(h.distance, __uint_as_float(h.triangleIndex), h.uv[0], h.uv[1]);
}

// One warp = a 16 x 2 pixel tile, one CTA = 4 warps = 16 x 8 pixels. Only on-screen pixels are traced
// (the reference dispatches (W/32+1) x (H/32+1) groups with no bounds guard, RaytracingMeshDrawer.cs:83).
// Warp shapes measured on B200, strict mode (tools/trace_lab.py): configs[1] at 1080p 8x4 0.910 ms, 4x8 0.929, 16x2 0.881,
// 32x1 0.867; configs[0] at 512x512 8x4 0.846, 4x8 0.824, 16x2 0.852, 32x1 0.922 -- scene dependent; 16x2 gains 3 % on
// the headline scene and loses under 1 % on the soup, 32x1 loses 9 % there. Also measured there, none faster: deferred
// siblings in shared memory (8-24 slots), 4- and 16-entry leaf FIFOs, register caps for 10 and 12 CTAs per SM, and the
// .L2::64B fetch-size qualifier on the node loads (0.879 -> 0.910 ms; incoherent rays over a 2^22-triangle soup 18.3 ->
// 19.0 ms: the other half of a 128-byte line is the sibling node, i.e. a prefetch that pays) or on the triangle loads.
#ifndef USRT_TRACE_WARP_W
#define USRT_TRACE_WARP_W 16                                      // lab switch: 8 (8x4), 4 (4x8), 16 (16x2), 32 (32x1)
#endif
#if USRT_TRACE_WARP_W == 32
constexpr int kTileW = 32, kTileH = 4;
#else
constexpr int kTileW = 16, kTileH = 8;
#endif

template <int kMode>                                              // 0 strict, 1 culled, 2 culled + near child first
__global__ void __launch_bounds__(128) k_trace_primary(TraceScene scene, PrimaryParams p, usrt_raycast_result* __restrict__ out,
                                                       HitMirrors mirrors) {
    constexpr bool kCulled = kMode != 0;
    __shared__ uint32_t s_fifo[kCulled ? 1 : kLeafFifo][128];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
#if USRT_TRACE_WARP_W == 8
    const uint32_t x = blockIdx.x * kTileW + (warp & 1u) * 8u + (lane & 7u);
    const uint32_t row = blockIdx.y * kTileH + (warp >> 1) * 4u + (lane >> 3);
#elif USRT_TRACE_WARP_W == 4
    const uint32_t x = blockIdx.x * kTileW + warp * 4u + (lane & 3u);
    const uint32_t row = blockIdx.y * kTileH + (lane >> 2);
#elif USRT_TRACE_WARP_W == 32
    const uint32_t x = blockIdx.x * kTileW + lane;
    const uint32_t row = blockIdx.y * kTileH + warp;
#else
    const uint32_t x = blockIdx.x * kTileW + (lane & 15u);
    const uint32_t row = blockIdx.y * kTileH + warp * 2u + (lane >> 4);
#endif
    uint32_t y, out_row;
    bool valid = x < (uint32_t)p.width;
    if (p.num_shards > 0) {                                            // ray sharding: interleaved row blocks
        const uint32_t blk = row / (uint32_t)p.block_rows, in_blk = row % (uint32_t)p.block_rows;
        y = (blk * (uint32_t)p.num_shards + (uint32_t)p.shard) * (uint32_t)p.block_rows + in_blk;
        out_row = row;
        valid = valid && row < (uint32_t)p.local_rows && y < (uint32_t)p.height;
    } else {
        y = (uint32_t)p.y0 + row;
        out_row = y;
        valid = valid && y < (uint32_t)p.y1;
    }
    const Ray ray = primary_ray(p, x, y);                              // harmless for off-frame lanes
    usrt_raycast_result h;
    if (kCulled) {
        if (!valid) return;
        h = traverse<true, kMode == 2>(scene, ray);
    } else {
        h = traverse_strict(scene, ray, valid, s_fifo);               // whole warps stay together (warp votes inside)
        if (!valid) return;
    }
    store_hit(out, (size_t)out_row * (size_t)p.width + x, h);          // frame mode: record index = y*W + x
#pragma unroll
    for (int j = 0; j < kMaxHitMirrors; ++j)                           // pinned host frame / peer GPUs' frame slots
        if (j < mirrors.count) store_hit(mirrors.ptr[j], (size_t)out_row * (size_t)p.width + x, h);
}

template <int kMode>
__global__ void __launch_bounds__(128) k_trace_rays(TraceScene scene, const float4* __restrict__ rays, uint64_t num_rays,
                                                    usrt_raycast_result* __restrict__ out) {
    constexpr bool kCulled = kMode != 0;
    __shared__ uint32_t s_fifo[kCulled ? 1 : kLeafFifo][128];
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = i < num_rays;
    const uint64_t j = valid ? i : 0;
    const float4 o = __ldg(rays + j * 2), d = __ldg(rays + j * 2 + 1);
    Ray r;
    r.ox = o.x; r.oy = o.y; r.oz = o.z;
    r.dx = d.x; r.dy = d.y; r.dz = d.z;
    r.ix = __fdiv_rn(1.0f, d.x); r.iy = __fdiv_rn(1.0f, d.y); r.iz = __fdiv_rn(1.0f, d.z);
    usrt_raycast_result h;
    if (kCulled) {
        if (!valid) return;
        h = traverse<true, kMode == 2>(scene, r);
    } else {
        h = traverse_strict(scene, r, valid, s_fifo);
        if (!valid) return;
    }
    store_hit(out, (size_t)i, h);
}

// ---- diffuse bounce rays (BASELINE config 5: "64 spp random diffuse rays") ------------------------------
// The reference casts primary rays only; the generator below is DEFINED by the test oracle's DiffuseRay
// function and restated here operation for operation: hit point on the
// primary ray, geometric normal of the hit triangle turned against the ray, plus a hashed, trig-free random
// unit vector (cosine-weighted hemisphere), seeded by (seed, pixel, sample). Pixels that hit nothing and
// degenerate triangles emit the null ray (all zeros), which can hit nothing (det = 0 in ray_triangle).
__device__ __forceinline__ uint64_t hash_u64(uint64_t x) {             // splitmix64 finaliser
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

__global__ void __launch_bounds__(256) k_diffuse_rays(PrimaryParams p, const usrt_raycast_result* __restrict__ hits,
                                                      VertexSource vertices, uint64_t seed, uint32_t s0,
                                                      uint64_t total, float4* __restrict__ rays_out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const uint64_t frame = (uint64_t)p.width * (uint64_t)p.height;
    const uint32_t sample = s0 + (uint32_t)(i / frame);
    const uint32_t pixel = (uint32_t)(i % frame);
    const uint32_t x = pixel % (uint32_t)p.width, y = pixel / (uint32_t)p.width;
    float4 o4 = make_float4(0.f, 0.f, 0.f, 0.f), d4 = o4;
    const float4 h = __ldg(reinterpret_cast<const float4*>(hits) + pixel);
    if (__float_as_uint(h.x) != 0x4EFF0000u) {                         // distance != MAX_FLOAT: the pixel hit something
        const Ray r = primary_ray(p, x, y);
        const float t = h.x;
        const float px = add(r.ox, mul(r.dx, t)), py = add(r.oy, mul(r.dy, t)), pz = add(r.oz, mul(r.dz, t));
        const float4* tri = vertices.base + (size_t)__float_as_uint(h.y) * vertices.stride;
        const float4 a = ldg_vertex(tri), b = ldg_vertex(tri + 1), c = ldg_vertex(tri + 2);
        const float e1x = sub(b.x, a.x), e1y = sub(b.y, a.y), e1z = sub(b.z, a.z);
        const float e2x = sub(c.x, a.x), e2y = sub(c.y, a.y), e2z = sub(c.z, a.z);
        float nx = sub(mul(e1y, e2z), mul(e1z, e2y));
        float ny = sub(mul(e1z, e2x), mul(e1x, e2z));
        float nz = sub(mul(e1x, e2y), mul(e1y, e2x));
        const float n2 = dot3(nx, ny, nz, nx, ny, nz);
        if (n2 > 0.0f) {
            const float nl = __fsqrt_rn(n2);
            nx = __fdiv_rn(nx, nl); ny = __fdiv_rn(ny, nl); nz = __fdiv_rn(nz, nl);
            if (dot3(nx, ny, nz, r.dx, r.dy, r.dz) > 0.0f) { nx = -nx; ny = -ny; nz = -nz; }
            float ux = nx, uy = ny, uz = nz;                           // fallback after 16 rejected draws
            const uint64_t stream = hash_u64(seed ^ hash_u64(((uint64_t)pixel << 16) | (uint64_t)(sample & 0xFFFFu)));
            for (uint32_t k = 0; k < 16u; ++k) {
                const uint64_t bits = hash_u64(stream + k);
                // three 21-bit fields -> [-1, 1): exact in fp32
                const float vx = sub(mul(__uint2float_rn((uint32_t)(bits & 0x1FFFFFu)), 9.5367431640625e-07f), 1.0f);
                const float vy = sub(mul(__uint2float_rn((uint32_t)((bits >> 21) & 0x1FFFFFu)), 9.5367431640625e-07f), 1.0f);
                const float vz = sub(mul(__uint2float_rn((uint32_t)((bits >> 42) & 0x1FFFFFu)), 9.5367431640625e-07f), 1.0f);
                const float l2 = dot3(vx, vy, vz, vx, vy, vz);
                if (l2 <= 1.0f && l2 > 1e-4f) {
                    const float l = __fsqrt_rn(l2);
                    ux = __fdiv_rn(vx, l); uy = __fdiv_rn(vy, l); uz = __fdiv_rn(vz, l);
                    break;
                }
            }
            float dx = add(nx, ux), dy = add(ny, uy), dz = add(nz, uz);
            const float d2 = dot3(dx, dy, dz, dx, dy, dz);
            if (d2 < 1e-8f) { dx = nx; dy = ny; dz = nz; }
            else { const float dl = __fsqrt_rn(d2); dx = __fdiv_rn(dx, dl); dy = __fdiv_rn(dy, dl); dz = __fdiv_rn(dz, dl); }
            o4 = make_float4(add(px, mul(nx, 0.001f)), add(py, mul(ny, 0.001f)), add(pz, mul(nz, 0.001f)), 0.f);
            d4 = make_float4(dx, dy, dz, 0.f);
        }
    }
    rays_out[i * 2] = o4;
    rays_out[i * 2 + 1] = d4;
}

__global__ void __launch_bounds__(256) k_fill_miss(float4* __restrict__ out, uint64_t count) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) out[i] = make_float4(max_float(), 0.0f, 0.0f, 0.0f);   // triangleIndex 0 == +0.0f bits
}

}  // namespace

cudaError_t launch_fill_miss(usrt_raycast_result* out, uint64_t count, cudaStream_t stream) {
    if (count == 0) return cudaSuccess;
    k_fill_miss<<<(unsigned)((count + 255) / 256), 256, 0, stream>>>(reinterpret_cast<float4*>(out), count);
    return cudaGetLastError();
}

cudaError_t launch_trace_primary(const TraceScene& scene, const PrimaryParams& p, usrt_raycast_result* out, int mode,
                                 cudaStream_t stream, const HitMirrors& mirrors) {
    const int rows = p.num_shards > 0 ? p.local_rows : p.y1 - p.y0;
    if (rows <= 0 || p.width <= 0) return cudaSuccess;
    const dim3 grid((p.width + kTileW - 1) / kTileW, (rows + kTileH - 1) / kTileH);
    if (mode == 2) k_trace_primary<2><<<grid, 128, 0, stream>>>(scene, p, out, mirrors);
    else if (mode == 1) k_trace_primary<1><<<grid, 128, 0, stream>>>(scene, p, out, mirrors);
    else k_trace_primary<0><<<grid, 128, 0, stream>>>(scene, p, out, mirrors);
    return cudaGetLastError();
}

cudaError_t launch_trace_rays(const TraceScene& scene, const float4* rays, uint64_t num_rays, usrt_raycast_result* out,
                              int mode, cudaStream_t stream) {
    if (num_rays == 0) return cudaSuccess;
    const uint32_t grid = (uint32_t)((num_rays + 127) / 128);
    if (mode == 2) k_trace_rays<2><<<grid, 128, 0, stream>>>(scene, rays, num_rays, out);
    else if (mode == 1) k_trace_rays<1><<<grid, 128, 0, stream>>>(scene, rays, num_rays, out);
    else k_trace_rays<0><<<grid, 128, 0, stream>>>(scene, rays, num_rays, out);
    return cudaGetLastError();
}

cudaError_t launch_diffuse_rays(const PrimaryParams& p, const usrt_raycast_result* hits, VertexSource vertices,
                                uint64_t seed, uint32_t s0, uint32_t s_count, float4* rays_out, cudaStream_t stream) {
    const uint64_t total = (uint64_t)p.width * (uint64_t)p.height * (uint64_t)s_count;
    if (total == 0) return cudaSuccess;
    k_diffuse_rays<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(p, hits, vertices, seed, s0, total, rays_out);
    return cudaGetLastError();
}

}  // namespace usrt
