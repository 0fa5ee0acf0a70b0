// usrt_oracle.cpp -- TEST INFRASTRUCTURE ONLY. Scalar CPU restatement ("C++ twin") of the
// UnitySimpleRaytracing hot path: Morton/AABB -> 4-pass 8-bit LSD key/value radix sort ->
// DistributeKeys -> Karras LBVH topology -> atomic-counter bottom-up AABB refit -> 64-entry
// stack traversal with the reference's slab and Moller-Trumbore tests.
//
// PINNED AGAINST THE REFERENCE'S OWN TEXT (round 2): oracle/build_ref.sh compiles BVH.compute, Raytracing.compute, the
// five Sorting/*.compute kernels and the static functions + DistributeKeys of MeshBufferContainer.cs with g++
// (oracle/_ref/libusrt_ref.so: a syntactic sed pass, shim headers for the HLSL / UnityEngine types, and a lock-step wave
// emulator -- coroutines per thread, group barriers, WavePrefixCountBits / WavePrefixSum -- for the sort kernels).
// tests/test_ref_pin.py requires this file to agree with it bit for bit: every build buffer, every intermediate of a
// sort pass (block-sorted pairs, per-block digit offsets, the count table before / after the scan) incl. one pass at the
// reference's full 512 x 1024 capacity, full primary frames (configs[0] 512x512; configs[1] 1920x1080 at 1,048,576
// triangles via tests/golden/ref_digests.json) and the shading epilogue. The reference ships no golden vectors of its
// own; its runtime self-checks are restated in tests/test_oracle.py.
// The arithmetic the HLSL leaves to the hardware (normalize -> rsqrt, 1/x -> rcp, FMA contraction, the bilinear sampler)
// is DEFINED here as IEEE fp32 per-operation, left-to-right, no contraction (build with -ffp-contract=off, no
// fast-math); the shim headers of the _ref build carry the same definitions and nothing else.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// load this library. The product (unitysimpleraytracing_b200/) never does.
//
// Each function cites the reference file:line it follows (paths relative to /root/reference).

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

namespace {

// ---- Assets/_Shaders/Constants.cginc:1-7 ------------------------------------------------------
constexpr uint32_t RADIX = 8;
constexpr uint32_t BUCKET_SIZE = 256;          // 2 ^ RADIX
constexpr uint32_t THREADS_PER_BLOCK = 1024;   // elements per LocalRadixSort group
constexpr uint32_t WARP_SIZE = 32;
// Constants.cginc:7 -- MAX_FLOAT is the INTEGER literal 0x7F7FFFFF assigned to a float
// (Raytracing.compute:49,129): (float)2139095039 == 2139095040.0f (bits 0x4EFF0000), not FLT_MAX.
const float MAX_FLOAT = (float)0x7F7FFFFF;

// ---- Constants.cginc:9-54 / SceneDataTypes.cs:4-90 ---------------------------------------------
struct AABB { float min[3]; float _dummy0; float max[3]; float _dummy1; };
struct InternalNode { uint32_t leftNode, leftNodeType, rightNode, rightNodeType, parent, index; };
struct LeafNode { uint32_t parent, index; };
struct Triangle {
    float a[3]; float _d0; float b[3]; float _d1; float c[3]; float _d2;
    float a_uv[2], b_uv[2], c_uv[2], _d3[2];
    float a_n[3]; float _d4; float b_n[3]; float _d5; float c_n[3]; float _d6;
};
// Raytracing.compute:30-35
struct RaycastResult { float distance; uint32_t triangleIndex; float uv[2]; };

static_assert(sizeof(AABB) == 32, "AABB");                 // MeshBufferContainer.cs:103
static_assert(sizeof(Triangle) == 128, "Triangle");        // MeshBufferContainer.cs:98
static_assert(sizeof(InternalNode) == 24, "InternalNode");
static_assert(sizeof(LeafNode) == 8, "LeafNode");
static_assert(sizeof(RaycastResult) == 16, "RaycastResult");

constexpr uint32_t INTERNAL_NODE = 0;  // Constants.cginc:17
constexpr uint32_t LEAF_NODE = 1;      // Constants.cginc:18

// C# Math.Min/Max(float,float) as used on finite inputs (MeshBufferContainer.cs:55-62) and HLSL
// min/max in MergeAABB (BVH.compute:152-170): plain compare-select; inputs are never NaN here.
inline float minf_sel(float a, float b) { return a < b ? a : b; }
inline float maxf_sel(float a, float b) { return a > b ? a : b; }

// ---- MeshBufferContainer.cs:32-39 ------------------------------------------------------------
inline uint32_t ExpandBits(uint32_t v) {
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}

// ---- MeshBufferContainer.cs:41-50 ------------------------------------------------------------
inline uint32_t Morton3D(float x, float y, float z) {
    x = minf_sel(maxf_sel(x * 1024.0f, 0.0f), 1023.0f);
    y = minf_sel(maxf_sel(y * 1024.0f, 0.0f), 1023.0f);
    z = minf_sel(maxf_sel(z * 1024.0f, 0.0f), 1023.0f);
    uint32_t xx = ExpandBits((uint32_t)x);   // C# (uint)float truncates toward zero
    uint32_t yy = ExpandBits((uint32_t)y);
    uint32_t zz = ExpandBits((uint32_t)z);
    return xx * 4 + yy * 2 + zz;
}

// ---- MeshBufferContainer.cs:52-71 ------------------------------------------------------------
inline void GetCentroidAndAABB(const float* a, const float* b, const float* c, float* centroid, AABB* aabb) {
    for (int k = 0; k < 3; ++k) {
        float mn = minf_sel(minf_sel(a[k], b[k]), c[k]) - 0.001f;
        float mx = maxf_sel(maxf_sel(a[k], b[k]), c[k]) + 0.001f;
        centroid[k] = (mn + mx) * 0.5f;     // centroid of the PADDED box, not the vertex mean
        aabb->min[k] = mn;
        aabb->max[k] = mx;
    }
    aabb->_dummy0 = 0.0f;  // C# struct default
    aabb->_dummy1 = 0.0f;
}

// ---- MeshBufferContainer.cs:9-15,73-83 -------------------------------------------------------
inline void NormalizeCentroid(float* c, float whole_min, float whole_max) {
    for (int k = 0; k < 3; ++k) {
        c[k] -= whole_min;
        c[k] /= (whole_max - whole_min);   // true fp32 division by 250f, not a reciprocal multiply
    }
}

// ---- BVH.compute:18-21 -----------------------------------------------------------------------
inline int clz32(uint32_t v) {
    // 31 - firstbithigh(v); firstbithigh(0) == -1 (0xFFFFFFFF) in HLSL => clz32(0) == 32
    if (v == 0) return 32;
    return __builtin_clz(v);
}

struct TreeCtx { const uint32_t* sortedMortonCodes; };

// ---- BVH.compute:23-33 -----------------------------------------------------------------------
inline int delta(const TreeCtx& t, int x, int y, int numObjects) {
    if (x >= 0 && x <= numObjects - 1 && y >= 0 && y <= numObjects - 1) {
        const uint32_t x_code = t.sortedMortonCodes[x];
        const uint32_t y_code = t.sortedMortonCodes[y];
        return clz32(x_code ^ y_code);   // "we guarantee that x_code != y_code" (DistributeKeys)
    }
    return -1;
}

inline int sign_i(int v) { return (v > 0) - (v < 0); }

// ---- BVH.compute:35-52 -----------------------------------------------------------------------
// HLSL evaluates `idx + lmax * d` with lmax:uint, d:int in 32-bit two's complement (wraps); C#
// would promote uint*int to long (SURVEY 8a trap) -- use explicit uint32 arithmetic here.
inline void DetermineRange(const TreeCtx& t, int numObjects, int idx, int* first, int* last) {
    const int d = sign_i(delta(t, idx, idx + 1, numObjects) - delta(t, idx, idx - 1, numObjects));
    const int dmin = delta(t, idx, idx - d, numObjects);
    uint32_t lmax = 2;
    while (delta(t, idx, (int)((uint32_t)idx + lmax * (uint32_t)d), numObjects) > dmin)
        lmax = lmax * 2;
    int l = 0;
    for (uint32_t s = lmax / 2; s >= 1; s /= 2) {
        if (delta(t, idx, (int)((uint32_t)idx + ((uint32_t)l + s) * (uint32_t)d), numObjects) > dmin)
            l += (int)s;
    }
    const int j = idx + l * d;
    *first = std::min(idx, j);
    *last = std::max(idx, j);
}

// ---- BVH.compute:54-92 -----------------------------------------------------------------------
inline int FindSplit(const TreeCtx& t, int first, int last) {
    const uint32_t firstCode = t.sortedMortonCodes[first];
    const uint32_t lastCode = t.sortedMortonCodes[last];
    if (firstCode == lastCode) return (first + last) >> 1;
    const int commonPrefix = clz32(firstCode ^ lastCode);
    int split = first;
    int step = last - first;
    do {
        step = (step + 1) >> 1;
        const int newSplit = split + step;
        if (newSplit < last) {
            const uint32_t splitCode = t.sortedMortonCodes[newSplit];
            const int splitPrefix = clz32(firstCode ^ splitCode);
            if (splitPrefix > commonPrefix) split = newSplit;
        }
    } while (step > 1);
    return split;
}

// ---- BVH.compute:152-170 ---------------------------------------------------------------------
inline AABB MergeAABB(const AABB& l, const AABB& r) {
    AABB ret;
    for (int k = 0; k < 3; ++k) {
        ret.min[k] = minf_sel(l.min[k], r.min[k]);
        ret.max[k] = maxf_sel(l.max[k], r.max[k]);
    }
    ret._dummy0 = 0;
    ret._dummy1 = 0;
    return ret;
}

// ---- Raytracing.compute:23-28 ----------------------------------------------------------------
struct Ray { float origin[3]; float dir[3]; float inv_dir[3]; };

// canonical expansions (SURVEY 8a "transliteration traps"): left-to-right, no FMA
inline float dot3(const float* a, const float* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
inline void cross3(const float* a, const float* b, float* o) {
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}

// ---- Raytracing.compute:37-73 ----------------------------------------------------------------
inline RaycastResult RayTriangleIntersection(const float* orig, const float* dir, const float* v0,
                                             const float* v1, const float* v2) {
    RaycastResult result;
    result.triangleIndex = 0; result.uv[0] = 0; result.uv[1] = 0;  // HLSL leaves these undefined
    float e1[3], e2[3], pvec[3], tvec[3], qvec[3];
    for (int k = 0; k < 3; ++k) { e1[k] = v1[k] - v0[k]; e2[k] = v2[k] - v0[k]; }
    cross3(dir, e2, pvec);
    const float det = dot3(e1, pvec);
    if (det < 1e-8f && det > -1e-8f) { result.distance = MAX_FLOAT; return result; }
    const float inv_det = 1.0f / det;
    for (int k = 0; k < 3; ++k) tvec[k] = orig[k] - v0[k];
    const float u = dot3(tvec, pvec) * inv_det;
    if (u < 0 || u > 1) { result.distance = MAX_FLOAT; return result; }
    cross3(tvec, e1, qvec);
    const float v = dot3(dir, qvec) * inv_det;
    if (v < 0 || u + v > 1) { result.distance = MAX_FLOAT; return result; }
    result.distance = dot3(e2, qvec) * inv_det;   // no t>0 test: negative t is accepted
    result.uv[0] = u; result.uv[1] = v;
    return result;
}

// ---- Raytracing.compute:75-87 (CPU twin: _debugRayBoxIntersectionTester.cs:33-45) --------------
// HLSL min/max return the non-NaN operand (0*inf cases) == C fminf/fmaxf.
inline bool RayBoxIntersection(const AABB& b, const Ray& r) {
    float tmin1[3], tmax1[3];
    for (int k = 0; k < 3; ++k) {
        const float t1 = (b.min[k] - r.origin[k]) * r.inv_dir[k];
        const float t2 = (b.max[k] - r.origin[k]) * r.inv_dir[k];
        tmin1[k] = fminf(t1, t2);
        tmax1[k] = fmaxf(t1, t2);
    }
    const float tmin = fmaxf(tmin1[0], fmaxf(tmin1[1], tmin1[2]));
    const float tmax = fminf(tmax1[0], fminf(tmax1[1], tmax1[2]));
    return tmax > tmin && tmax > 0;
}

struct Scene {
    const uint32_t* sortedTriangleIndices;
    const AABB* triangleAABB;
    const InternalNode* internalNodes;
    const LeafNode* leafNodes;
    const AABB* bvhData;
    const Triangle* triangleData;
};

struct TraceCounters { uint64_t boxTests, triBoxTests, triTests, maxStack; };

// ---- Raytracing.compute:89-103 ---------------------------------------------------------------
inline RaycastResult CheckTriangle(const Scene& s, uint32_t triangleIndex, const Ray& ray, RaycastResult result,
                                   TraceCounters* c) {
    if (c) c->triBoxTests++;
    if (RayBoxIntersection(s.triangleAABB[triangleIndex], ray)) {
        if (c) c->triTests++;
        const Triangle& t = s.triangleData[triangleIndex];
        RaycastResult newResult = RayTriangleIntersection(ray.origin, ray.dir, t.a, t.b, t.c);
        if (newResult.distance < result.distance) {   // strict: first visited wins ties
            newResult.triangleIndex = triangleIndex;
            return newResult;
        }
        return result;
    }
    return result;
}

// ---- Raytracing.compute:128-176 (the traversal loop, given a ray) --------------------------------
inline RaycastResult TraverseRay(const Scene& s, const Ray& ray, TraceCounters* c) {
    RaycastResult result;
    result.distance = MAX_FLOAT;
    result.triangleIndex = 0;
    result.uv[0] = 0; result.uv[1] = 0;

    uint32_t stack[64];
    uint32_t currentStackIndex = 0;
    stack[currentStackIndex] = 0;
    currentStackIndex = 1;

    while (currentStackIndex != 0) {
        currentStackIndex--;
        const uint32_t index = stack[currentStackIndex];
        if (c) c->boxTests++;
        if (!RayBoxIntersection(s.bvhData[index], ray)) continue;

        const uint32_t leftIndex = s.internalNodes[index].leftNode;
        const uint32_t leftType = s.internalNodes[index].leftNodeType;
        if (leftType == INTERNAL_NODE) {
            stack[currentStackIndex] = leftIndex;
            currentStackIndex++;
        } else {
            const uint32_t triangleIndex = s.sortedTriangleIndices[s.leafNodes[leftIndex].index];
            result = CheckTriangle(s, triangleIndex, ray, result, c);
        }
        const uint32_t rightIndex = s.internalNodes[index].rightNode;
        const uint32_t rightType = s.internalNodes[index].rightNodeType;
        if (rightType == INTERNAL_NODE) {
            stack[currentStackIndex] = rightIndex;
            currentStackIndex++;
        } else {
            const uint32_t triangleIndex = s.sortedTriangleIndices[s.leafNodes[rightIndex].index];
            result = CheckTriangle(s, triangleIndex, ray, result, c);
        }
        if (c && currentStackIndex > c->maxStack) c->maxStack = currentStackIndex;
    }
    return result;
}

// ---- Raytracing.compute:108-126 + RaytracingMeshDrawer.cs:78-81 ----------------------------------
// m is row-major: m[r*4+c]; mul(M, v) = row . vector, left to right, w term included.
inline Ray PrimaryRay(uint32_t x, uint32_t y, int screenWidth, int screenHeight, float near, float cameraFov,
                      const float* m) {
    const float fov = cameraFov;                 // = tan(fovDeg * Deg2Rad / 2), precomputed on host
    const float height = 2 * near * fov;
    const float width = (float)screenWidth * height / (float)screenHeight;
    float o4[4] = {0, 0, 0, 1};
    float d4[4] = {
        -width / 2 + width / (float)screenWidth * ((float)x + 0.5f),
        -height / 2 + height / (float)screenHeight * ((float)y + 0.5f),
        -near, 0};
    Ray ray;
    float dir[3];
    for (int r = 0; r < 3; ++r) {
        ray.origin[r] = m[r * 4 + 0] * o4[0] + m[r * 4 + 1] * o4[1] + m[r * 4 + 2] * o4[2] + m[r * 4 + 3] * o4[3];
        dir[r] = m[r * 4 + 0] * d4[0] + m[r * 4 + 1] * d4[1] + m[r * 4 + 2] * d4[2] + m[r * 4 + 3] * d4[3];
    }
    const float len = sqrtf(dot3(dir, dir));     // normalize(v) = v / sqrt(dot(v,v)), per component
    for (int k = 0; k < 3; ++k) {
        ray.dir[k] = dir[k] / len;
        ray.inv_dir[k] = 1.0f / ray.dir[k];        // +-inf allowed
    }
    return ray;
}

inline Ray BufferRay(const float* origin_dir8) {
    // ray-buffer ABI: 2 x float4 (origin.xyz,_ ; dir.xyz,_) ; dir used as given; inv_dir = 1/dir
    Ray ray;
    for (int k = 0; k < 3; ++k) {
        ray.origin[k] = origin_dir8[k];
        ray.dir[k] = origin_dir8[4 + k];
        ray.inv_dir[k] = 1.0f / ray.dir[k];
    }
    return ray;
}

template <class F>
void parallel_for(uint64_t n, int threads, F f) {
    if (threads <= 1 || n < 1024) { f(0, n, 0); return; }
    std::vector<std::thread> pool;
    std::atomic<uint64_t> next(0);
    const uint64_t chunk = 4096;
    for (int t = 0; t < threads; ++t) {
        pool.emplace_back([&, t]() {
            for (;;) {
                uint64_t b = next.fetch_add(chunk);
                if (b >= n) break;
                f(b, std::min(n, b + chunk), t);
            }
        });
    }
    for (auto& th : pool) th.join();
}

}  // namespace

extern "C" {

float usrt_oracle_max_float() { return MAX_FLOAT; }

// MeshBufferContainer.cs:123-146 (positions only; uv/normal packing is mesh ingest, not this path).
// whole = +-125 (MeshBufferContainer.cs:9-15) unless the caller overrides it.
// The same loop with a per-axis Whole box, and the scene box the reference leaves as a TODO
// ("reduce scene data for finding AABB scene in runtime", MeshBufferContainer.cs:7): per-axis min / max over all
// vertices; an axis on which the mesh is flat gets max = min + 1 so that NormalizeCentroid never divides by 0.
// Opt-in, outside reference parity (the reference's keys come from the fixed +-125 cube).
void usrt_oracle_scene_box(const Triangle* tris, uint32_t n, float* box_min, float* box_max) {
    for (int k = 0; k < 3; ++k) { box_min[k] = INFINITY; box_max[k] = -INFINITY; }
    for (uint32_t i = 0; i < n; i++)
        for (const float* v : {tris[i].a, tris[i].b, tris[i].c})
            for (int k = 0; k < 3; ++k) {
                box_min[k] = minf_sel(box_min[k], v[k]);
                box_max[k] = maxf_sel(box_max[k], v[k]);
            }
    for (int k = 0; k < 3; ++k)
        if (!(box_max[k] > box_min[k])) box_max[k] = box_min[k] + 1.0f;
}

void usrt_oracle_morton_box(const Triangle* tris, uint32_t n, const float* box_min, const float* box_max,
                            uint32_t* keys, uint32_t* values, AABB* aabbs) {
    for (uint32_t i = 0; i < n; i++) {
        float centroid[3];
        AABB aabb;
        GetCentroidAndAABB(tris[i].a, tris[i].b, tris[i].c, centroid, &aabb);
        for (int k = 0; k < 3; ++k) {                 // NormalizeCentroid (:73-83) with Whole.min/max per axis
            centroid[k] -= box_min[k];
            centroid[k] /= (box_max[k] - box_min[k]);
        }
        keys[i] = Morton3D(centroid[0], centroid[1], centroid[2]);
        values[i] = i;
        aabbs[i] = aabb;
    }
}

void usrt_oracle_morton(const Triangle* tris, uint32_t n, float whole_min, float whole_max,
                        uint32_t* keys, uint32_t* values, AABB* aabbs) {
    for (uint32_t i = 0; i < n; i++) {
        float centroid[3];
        AABB aabb;
        GetCentroidAndAABB(tris[i].a, tris[i].b, tris[i].c, centroid, &aabb);
        NormalizeCentroid(centroid, whole_min, whole_max);
        keys[i] = Morton3D(centroid[0], centroid[1], centroid[2]);
        values[i] = i;
        aabbs[i] = aabb;
    }
}

// One reference sort pass over `count` elements (count must be a multiple of 1024; caller pads with
// 0xFFFFFFFF like MeshBufferContainer.cs:108-109). Emits every intermediate the reference validators
// read back (ComputeBufferSorter.cs:128-134):
//   LocalRadixSort.compute:53-134  -> sortedBlocksKeys/Values, offsets[block*256+digit],
//                                     sizesBefore[digit*numBlocks+block]
//   Scan.compute:15-96             -> sizesAfter = exclusive prefix sum of sizesBefore (linear order)
//   GlobalRadixSort.compute:20-40  -> keys/values rewritten in place (ComputeBufferSorter.cs:87-88)
void usrt_oracle_sort_pass(uint32_t* keys, uint32_t* values, uint32_t count, int bitOffset,
                           uint32_t* sortedBlocksKeys, uint32_t* sortedBlocksValues, uint32_t* offsets,
                           uint32_t* sizesBefore, uint32_t* sizesAfter) {
    const uint32_t numBlocks = count / THREADS_PER_BLOCK;
    std::vector<uint32_t> sortTile(THREADS_PER_BLOCK), valuesTile(THREADS_PER_BLOCK);
    std::vector<uint32_t> nk(THREADS_PER_BLOCK), nv(THREADS_PER_BLOCK);
    for (uint32_t groupId = 0; groupId < numBlocks; ++groupId) {
        // LocalRadixSort.compute:59-60
        for (uint32_t t = 0; t < THREADS_PER_BLOCK; ++t) {
            sortTile[t] = keys[groupId * THREADS_PER_BLOCK + t];
            valuesTile[t] = values[groupId * THREADS_PER_BLOCK + t];
        }
        // LocalRadixSort.compute:64-91 -- eight stable one-bit splits (IntraBlockScan :29-51 is an
        // exclusive count of set predicates before each thread)
        for (uint32_t shift = bitOffset; shift < (uint32_t)bitOffset + RADIX; shift++) {
            uint32_t trueTotal = 0;
            for (uint32_t t = 0; t < THREADS_PER_BLOCK; ++t) trueTotal += (sortTile[t] >> shift) & 1;
            const uint32_t falseTotal = THREADS_PER_BLOCK - trueTotal;   // :81-84
            uint32_t trueBefore = 0;
            for (uint32_t t = 0; t < THREADS_PER_BLOCK; ++t) {
                const uint32_t key = sortTile[t], value = valuesTile[t];
                const bool pred = (key >> shift) & 1;
                const uint32_t dst = pred ? trueBefore + falseTotal : t - trueBefore;   // :87-88
                nk[dst] = key; nv[dst] = value;
                trueBefore += pred;
            }
            sortTile.swap(nk); valuesTile.swap(nv);
        }
        // :99-100
        for (uint32_t t = 0; t < THREADS_PER_BLOCK; ++t) {
            sortedBlocksKeys[groupId * THREADS_PER_BLOCK + t] = sortTile[t];
            sortedBlocksValues[groupId * THREADS_PER_BLOCK + t] = valuesTile[t];
        }
        // :102-133 -- first position and run length of each digit in the block
        uint32_t offsetsTile[BUCKET_SIZE] = {0}, sizesTile[BUCKET_SIZE] = {0};
        auto radixOf = [&](uint32_t t) { return (sortTile[t] >> bitOffset) & (BUCKET_SIZE - 1); };
        for (uint32_t t = 1; t < THREADS_PER_BLOCK; ++t)
            if (radixOf(t - 1) != radixOf(t)) offsetsTile[radixOf(t)] = t;
        for (uint32_t t = 1; t < THREADS_PER_BLOCK; ++t)
            if (radixOf(t - 1) != radixOf(t)) { uint32_t r = radixOf(t - 1); sizesTile[r] = t - offsetsTile[r]; }
        { uint32_t r = radixOf(THREADS_PER_BLOCK - 1); sizesTile[r] = THREADS_PER_BLOCK - offsetsTile[r]; }
        for (uint32_t d = 0; d < BUCKET_SIZE; ++d) {
            offsets[groupId * BUCKET_SIZE + d] = offsetsTile[d];
            sizesBefore[groupId + d * numBlocks] = sizesTile[d];   // digit-major (:132, BLOCK_SIZE -> numBlocks)
        }
    }
    // Scan.compute:15-96 -- PreScan + BlockSum + GlobalScan == exclusive prefix over the table
    {
        uint32_t run = 0;
        const uint64_t total = (uint64_t)BUCKET_SIZE * numBlocks;
        for (uint64_t i = 0; i < total; ++i) { sizesAfter[i] = run; run += sizesBefore[i]; }
    }
    // GlobalRadixSort.compute:20-40
    for (uint32_t groupId = 0; groupId < numBlocks; ++groupId) {
        for (uint32_t t = 0; t < THREADS_PER_BLOCK; ++t) {
            const uint32_t key = sortedBlocksKeys[groupId * THREADS_PER_BLOCK + t];
            const uint32_t value = sortedBlocksValues[groupId * THREADS_PER_BLOCK + t];
            const uint32_t radix = (key >> bitOffset) & (BUCKET_SIZE - 1);
            const uint32_t indexOutput = sizesAfter[groupId + radix * numBlocks] + t - offsets[groupId * BUCKET_SIZE + radix];
            keys[indexOutput] = key;
            values[indexOutput] = value;
        }
    }
}

// ComputeBufferSorter.cs:100-126 -- bitOffset = 0, 8, 16, 24 over all 32 bits. `count` is arbitrary:
// the tail block is padded with 0xFFFFFFFF pairs exactly as the fixed-capacity buffers are
// (MeshBufferContainer.cs:108-109); padding sorts to the very end (stable) and is dropped again.
void usrt_oracle_sort(uint32_t* keys, uint32_t* values, uint64_t count) {
    if (count == 0) return;
    const uint64_t padded = (count + THREADS_PER_BLOCK - 1) / THREADS_PER_BLOCK * THREADS_PER_BLOCK;
    std::vector<uint32_t> k(padded, 0xFFFFFFFFu), v(padded, 0xFFFFFFFFu);
    std::memcpy(k.data(), keys, count * 4);
    std::memcpy(v.data(), values, count * 4);
    const uint64_t nb = padded / THREADS_PER_BLOCK;
    std::vector<uint32_t> sbk(padded), sbv(padded), off(nb * BUCKET_SIZE), sb(nb * BUCKET_SIZE), sa(nb * BUCKET_SIZE);
    for (int bitOffset = 0; bitOffset < 32; bitOffset += RADIX)
        usrt_oracle_sort_pass(k.data(), v.data(), (uint32_t)padded, bitOffset, sbk.data(), sbv.data(), off.data(),
                              sb.data(), sa.data());
    std::memcpy(keys, k.data(), count * 4);
    std::memcpy(values, v.data(), count * 4);
}

// Cross-check used by the tests: the net contract of Sort() is std::stable_sort by the 32-bit key.
void usrt_oracle_stable_sort(uint32_t* keys, uint32_t* values, uint64_t count) {
    std::vector<uint64_t> idx(count);
    for (uint64_t i = 0; i < count; ++i) idx[i] = i;
    std::stable_sort(idx.begin(), idx.end(), [&](uint64_t a, uint64_t b) { return keys[a] < keys[b]; });
    std::vector<uint32_t> k(count), v(count);
    for (uint64_t i = 0; i < count; ++i) { k[i] = keys[idx[i]]; v[i] = values[idx[i]]; }
    std::memcpy(keys, k.data(), count * 4);
    std::memcpy(values, v.data(), count * 4);
}

// MeshBufferContainer.cs:154-169 -- unchecked uint32 arithmetic; Math.Max(uint, 1) is unsigned.
void usrt_oracle_distribute_keys(uint32_t* keys, uint32_t trianglesLength) {
    if (trianglesLength == 0) return;
    uint32_t newCurrentValue = 0;
    uint32_t oldCurrentValue = keys[0];
    keys[0] = newCurrentValue;
    for (uint32_t i = 1; i < trianglesLength; i++) {
        newCurrentValue += std::max<uint32_t>(keys[i] - oldCurrentValue, 1u);
        oldCurrentValue = keys[i];
        keys[i] = newCurrentValue;
    }
}

// BVH.compute:94-149 -- one "thread" per internal node; buffers must be pre-filled with NullLeaf
// (all 0xFFFFFFFF, MeshBufferContainer.cs:114-115) by the caller; root.parent is never written.
void usrt_oracle_construct_tree(const uint32_t* sortedMortonCodes, uint32_t trianglesCount,
                                InternalNode* internalNodes, LeafNode* leafNodes) {
    TreeCtx t{sortedMortonCodes};
    if (trianglesCount < 2) return;   // BVH.compute:101 with n<2 builds nothing
    for (uint32_t threadId = 0; threadId < trianglesCount - 1; ++threadId) {
        int first, last;
        DetermineRange(t, (int)trianglesCount, (int)threadId, &first, &last);
        const int split = FindSplit(t, first, last);
        internalNodes[threadId].index = threadId;
        if (split == first) {
            leafNodes[split] = LeafNode{threadId, (uint32_t)split};
            internalNodes[threadId].leftNode = split;
            internalNodes[threadId].leftNodeType = LEAF_NODE;
        } else {
            internalNodes[split].parent = threadId;
            internalNodes[threadId].leftNode = split;
            internalNodes[threadId].leftNodeType = INTERNAL_NODE;
        }
        if (split + 1 == last) {
            leafNodes[split + 1] = LeafNode{threadId, (uint32_t)(split + 1)};
            internalNodes[threadId].rightNode = split + 1;
            internalNodes[threadId].rightNodeType = LEAF_NODE;
        } else {
            internalNodes[split + 1].parent = threadId;
            internalNodes[threadId].rightNode = split + 1;
            internalNodes[threadId].rightNodeType = INTERNAL_NODE;
        }
    }
}

// ---- SURVEY 8(f)-4 key variants: oracle twins ---------------------------------------------------------------------
// The reference has neither (its sorter is generic over uint / ulong keys, ComputeBufferSorter.cs:179-191, and
// _debug/debugShader.compute probes uint64_t, but no 64-bit Morton or tree code ships): these restate the SAME
// functions above with 64-bit keys, so they are defined here, not pinned by the reference's text.

// Morton3D with 21 bits per axis (the classic 64-bit bit spread), same clamp-and-truncate quantisation as :43-48
static inline uint64_t ExpandBits64(uint64_t v) {
    v &= 0x1FFFFFull;
    v = (v | (v << 32)) & 0x001F00000000FFFFull;
    v = (v | (v << 16)) & 0x001F0000FF0000FFull;
    v = (v | (v << 8)) & 0x100F00F00F00F00Full;
    v = (v | (v << 4)) & 0x10C30C30C30C30C3ull;
    v = (v | (v << 2)) & 0x1249249249249249ull;
    return v;
}
static inline uint64_t Morton3D64(float x, float y, float z) {
    x = minf_sel(maxf_sel(x * 2097152.0f, 0.0f), 2097151.0f);
    y = minf_sel(maxf_sel(y * 2097152.0f, 0.0f), 2097151.0f);
    z = minf_sel(maxf_sel(z * 2097152.0f, 0.0f), 2097151.0f);
    return ExpandBits64((uint32_t)x) * 4 + ExpandBits64((uint32_t)y) * 2 + ExpandBits64((uint32_t)z);
}

void usrt_oracle_morton64(const Triangle* tris, uint32_t n, float whole_min, float whole_max, uint64_t* keys, uint32_t* values,
                          AABB* aabbs) {
    for (uint32_t i = 0; i < n; i++) {
        float centroid[3];
        GetCentroidAndAABB(tris[i].a, tris[i].b, tris[i].c, centroid, &aabbs[i]);
        NormalizeCentroid(centroid, whole_min, whole_max);
        keys[i] = Morton3D64(centroid[0], centroid[1], centroid[2]);
        values[i] = i;
    }
}

// ComputeBufferSorter.Sort() over ulong keys: bitOffset = 0, 8, ..., 56; every pass a stable split by one 8-bit digit
// (GetRadix :185-188). The block / scan / scatter structure of one pass is restated in usrt_oracle_sort_pass above;
// its net effect -- a stable counting sort by that digit -- is all a wider key changes, so it is written directly.
void usrt_oracle_sort64(uint64_t* keys, uint32_t* values, uint64_t count) {
    std::vector<uint64_t> k2(count);
    std::vector<uint32_t> v2(count);
    uint64_t* ka = keys; uint64_t* kb = k2.data();
    uint32_t* va = values; uint32_t* vb = v2.data();
    for (int bitOffset = 0; bitOffset < 64; bitOffset += RADIX) {
        uint64_t start[BUCKET_SIZE + 1] = {0};
        for (uint64_t i = 0; i < count; ++i) start[((ka[i] >> bitOffset) & (BUCKET_SIZE - 1)) + 1]++;
        for (uint32_t d = 0; d < BUCKET_SIZE; ++d) start[d + 1] += start[d];
        for (uint64_t i = 0; i < count; ++i) {
            const uint64_t dst = start[(ka[i] >> bitOffset) & (BUCKET_SIZE - 1)]++;
            kb[dst] = ka[i]; vb[dst] = va[i];
        }
        std::swap(ka, kb); std::swap(va, vb);
    }   // 8 passes: the result is back in (keys, values)
}

void usrt_oracle_stable_sort64(uint64_t* keys, uint32_t* values, uint64_t count) {
    std::vector<uint64_t> idx(count);
    for (uint64_t i = 0; i < count; ++i) idx[i] = i;
    std::stable_sort(idx.begin(), idx.end(), [&](uint64_t a, uint64_t b) { return keys[a] < keys[b]; });
    std::vector<uint64_t> k(count); std::vector<uint32_t> v(count);
    for (uint64_t i = 0; i < count; ++i) { k[i] = keys[idx[i]]; v[i] = values[idx[i]]; }
    std::memcpy(keys, k.data(), count * 8);
    std::memcpy(values, v.data(), count * 4);
}

// MeshBufferContainer.cs:154-169 in unchecked ulong arithmetic
void usrt_oracle_distribute_keys64(uint64_t* keys, uint32_t trianglesLength) {
    if (trianglesLength == 0) return;
    uint64_t newCurrentValue = 0;
    uint64_t oldCurrentValue = keys[0];
    keys[0] = newCurrentValue;
    for (uint32_t i = 1; i < trianglesLength; i++) {
        newCurrentValue += std::max<uint64_t>(keys[i] - oldCurrentValue, 1ull);
        oldCurrentValue = keys[i];
        keys[i] = newCurrentValue;
    }
}

// BVH.compute:18-149 over 64-bit keys: clz64 for clz32, everything else word for word. Also the tree of the index
// tie-break variant: the caller passes (code << 32 | sorted position), Karras 2012's augmented key.
static inline int clz64(uint64_t v) { return v == 0 ? 64 : __builtin_clzll(v); }
static inline int delta64(const uint64_t* codes, int x, int y, int numObjects) {
    if (x >= 0 && x <= numObjects - 1 && y >= 0 && y <= numObjects - 1) return clz64(codes[x] ^ codes[y]);
    return -1;
}
void usrt_oracle_construct_tree64(const uint64_t* codes, uint32_t trianglesCount, InternalNode* internalNodes, LeafNode* leafNodes) {
    if (trianglesCount < 2) return;
    const int numObjects = (int)trianglesCount;
    for (uint32_t threadId = 0; threadId < trianglesCount - 1; ++threadId) {
        const int idx = (int)threadId;
        // DetermineRange (:35-52)
        const int d = sign_i(delta64(codes, idx, idx + 1, numObjects) - delta64(codes, idx, idx - 1, numObjects));
        const int dmin = delta64(codes, idx, idx - d, numObjects);
        uint32_t lmax = 2;
        while (delta64(codes, idx, (int)((uint32_t)idx + lmax * (uint32_t)d), numObjects) > dmin) lmax = lmax * 2;
        int l = 0;
        for (uint32_t t = lmax / 2; t >= 1; t /= 2)
            if (delta64(codes, idx, (int)((uint32_t)idx + ((uint32_t)l + t) * (uint32_t)d), numObjects) > dmin) l += (int)t;
        const int j = idx + l * d;
        const int first = std::min(idx, j), last = std::max(idx, j);
        // FindSplit (:54-92)
        int split;
        const uint64_t firstCode = codes[first], lastCode = codes[last];
        if (firstCode == lastCode) {
            split = (first + last) >> 1;
        } else {
            const int commonPrefix = clz64(firstCode ^ lastCode);
            split = first;
            int step = last - first;
            do {
                step = (step + 1) >> 1;
                const int newSplit = split + step;
                if (newSplit < last && clz64(firstCode ^ codes[newSplit]) > commonPrefix) split = newSplit;
            } while (step > 1);
        }
        // TreeConstructor (:111-147)
        internalNodes[threadId].index = threadId;
        if (split == first) {
            leafNodes[split] = LeafNode{threadId, (uint32_t)split};
            internalNodes[threadId].leftNode = split; internalNodes[threadId].leftNodeType = LEAF_NODE;
        } else {
            internalNodes[split].parent = threadId;
            internalNodes[threadId].leftNode = split; internalNodes[threadId].leftNodeType = INTERNAL_NODE;
        }
        if (split + 1 == last) {
            leafNodes[split + 1] = LeafNode{threadId, (uint32_t)(split + 1)};
            internalNodes[threadId].rightNode = split + 1; internalNodes[threadId].rightNodeType = LEAF_NODE;
        } else {
            internalNodes[split + 1].parent = threadId;
            internalNodes[threadId].rightNode = split + 1; internalNodes[threadId].rightNodeType = INTERNAL_NODE;
        }
    }
}

// BVH.compute:172-220 -- leaves climb; the first arrival at a node stops, the second merges. Run
// serially the outcome is identical to any parallel schedule (min/max are exact and the second
// arrival always sees both children complete). Counter buffer = BVHConstructor.cs:41 (zeroed).
void usrt_oracle_construct_bvh(uint32_t trianglesCount, const uint32_t* sortedTriangleIndices,
                               const AABB* triangleAABB, const InternalNode* internalNodes,
                               const LeafNode* leafNodes, AABB* BVHData) {
    std::vector<uint32_t> atomicsData(trianglesCount, 0);
    for (uint32_t threadId = 0; threadId < trianglesCount; ++threadId) {
        uint32_t parent = leafNodes[threadId].parent;
        while (parent != 0xFFFFFFFFu) {
            uint32_t old = atomicsData[parent];
            if (old == 0) atomicsData[parent] = 1;   // InterlockedCompareExchange(.., 0, 1, old)
            if (old == 0) break;
            const uint32_t leftId = internalNodes[parent].leftNode;
            const uint32_t leftType = internalNodes[parent].leftNodeType;
            const uint32_t rightId = internalNodes[parent].rightNode;
            const uint32_t rightType = internalNodes[parent].rightNodeType;
            const AABB leftAABB = (leftType == INTERNAL_NODE) ? BVHData[leftId] : triangleAABB[sortedTriangleIndices[leftId]];
            const AABB rightAABB = (rightType == INTERNAL_NODE) ? BVHData[rightId] : triangleAABB[sortedTriangleIndices[rightId]];
            BVHData[parent] = MergeAABB(leftAABB, rightAABB);
            parent = internalNodes[parent].parent;
        }
    }
}

// Raytracing.compute:105-176 for every pixel of a W x H frame; hit record index = y*W + x, row 0 is
// the most negative camera-space y. counters (optional, 4 x u64): node box tests, triangle box
// tests, triangle tests, max stack depth -- summed/maxed over rays.
void usrt_oracle_trace_primary(const uint32_t* sortedTriangleIndices, const AABB* triangleAABB,
                               const InternalNode* internalNodes, const LeafNode* leafNodes, const AABB* bvhData,
                               const Triangle* triangleData, int screenWidth, int screenHeight, float near,
                               float cameraFov, const float* cameraToWorld, uint32_t y0, uint32_t y1,
                               RaycastResult* out, int threads, uint64_t* counters) {
    Scene s{sortedTriangleIndices, triangleAABB, internalNodes, leafNodes, bvhData, triangleData};
    const uint64_t W = (uint64_t)screenWidth;
    const uint64_t n = (uint64_t)(y1 - y0) * W;
    std::vector<TraceCounters> tc(std::max(threads, 1), TraceCounters{0, 0, 0, 0});
    parallel_for(n, threads, [&](uint64_t b, uint64_t e, int t) {
        for (uint64_t i = b; i < e; ++i) {
            const uint32_t y = y0 + (uint32_t)(i / W), x = (uint32_t)(i % W);
            Ray ray = PrimaryRay(x, y, screenWidth, screenHeight, near, cameraFov, cameraToWorld);
            out[(uint64_t)y * W + x] = TraverseRay(s, ray, counters ? &tc[t] : nullptr);
        }
    });
    if (counters) {
        counters[0] = counters[1] = counters[2] = counters[3] = 0;
        for (auto& c : tc) {
            counters[0] += c.boxTests; counters[1] += c.triBoxTests; counters[2] += c.triTests;
            counters[3] = std::max<uint64_t>(counters[3], c.maxStack);
        }
    }
}

// Same traversal for caller-supplied rays (2 x float4 per ray: origin.xyz_, dir.xyz_).
void usrt_oracle_trace_rays(const uint32_t* sortedTriangleIndices, const AABB* triangleAABB,
                            const InternalNode* internalNodes, const LeafNode* leafNodes, const AABB* bvhData,
                            const Triangle* triangleData, const float* rays, uint64_t numRays, RaycastResult* out,
                            int threads, uint64_t* counters) {
    Scene s{sortedTriangleIndices, triangleAABB, internalNodes, leafNodes, bvhData, triangleData};
    std::vector<TraceCounters> tc(std::max(threads, 1), TraceCounters{0, 0, 0, 0});
    parallel_for(numRays, threads, [&](uint64_t b, uint64_t e, int t) {
        for (uint64_t i = b; i < e; ++i) {
            Ray ray = BufferRay(rays + i * 8);
            out[i] = TraverseRay(s, ray, counters ? &tc[t] : nullptr);
        }
    });
    if (counters) {
        counters[0] = counters[1] = counters[2] = counters[3] = 0;
        for (auto& c : tc) {
            counters[0] += c.boxTests; counters[1] += c.triBoxTests; counters[2] += c.triTests;
            counters[3] = std::max<uint64_t>(counters[3], c.maxStack);
        }
    }
}

// Per-ray node-visit counts of a primary frame (analysis aid: warp-tile load balance of the GPU kernel).
void usrt_oracle_primary_visit_counts(const uint32_t* sortedTriangleIndices, const AABB* triangleAABB,
                                      const InternalNode* internalNodes, const LeafNode* leafNodes, const AABB* bvhData,
                                      const Triangle* triangleData, int screenWidth, int screenHeight, float near,
                                      float cameraFov, const float* cameraToWorld, uint32_t* visits, int threads) {
    Scene s{sortedTriangleIndices, triangleAABB, internalNodes, leafNodes, bvhData, triangleData};
    const uint64_t W = (uint64_t)screenWidth, n = W * (uint64_t)screenHeight;
    parallel_for(n, threads, [&](uint64_t b, uint64_t e, int) {
        for (uint64_t i = b; i < e; ++i) {
            TraceCounters c{0, 0, 0, 0};
            Ray ray = PrimaryRay((uint32_t)(i % W), (uint32_t)(i / W), screenWidth, screenHeight, near, cameraFov, cameraToWorld);
            TraverseRay(s, ray, &c);
            visits[i] = (uint32_t)c.boxTests;
        }
    });
}

// Primary-ray generation only (for tests of K6a and for feeding trace_rays with identical rays).
void usrt_oracle_primary_rays(int screenWidth, int screenHeight, float near, float cameraFov,
                              const float* cameraToWorld, float* rays_out /* W*H*8 floats */) {
    for (int y = 0; y < screenHeight; ++y)
        for (int x = 0; x < screenWidth; ++x) {
            Ray r = PrimaryRay((uint32_t)x, (uint32_t)y, screenWidth, screenHeight, near, cameraFov, cameraToWorld);
            float* o = rays_out + ((uint64_t)y * screenWidth + x) * 8;
            for (int k = 0; k < 3; ++k) { o[k] = r.origin[k]; o[4 + k] = r.dir[k]; }
            o[3] = 0; o[7] = 0;
        }
}

// Brute force closest hit over triangles in a caller-given visiting order (oracle self-check v of
// SURVEY 8c): same CheckTriangle sequence, no BVH.
void usrt_oracle_brute_force(const AABB* triangleAABB, const Triangle* triangleData, const uint32_t* order,
                             uint32_t numTriangles, const float* rays, uint64_t numRays, RaycastResult* out, int threads) {
    Scene s{nullptr, triangleAABB, nullptr, nullptr, nullptr, triangleData};
    parallel_for(numRays, threads, [&](uint64_t b, uint64_t e, int) {
        for (uint64_t i = b; i < e; ++i) {
            Ray ray = BufferRay(rays + i * 8);
            RaycastResult result;
            result.distance = MAX_FLOAT; result.triangleIndex = 0; result.uv[0] = 0; result.uv[1] = 0;
            for (uint32_t j = 0; j < numTriangles; ++j)
                result = CheckTriangle(s, order ? order[j] : j, ray, result, nullptr);
            out[i] = result;
        }
    });
}

// ---- shading epilogue, Raytracing.compute:178-184 --------------------------------
// fp32 -> fp16 bits, round to nearest even (the RGBA16F render target, RaytracingMeshDrawer.cs:56).
static uint16_t float_to_half_rn(float f) {
    uint32_t x; std::memcpy(&x, &f, 4);
    const uint32_t sign = (x >> 16) & 0x8000u;
    const uint32_t absx = x & 0x7FFFFFFFu;
    if (absx >= 0x7F800000u) return (uint16_t)(sign | 0x7C00u | ((absx > 0x7F800000u) ? 0x0200u : 0u));   // inf / nan
    if (absx >= 0x477FF000u) return (uint16_t)(sign | 0x7C00u);                                             // overflow -> inf
    if (absx < 0x33000001u) return (uint16_t)sign;                                                          // underflow -> 0
    int exp = (int)(absx >> 23) - 127 + 15;
    uint32_t mant = absx & 0x007FFFFFu;
    if (exp <= 0) {                                   // subnormal half
        mant |= 0x00800000u;
        const int shift = 14 - exp;                   // 14..24
        uint32_t half_mant = mant >> shift;
        const uint32_t rem = mant & ((1u << shift) - 1u), halfway = 1u << (shift - 1);
        if (rem > halfway || (rem == halfway && (half_mant & 1u))) half_mant++;
        return (uint16_t)(sign | half_mant);
    }
    uint32_t h = ((uint32_t)exp << 10) | (mant >> 13);
    const uint32_t rem = mant & 0x1FFFu;
    if (rem > 0x1000u || (rem == 0x1000u && (h & 1u))) h++;   // may carry into the exponent: still correct
    return (uint16_t)(sign | h);
}

// Texture2D.SampleLevel(linearClampSampler, uv, 0) DEFINED as: texel centres at (i + 0.5) / size, clamp
// addressing, fp32 weights, lerp(a, b, t) = a + (b - a) * t, x first then y. Texels are float4, row 0 at v = 0.
static void sample_bilinear_clamp(const float* tex, int tw, int th, float u, float v, float out[4]) {
    const float x = u * (float)tw - 0.5f, y = v * (float)th - 0.5f;
    const float x0f = floorf(x), y0f = floorf(y);
    const float fx = x - x0f, fy = y - y0f;
    auto clampi = [](float f, int hi) { if (!(f >= 0.0f)) return 0; if (f > (float)hi) return hi; return (int)f; };
    const int x0 = clampi(x0f, tw - 1), x1 = clampi(x0f + 1.0f, tw - 1);
    const int y0 = clampi(y0f, th - 1), y1 = clampi(y0f + 1.0f, th - 1);
    const float* c00 = tex + ((size_t)y0 * tw + x0) * 4; const float* c10 = tex + ((size_t)y0 * tw + x1) * 4;
    const float* c01 = tex + ((size_t)y1 * tw + x0) * 4; const float* c11 = tex + ((size_t)y1 * tw + x1) * 4;
    for (int k = 0; k < 4; ++k) {
        const float top = c00[k] + (c10[k] - c00[k]) * fx;
        const float bot = c01[k] + (c11[k] - c01[k]) * fx;
        out[k] = top + (bot - top) * fy;
    }
}

extern "C" void usrt_oracle_shade(const RaycastResult* hits, uint64_t count, const Triangle* triangleData,
                                  const float* texture_rgba, int tex_w, int tex_h, uint16_t* out_rgba16f) {
    // :181 `const float lightDir = normalize(float3(1,1,1))` is declared SCALAR: it keeps .x only
    const float len = sqrtf(1.0f * 1.0f + 1.0f * 1.0f + 1.0f * 1.0f);
    const float lightDir = 1.0f / len;
    for (uint64_t i = 0; i < count; ++i) {
        const RaycastResult& r = hits[i];
        const Triangle& t = triangleData[r.triangleIndex];                       // :178 (triangle 0 for a miss)
        const float bu = r.uv[0], bv = r.uv[1];
        const float bw = 1 - bu - bv;
        float uv[2], normal[3];
        for (int k = 0; k < 2; ++k) uv[k] = bw * t.a_uv[k] + bu * t.b_uv[k] + bv * t.c_uv[k];           // :179
        for (int k = 0; k < 3; ++k) normal[k] = bw * t.a_n[k] + bu * t.b_n[k] + bv * t.c_n[k];         // :180
        // dot(scalar, float3): the scalar is splatted
        const float ndotl = lightDir * normal[0] + lightDir * normal[1] + lightDir * normal[2];
        const float shade = fmaxf(0.4f, ndotl);                                  // :183
        float texel[4];
        sample_bilinear_clamp(texture_rgba, tex_w, tex_h, uv[0], uv[1], texel);
        const float a = (r.distance != MAX_FLOAT) ? 1.0f : 0.0f;                 // :184
        out_rgba16f[i * 4 + 0] = float_to_half_rn(texel[0] * shade);
        out_rgba16f[i * 4 + 1] = float_to_half_rn(texel[1] * shade);
        out_rgba16f[i * 4 + 2] = float_to_half_rn(texel[2] * shade);
        out_rgba16f[i * 4 + 3] = float_to_half_rn(a);
    }
}

extern "C" uint16_t usrt_oracle_float_to_half(float f) { return float_to_half_rn(f); }

// ---- diffuse bounce rays: BASELINE.json configs[4] ("3840x2160 x 64 spp random diffuse rays") ------------------
// The reference casts primary rays only (Raytracing.compute:105-176); nothing in it generates secondary rays, so
// this function DEFINES them for the config (parity unpinned, like the rest of this file): hit point on the primary
// ray (PrimaryRay above, Raytracing.compute:108-126), geometric normal of the hit triangle (cross/normalize in the
// canonical expansions of SURVEY 8a) turned against the ray, plus a random unit vector from a counter-based hash
// (splitmix64 finaliser, rejection sampling in the cube, trig-free) -- a cosine-weighted hemisphere direction.
static inline uint64_t hash_u64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

static inline void DiffuseRay(const RaycastResult& hit, const Ray& primary, const Triangle* triangleData, uint64_t seed,
                              uint32_t pixel, uint32_t sample, float* out8) {
    for (int k = 0; k < 8; ++k) out8[k] = 0.0f;                      // the null ray: hits nothing (det = 0)
    if (hit.distance == MAX_FLOAT) return;
    const float t = hit.distance;
    float P[3];
    for (int k = 0; k < 3; ++k) P[k] = primary.origin[k] + primary.dir[k] * t;
    const Triangle& tri = triangleData[hit.triangleIndex];
    float e1[3], e2[3], n[3];
    for (int k = 0; k < 3; ++k) { e1[k] = tri.b[k] - tri.a[k]; e2[k] = tri.c[k] - tri.a[k]; }
    cross3(e1, e2, n);
    const float n2 = dot3(n, n);
    if (!(n2 > 0.0f)) return;
    const float nl = sqrtf(n2);
    for (int k = 0; k < 3; ++k) n[k] = n[k] / nl;
    if (dot3(n, primary.dir) > 0.0f) for (int k = 0; k < 3; ++k) n[k] = -n[k];
    float u[3] = {n[0], n[1], n[2]};                                  // fallback after 16 rejected draws
    const uint64_t stream = hash_u64(seed ^ hash_u64(((uint64_t)pixel << 16) | (uint64_t)(sample & 0xFFFFu)));
    for (uint32_t k = 0; k < 16u; ++k) {
        const uint64_t bits = hash_u64(stream + k);
        float v[3];                                                   // three 21-bit fields -> [-1, 1), exact in fp32
        v[0] = (float)(uint32_t)(bits & 0x1FFFFFu) * 9.5367431640625e-07f - 1.0f;
        v[1] = (float)(uint32_t)((bits >> 21) & 0x1FFFFFu) * 9.5367431640625e-07f - 1.0f;
        v[2] = (float)(uint32_t)((bits >> 42) & 0x1FFFFFu) * 9.5367431640625e-07f - 1.0f;
        const float l2 = dot3(v, v);
        if (l2 <= 1.0f && l2 > 1e-4f) {
            const float l = sqrtf(l2);
            for (int c = 0; c < 3; ++c) u[c] = v[c] / l;
            break;
        }
    }
    float d[3];
    for (int k = 0; k < 3; ++k) d[k] = n[k] + u[k];
    const float d2 = dot3(d, d);
    if (d2 < 1e-8f) { for (int k = 0; k < 3; ++k) d[k] = n[k]; }
    else { const float dl = sqrtf(d2); for (int k = 0; k < 3; ++k) d[k] = d[k] / dl; }
    for (int k = 0; k < 3; ++k) { out8[k] = P[k] + n[k] * 0.001f; out8[4 + k] = d[k]; }
}

// rays_out: 8 floats per ray at index (sample - first_sample) * W * H + pixel
extern "C" void usrt_oracle_diffuse_rays(const RaycastResult* primary_hits, const Triangle* triangleData, int screenWidth,
                                         int screenHeight, float near, float cameraFov, const float* m, uint64_t seed,
                                         uint32_t first_sample, uint32_t num_samples, float* rays_out) {
    const uint64_t frame = (uint64_t)screenWidth * (uint64_t)screenHeight;
    for (uint32_t s = 0; s < num_samples; ++s)
        for (uint32_t y = 0; y < (uint32_t)screenHeight; ++y)
            for (uint32_t x = 0; x < (uint32_t)screenWidth; ++x) {
                const uint32_t pixel = y * (uint32_t)screenWidth + x;
                const Ray primary = PrimaryRay(x, y, screenWidth, screenHeight, near, cameraFov, m);
                DiffuseRay(primary_hits[pixel], primary, triangleData, seed, pixel, first_sample + s,
                           rays_out + ((uint64_t)s * frame + pixel) * 8);
            }
}

// Leaf visiting order of the reference DFS when every box test passes (left leaf, right leaf, then
// the RIGHT internal subtree before the LEFT one -- Raytracing.compute:148-175 push order).
void usrt_oracle_visit_order(const uint32_t* sortedTriangleIndices, const InternalNode* internalNodes,
                             const LeafNode* leafNodes, uint32_t trianglesCount, uint32_t* order) {
    std::vector<uint32_t> stack;
    stack.push_back(0);
    uint32_t k = 0;
    while (!stack.empty()) {
        const uint32_t index = stack.back(); stack.pop_back();
        const InternalNode& nd = internalNodes[index];
        if (nd.leftNodeType == INTERNAL_NODE) stack.push_back(nd.leftNode);
        else order[k++] = sortedTriangleIndices[leafNodes[nd.leftNode].index];
        if (nd.rightNodeType == INTERNAL_NODE) stack.push_back(nd.rightNode);
        else order[k++] = sortedTriangleIndices[leafNodes[nd.rightNode].index];
    }
    (void)trianglesCount;
}

}  // extern "C"
