// usrt_internal.cuh -- shared declarations of libusrt_b200.so (not part of the public ABI).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/usrt.h"

static_assert(sizeof(usrt_aabb) == 32, "AABB layout (Constants.cginc:9-15)");
static_assert(sizeof(usrt_internal_node) == 24, "InternalNode layout (Constants.cginc:20-28)");
static_assert(sizeof(usrt_leaf_node) == 8, "LeafNode layout (Constants.cginc:30-34)");
static_assert(sizeof(usrt_triangle) == 128, "Triangle layout (Constants.cginc:36-54)");
static_assert(sizeof(usrt_raycast_result) == 16, "RaycastResult layout (Raytracing.compute:30-35)");

namespace usrt {

constexpr int kNumSMs = 148;          // B200: 2 dies x 74 SMs
constexpr int kRadixBits = 8;         // Constants.cginc:1  RADIX
constexpr int kRadix = 256;           // Constants.cginc:2  BUCKET_SIZE
constexpr int kSortPasses = 4;        // ComputeBufferSorter.cs:102  bitOffset = 0,8,16,24

// ---- radix sort scratch (K2) ----------------------------------------------------------------
struct SortScratch {
    uint32_t* keys_alt = nullptr;      // ping-pong partners for the standalone sorter
    uint32_t* vals_alt = nullptr;
    uint64_t alt_capacity = 0;         // elements
    uint2* pairs_x = nullptr;          // interleaved {key, value} records of the intermediate passes of large sorts
    uint2* pairs_y = nullptr;
    uint64_t pairs_capacity = 0;
    uint64_t* keys64_alt = nullptr;    // the same for the 64-bit-key sorter
    uint32_t* vals64_alt = nullptr;
    uint64_t alt64_capacity = 0;
    uint32_t* hist = nullptr;          // [8][256] digit histograms (4 used by 32-bit keys), then exclusive digit bases
    void* status = nullptr;            // [4 tile counters (as 64 x u32 header)] + [passes][tiles][256] look-back words
    uint64_t status_bytes = 0;
    uint64_t generation = 0;           // bumped whenever hist / status are (re)allocated: captured graphs hold these pointers
};

// Launches enqueue on `stream`; every function returns the first CUDA error it saw.
// K2: stable LSD sort of count pairs, result back in (keys, vals). alt buffers must hold count elements.
// events (optional, 6): recorded before the histogram, after it, and after each of the 4 passes.
cudaError_t sort_pairs(uint32_t* keys, uint32_t* vals, uint32_t* keys_alt, uint32_t* vals_alt, uint64_t count,
                       SortScratch& scratch, cudaStream_t stream, uint64_t* launches, cudaEvent_t* events = nullptr);
// 64-bit keys, 8 passes; same contract (ComputeBufferSorter<ulong, uint>)
cudaError_t sort_pairs64(uint64_t* keys, uint32_t* vals, uint64_t* keys_alt, uint32_t* vals_alt, uint64_t count, SortScratch& scratch,
                         cudaStream_t stream, uint64_t* launches);
cudaError_t sort_scratch_reserve64(SortScratch& scratch, uint64_t count, bool need_alt);
// One stable partition pass src -> dst by digit (key >> bit_offset) & 255.
cudaError_t partition_pass(const uint32_t* src_keys, const uint32_t* src_vals, uint32_t* dst_keys, uint32_t* dst_vals,
                           uint64_t count, int bit_offset, uint32_t* histogram_out, SortScratch& scratch,
                           cudaStream_t stream, uint64_t* launches);
// multi-GPU bucket exchange: raw digit counts, then the partition pass scattering straight into per-digit
// destination addresses (peer memory over NVLink)
cudaError_t digit_histogram(const uint32_t* keys, uint64_t count, int bit_offset, uint32_t* hist_out, SortScratch& scratch,
                            cudaStream_t stream, uint64_t* launches);
cudaError_t partition_scatter(const uint32_t* src_keys, const uint32_t* src_vals, uint64_t count, int bit_offset,
                              const unsigned long long* key_ptrs, const unsigned long long* val_ptrs, SortScratch& scratch,
                              cudaStream_t stream, uint64_t* launches);
// the landing plan of the bucket exchange from the all-gathered histograms (device in, device out)
cudaError_t peer_scatter_plan(const uint32_t* all_hist, int world, int rank, const unsigned long long* peer_base,
                              unsigned long long capacity, unsigned long long* key_ptrs, unsigned long long* val_ptrs,
                              unsigned long long* recv_total, uint32_t* bounds_out, cudaStream_t stream, uint64_t* launches);
cudaError_t sort_scratch_reserve(SortScratch& scratch, uint64_t count, bool need_alt);
void sort_scratch_free(SortScratch& scratch);

// K1
struct WorldBox { float min[3], max[3]; };            // NormalizeCentroid's box, per axis
// Where K1 reads the three vertices of triangle i: base + i * stride float4s, a / b / c in slots 0..2. stride 8 = the
// reference's 128-byte Triangle array; stride 3 = the compact 48-byte position records (usrt_upload_positions).
struct VertexSource { const float4* base; uint32_t stride; };
// One 16-byte vertex slot of a Triangle record. The build reads the first 48 of a record's 128 bytes: asking L2 to fetch
// 64 bytes around the access instead of the default 128-byte line halves the DRAM reads of K1 / K5 on the Triangle array.
#ifdef __CUDACC__
__device__ __forceinline__ float4 ldg_vertex(const float4* p) {
#ifdef USRT_NO_L2_64B
    return __ldg(p);
#else
    float4 v;
    asm volatile("ld.global.nc.L2::64B.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
#endif
}
#endif
// (Measured: letting K1 also EMIT compact records for K5 costs more than it saves -- K1 34.6 -> 43.0 us for 48 MB of extra
// writes against K5 102.4 -> 96.3 us at 1M triangles -- so K1, K5 and the bounce-ray generator all read whichever source
// the caller uploaded.)
cudaError_t launch_morton(VertexSource src, uint32_t n, const WorldBox& whole, uint32_t* keys, uint32_t* values, usrt_aabb* aabbs,
                          cudaStream_t stream);
// per-axis min / max over all vertices -> out6 (device, min xyz then max xyz)
cudaError_t launch_scene_box(VertexSource src, uint32_t n, float* out6, cudaStream_t stream);

// K3 (src != dst; dst receives the distributed keys). scan_status: >= (tiles+1) x 8 bytes, zeroed by the call.
cudaError_t launch_distribute_keys(const uint32_t* src, uint32_t* dst, uint32_t n, void* scan_status,
                                   cudaStream_t stream, int* launches);
uint64_t distribute_status_bytes(uint32_t n);
// K4
// key_mode 0: 32-bit keys made unique by DistributeKeys (the reference); 1: 32-bit Morton codes with ties broken by
// sorted position (no DistributeKeys); 2: 64-bit keys (keys = const uint64_t*) made unique by the 64-bit DistributeKeys
cudaError_t launch_construct_tree(const void* keys, int key_mode, uint32_t n, usrt_internal_node* internal,
                                  usrt_leaf_node* leaf, uint32_t* up_internal, uint32_t* up_leaf, cudaStream_t stream);
cudaError_t launch_distribute_keys64(const uint64_t* src, uint64_t* dst, uint32_t n, void* scan_status, cudaStream_t stream,
                                     int* launches);
uint64_t distribute_status_bytes64(uint32_t n);
// K1 with 21 bits per axis: 63-bit Morton keys
cudaError_t launch_morton64(VertexSource src, uint32_t n, const WorldBox& whole, uint64_t* keys, uint32_t* values,
                            usrt_aabb* aabbs, cudaStream_t stream);
// K5 (+ packed traversal arrays)
cudaError_t launch_construct_bvh(uint32_t n, const uint32_t* sorted_indices, const usrt_aabb* tri_aabb,
                                 VertexSource vertices, const usrt_internal_node* internal,
                                 const uint32_t* up_internal, const uint32_t* up_leaf, usrt_aabb* bvh, float4* slots,
                                 float4* packed_nodes, float4* packed_tris, cudaStream_t stream);
// packed traversal arrays from the reference-layout buffers (imported BVH)
cudaError_t launch_pack_traversal(uint32_t n, const uint32_t* sorted_indices, const usrt_aabb* tri_aabb,
                                  const usrt_triangle* tris, const usrt_internal_node* internal, const usrt_leaf_node* leaf,
                                  const usrt_aabb* bvh, float4* packed_nodes, float4* packed_tris, cudaStream_t stream);
// imported trees: (1) index ranges + one parent per node, (2) the K4 -> K5 up links + "every leaf reaches node 0 within 64
// steps". err2 (device, 2 words): [0] violations, [1] leaves whose index is not their slot. Step 2 only if err2[0] == 0.
cudaError_t launch_import_validate(uint32_t n, const uint32_t* sorted_indices, const usrt_internal_node* internal,
                                   const usrt_leaf_node* leaf, uint32_t* up_internal, uint32_t* up_leaf, uint32_t* err2,
                                   cudaStream_t stream);
cudaError_t launch_import_links(uint32_t n, const usrt_internal_node* internal, uint32_t* up_internal, uint32_t* up_leaf,
                                uint32_t* err2, cudaStream_t stream);
// validator (MeshBufferContainer.cs:181-195)
cudaError_t launch_count_corrupted(const usrt_leaf_node* leaf, const usrt_internal_node* internal, uint32_t n,
                                   uint32_t* out2, cudaStream_t stream);

// K6
struct TraceScene {
    const float4* packed_nodes;   // 4 x float4 per internal node (both child boxes + child refs)
    const float4* packed_tris;    // 3 x float4 per leaf, in sorted (leaf) order; .w of the first = triangle id
    const usrt_aabb* bvh;         // root box = bvh[0]
};
struct PrimaryParams {
    int width, height;
    float near_plane, tan_half_fov;
    float m[16];                  // row-major cameraToWorld
    int y0, y1;                   // frame mode: rows [y0, y1), record index y*width + x
    // sharded mode (num_shards > 0): this shard owns the row blocks b with b % num_shards == shard,
    // block = block_rows consecutive rows; local row lr <-> frame row
    // ((lr / block_rows) * num_shards + shard) * block_rows + lr % block_rows; record index lr*width + x
    int block_rows, shard, num_shards, local_rows;
};
// Extra destinations of every hit record, written by the trace kernel itself at the same record index as `out`:
// the device-visible alias of a page-locked HOST frame (zero-copy readback) and/or peer GPUs' frame slots mapped
// through CUDA IPC (the ray-sharded all-gather done by the kernel's own stores over NVLink).
constexpr int kMaxHitMirrors = 9;    // 8 frame slots (own + 7 peers) + the pinned host alias
struct HitMirrors {
    usrt_raycast_result* ptr[kMaxHitMirrors];
    int count;
};
cudaError_t launch_trace_primary(const TraceScene& scene, const PrimaryParams& p, usrt_raycast_result* out, int mode,
                                 cudaStream_t stream, const HitMirrors& mirrors);
cudaError_t launch_trace_rays(const TraceScene& scene, const float4* rays, uint64_t num_rays, usrt_raycast_result* out,
                              int mode, cudaStream_t stream);

// count miss records {MAX_FLOAT, 0, (0,0)} (Raytracing.compute:129-131)
cudaError_t launch_fill_miss(usrt_raycast_result* out, uint64_t count, cudaStream_t stream);

// diffuse bounce rays from the primary hit records (BASELINE config 5), s_count samples per pixel starting at s0;
// ray index = (sample - s0) * W * H + pixel
cudaError_t launch_diffuse_rays(const PrimaryParams& p, const usrt_raycast_result* hits, VertexSource vertices,
                                uint64_t seed, uint32_t s0, uint32_t s_count, float4* rays_out, cudaStream_t stream);

// shading epilogue
cudaError_t launch_shade(const usrt_raycast_result* hits, uint64_t count, const usrt_triangle* tris, const float4* tex,
                         int tw, int th, void* out_rgba16f, cudaStream_t stream);

// ---- small device helpers -------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }
__device__ __forceinline__ uint32_t lanemask_lt() {
    uint32_t m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}
__device__ __forceinline__ uint32_t ld_relaxed_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u32(uint32_t* p, uint32_t v) {
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint64_t ld_relaxed_u64(const uint64_t* p) {
    uint64_t v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u64(uint64_t* p, uint64_t v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
#endif

}  // namespace usrt
