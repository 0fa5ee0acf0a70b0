/* usrt.h -- C ABI of libusrt_b200.so: the B200-native (sm_100a) LBVH build + ray-cast path that
 * replaces UnitySimpleRaytracing's HLSL compute kernels and their C# dispatch layer.
 *
 * Boundary. Every entry point below stands in for one reference entry point; the
 * citation is the reference file:line it replaces (paths relative to the reference repo root).
 * Buffer layouts are the reference's, byte for byte (Assets/_Shaders/Constants.cginc:9-54,
 * Assets/_Scripts/SceneDataTypes.cs:4-90).
 *
 * Conventions
 *   - plain C, pointers and sizes only; every function returns 0 on success or a negative
 *     usrt_status; usrt_last_error(ctx) gives the message. Nothing aborts, nothing falls back to a
 *     CPU path: without a CUDA device usrt_create fails with USRT_ERR_CUDA.
 *   - a context owns all device memory and one CUDA stream. Calls enqueue on that stream and are
 *     asynchronous unless they take or return HOST data (upload/download/..._host), which
 *     synchronise. A context is not thread-safe; distinct contexts (one per GPU) are independent.
 *   - host pointers are caller-owned and only touched during the call.
 */
#ifndef USRT_H
#define USRT_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- struct layouts: Constants.cginc:9-54 / SceneDataTypes.cs:4-90 --------------------------- */
typedef struct usrt_aabb { float min[3]; float _dummy0; float max[3]; float _dummy1; } usrt_aabb;            /* 32 B */
typedef struct usrt_internal_node {                                                                           /* 24 B */
    uint32_t leftNode, leftNodeType, rightNode, rightNodeType, parent, index;
} usrt_internal_node;
typedef struct usrt_leaf_node { uint32_t parent, index; } usrt_leaf_node;                                     /*  8 B */
typedef struct usrt_triangle {                                                                                /* 128 B */
    float a[3], _dummy0, b[3], _dummy1, c[3], _dummy2;
    float a_uv[2], b_uv[2], c_uv[2], _dummy3[2];
    float a_normal[3], _dummy4, b_normal[3], _dummy5, c_normal[3], _dummy6;
} usrt_triangle;
/* Raytracing.compute:30-35 -- the hit record. Miss = { (float)0x7F7FFFFF, 0, (0,0) } (Constants.cginc:7). */
typedef struct usrt_raycast_result { float distance; uint32_t triangleIndex; float uv[2]; } usrt_raycast_result; /* 16 B */

#define USRT_INTERNAL_NODE 0u          /* Constants.cginc:17 */
#define USRT_LEAF_NODE 1u              /* Constants.cginc:18 */
#define USRT_NULL 0xFFFFFFFFu          /* SceneDataTypes.cs:63-89 NullLeaf; MeshBufferContainer.cs:108-109 padding */

typedef enum usrt_status {
    USRT_OK = 0,
    USRT_ERR_ARG = -1,        /* bad argument (null pointer, n > capacity, n < 2 for the tree, ...) */
    USRT_ERR_CUDA = -2,       /* a CUDA runtime call or kernel failed; see usrt_last_error */
    USRT_ERR_STATE = -3,      /* stage called before the stage it depends on */
    USRT_ERR_NOMEM = -4
} usrt_status;

/* The seven scene buffers MeshBufferContainer owns (MeshBufferContainer.cs:87-94), by name. */
typedef enum usrt_buffer {
    USRT_BUF_KEYS = 0,            /* uint32[capacity]  _keysBuffer (Morton, later sorted, later distributed) */
    USRT_BUF_TRIANGLE_INDEX = 1,  /* uint32[capacity]  _triangleIndexBuffer (sorted triangle ids) */
    USRT_BUF_TRIANGLE_DATA = 2,   /* usrt_triangle[capacity] */
    USRT_BUF_TRIANGLE_AABB = 3,   /* usrt_aabb[capacity] */
    USRT_BUF_BVH_DATA = 4,        /* usrt_aabb[capacity]  internal-node boxes */
    USRT_BUF_LEAF_NODES = 5,      /* usrt_leaf_node[capacity] */
    USRT_BUF_INTERNAL_NODES = 6,  /* usrt_internal_node[capacity] */
    USRT_BUF_KEYS64 = 7,          /* uint64[capacity]  the keys in key mode USRT_KEYS_MORTON64 (then USRT_BUF_KEYS is void) */
    USRT_BUF_COUNT = 8
} usrt_buffer;

/* Key variants, chosen per context with usrt_set_key_mode; every stage entry point keeps its meaning.
 * Results of modes 1 and 2 differ from the reference's by construction (different keys => different tree); each has an
 * oracle twin and is bit-exact against it. */
typedef enum usrt_key_mode {
    USRT_KEYS_REFERENCE = 0,       /* 30-bit Morton in uint32, made unique by DistributeKeys: the reference, the default */
    USRT_KEYS_INDEX_TIEBREAK = 1,  /* same codes, NO DistributeKeys: delta() breaks ties between equal codes with the sorted
                                      position, i.e. works on (code << 32 | index) -- Karras 2012's own rule, instead of the
                                      reference's "we guarantee that x_code != y_code" (BVH.compute:29). usrt_distribute_keys
                                      returns USRT_ERR_STATE in this mode. */
    USRT_KEYS_MORTON64 = 2         /* 63-bit Morton (21 bits per axis) in uint64, 8-pass sort (the reference's sorter is generic
                                      over uint / ulong keys, ComputeBufferSorter.cs:179-191), 64-bit DistributeKeys, clz64 delta */
} usrt_key_mode;

typedef struct usrt_context usrt_context;

/* ---- lifetime ------------------------------------------------------------------------------ */
/* Replaces the fixed 2^19-slot allocation of MeshBufferContainer.cs:108-115 / Constants.cs:3-6 with
 * a runtime capacity. Keys/indices are filled with 0xFFFFFFFF, node buffers with NullLeaf. */
int usrt_create(int device, uint32_t capacity, usrt_context** out);
int usrt_destroy(usrt_context* ctx);                    /* MeshBufferContainer.Dispose :207-216 */
const char* usrt_last_error(const usrt_context* ctx);   /* stands in for Debug.LogError strings */
const char* usrt_version(void);
int usrt_sync(usrt_context* ctx);
/* Enqueue on an existing CUDA stream (e.g. torch's current stream) instead of the context's own. */
int usrt_set_stream(usrt_context* ctx, void* cuda_stream);
/* Select a key variant (usrt_key_mode). Keys, tree and boxes built in another mode become void (triangles stay). */
int usrt_set_key_mode(usrt_context* ctx, int mode /* usrt_key_mode */);
/* World box of NormalizeCentroid, default -125/+125 (MeshBufferContainer.cs:9-15). */
int usrt_set_world_bounds(usrt_context* ctx, float whole_min, float whole_max);
/* Per-axis box instead of the cube, and the reference's own TODO ("reduce scene data for finding AABB scene in
 * runtime", MeshBufferContainer.cs:7) as an OPT-IN: usrt_fit_world_box reduces the uploaded vertices on the device
 * to their per-axis min / max (an axis on which the mesh is flat gets max = min + 1), makes that the box of
 * NormalizeCentroid and returns it (out pointers may be NULL). Synchronises. Morton keys then differ from the
 * reference's fixed +-125 box by construction; the oracle takes the same box. */
int usrt_set_world_box(usrt_context* ctx, const float box_min[3], const float box_max[3]);
int usrt_fit_world_box(usrt_context* ctx, float out_min[3], float out_max[3]);
uint32_t usrt_capacity(const usrt_context* ctx);
uint32_t usrt_triangles_length(const usrt_context* ctx);   /* MeshBufferContainer.TrianglesLength :30 */

/* ---- MeshBufferContainer(Mesh) : MeshBufferContainer.cs:96-152 --------------------------------- */
/* Copy n packed triangles to the device (synchronous) and re-initialise keys/indices/nodes as the
 * constructor does (:108-115). Does not compute Morton codes; see usrt_morton. */
int usrt_upload_triangles(usrt_context* ctx, const usrt_triangle* host_triangles, uint32_t n);
/* Same, asynchronous: host_triangles must be page-locked and stay untouched until usrt_sync (or a later
 * synchronising call) returns. With two contexts on one GPU, frame i+1's upload overlaps frame i's kernels. */
int usrt_upload_triangles_async(usrt_context* ctx, const usrt_triangle* pinned_host_triangles, uint32_t n);
/* Page-locked host memory for the ..._async entry points and for zero-copy frames (usrt_trace_primary with a pinned
 * host_out): cudaHostAlloc / cudaFreeHost behind the ABI, so that a host without a CUDA binding (the C# P/Invoke host)
 * can get it. (A bare 128 MiB copy from cudaHostAlloc memory measured 55.5 GB/s on B200 against 51.7 GB/s from a malloc
 * block pinned after the fact; the pipelined frame time of bench.py is the same for both.) The buffer belongs to the
 * process, not to the context;
 * free it with usrt_host_free after the last call that uses it has been synchronised. */
int usrt_host_alloc(usrt_context* ctx, uint64_t bytes, void** host_ptr);
int usrt_host_free(usrt_context* ctx, void* host_ptr);
/* ADDITIVE to the reference's SetData of whole Triangle structs (MeshBufferContainer.cs:150): positions only -- 12 floats
 * per triangle, exactly the first 48 bytes of usrt_triangle (a.xyz, pad, b.xyz, pad, c.xyz, pad). Those are the only bytes the
 * build and the traversal ever read, so an animated mesh can re-send 48 instead of 128 bytes per triangle per frame. The
 * Triangle array keeps whatever was uploaded last (uv / normals for usrt_shade); hit records, nodes and boxes are
 * bit-identical to a full upload of the same positions. _async: page-locked memory, untouched until usrt_sync. */
int usrt_upload_positions(usrt_context* ctx, const float* host_positions, uint32_t n);
int usrt_upload_positions_async(usrt_context* ctx, const float* pinned_host_positions, uint32_t n);
/* Same, from a DEVICE pointer (async, device-to-device). */
int usrt_set_triangles_device(usrt_context* ctx, const void* dev_triangles, uint32_t n);
/* K1 -- the CPU loop of MeshBufferContainer.cs:123-146 as a kernel: padded AABB, centroid of the padded
 * box, NormalizeCentroid, Morton3D; keys[i], triangleIndex[i] = i, triangleAABB[i]. */
int usrt_morton(usrt_context* ctx);

/* ---- ComputeBufferSorter<uint,uint>.Sort() : ComputeBufferSorter.cs:100-126 -------------------- */
/* K2 -- stable ascending LSD radix sort, 4 passes x 8 bits over all 32 key bits, of the context's
 * (keys, triangleIndex) pairs [0, trianglesLength). Padding slots keep 0xFFFFFFFF. */
int usrt_sort(usrt_context* ctx);
/* Standalone sorter on caller data (ComputeBufferSorter ctor :44 takes arbitrary key/value buffers).
 * In place. _device: pointers are device memory of this context's GPU, async. _host: synchronous. */
int usrt_sort_pairs_device(usrt_context* ctx, uint32_t* dev_keys, uint32_t* dev_values, uint64_t count);
int usrt_sort_pairs_host(usrt_context* ctx, uint32_t* host_keys, uint32_t* host_values, uint64_t count);
/* ComputeBufferSorter<ulong, uint> (GetRadix is generic over uint / ulong keys, ComputeBufferSorter.cs:179-191): the same
 * stable ascending LSD sort over 64-bit keys, 8 passes x 8 bits, 32-bit values (may be NULL: keys only). In place. */
int usrt_sort_pairs64_device(usrt_context* ctx, uint64_t* dev_keys, uint32_t* dev_values, uint64_t count);
int usrt_sort_pairs64_host(usrt_context* ctx, uint64_t* host_keys, uint32_t* host_values, uint64_t count);
/* One stable partition pass by an arbitrary 8-bit digit (bit_offset in 0..24) from src to dst; used
 * as the bucket-split step of the multi-GPU sort. histogram_out (device, 256 x uint32) may be NULL. */
int usrt_partition_pass_device(usrt_context* ctx, const uint32_t* src_keys, const uint32_t* src_values,
                               uint32_t* dst_keys, uint32_t* dst_values, uint64_t count, int bit_offset,
                               uint32_t* histogram_out);

/* Multi-GPU bucket exchange fused into the partition pass (BASELINE config 3 beyond one GPU): first the raw
 * counts of one 8-bit digit (device, 256 x uint32), then -- once every rank knows every rank's counts and so
 * where its pairs go -- the same stable partition pass, but with one destination base ADDRESS per digit value
 * (device arrays of 256 x uint64): this rank's slice of the receive buffer of the GPU that owns that bucket,
 * opened through usrt_peer_buffer_open. The scatter writes go straight over NVLink; no separate all-to-all. */
int usrt_digit_histogram_device(usrt_context* ctx, const uint32_t* dev_keys, uint64_t count, int bit_offset, uint32_t* dev_hist_out);
int usrt_partition_scatter_device(usrt_context* ctx, const uint32_t* src_keys, const uint32_t* src_values, uint64_t count,
                                  int bit_offset, const uint64_t* dev_key_base, const uint64_t* dev_value_base);
/* The landing plan of that exchange, computed ON THE DEVICE from the all-gathered histograms (dev_all_hist: world x 256
 * uint32, rank-major): contiguous bucket ranges of ~equal mass (dev_bounds, world + 1 entries, may be NULL), this rank's
 * 256 destination addresses inside its owners' receive buffers (dev_peer_base[o] = this process's mapping of rank o's
 * buffer, laid out keys[capacity] | values[capacity]; source-rank-major landing order keeps the sort globally stable)
 * and the number of pairs every rank receives (dev_recv_total, world x uint64). Every rank computes the same plan from
 * the same input; only dev_recv_total has to be read by the host (the count of the local sort that follows). */
int usrt_peer_scatter_plan_device(usrt_context* ctx, const uint32_t* dev_all_hist, int world, int rank, const uint64_t* dev_peer_base,
                                  uint64_t capacity, uint64_t* dev_key_base, uint64_t* dev_value_base, uint64_t* dev_recv_total,
                                  uint32_t* dev_bounds);
/* Device buffers that other processes on the node can map (CUDA IPC): create returns the pointer and a 64-byte
 * handle to send to the peers; open maps a peer's buffer into this process; close with opened = 1 unmaps a
 * peer's buffer, opened = 0 frees an own one. */
int usrt_peer_buffer_create(usrt_context* ctx, uint64_t bytes, void** dev_ptr, unsigned char handle_out[64]);
int usrt_peer_buffer_open(usrt_context* ctx, const unsigned char handle[64], void** dev_ptr);
int usrt_peer_buffer_close(usrt_context* ctx, void* dev_ptr, int opened);

/* ---- MeshBufferContainer.DistributeKeys() : MeshBufferContainer.cs:154-169 -------------------- */
/* K3 -- new[0]=0; new[i]=new[i-1]+max(k[i]-k[i-1],1) in wrapping uint32 over [0, trianglesLength). */
int usrt_distribute_keys(usrt_context* ctx);

/* ---- BVHConstructor : BVHConstructor.cs:24-69 ---------------------------------------------- */
/* K4 -- ConstructTree() :61-64 -> kernel TreeConstructor (BVH.compute:94-149). Needs n >= 2. */
int usrt_construct_tree(usrt_context* ctx);
/* K5 -- ConstructBVH() :66-69 -> kernel BVHConstructor (BVH.compute:172-220). Re-runnable: the
 * counter buffer (BVHConstructor.cs:41) is self-resetting. Also emits the traversal-side packed
 * node/triangle arrays the trace kernels read. */
int usrt_construct_bvh(usrt_context* ctx);
/* RaytracingMeshDrawer.Awake() :34-51 as one enqueue: K1 -> K2 -> K3 -> K4 -> K5, no host sync. */
int usrt_rebuild(usrt_context* ctx);
/* Device time of the last usrt_rebuild, per stage, in ms: morton, sort, distribute, tree, bvh, total.
 * Synchronises. Only filled when timing was enabled before the rebuild. */
int usrt_enable_stage_timing(usrt_context* ctx, int enabled);
int usrt_last_rebuild_ms(usrt_context* ctx, float out_ms[6]);
/* Device time of the last sort (usrt_sort, usrt_sort_pairs_device, or the one inside usrt_rebuild), per
 * kernel group, in ms: histogram+scan, pass bitOffset 0, 8, 16, 24, total. Synchronises. */
int usrt_last_sort_ms(usrt_context* ctx, float out_ms[6]);

/* Ray-sharded frames without a separate all-gather: up to 8 extra DEVICE destinations (frame slots of this and
 * of peer GPUs, the latter opened with usrt_peer_buffer_open) that usrt_trace_primary / usrt_trace_primary_sharded write every hit record
 * to as well, at the same record index, by the trace kernel's own stores over NVLink. count = 0 clears. The
 * caller fences across ranks (e.g. a one-element all-reduce) before reading a peer-written slot. */
int usrt_set_hit_mirrors(usrt_context* ctx, int count, void* const* dev_ptrs);

/* ---- Dispatch(Raytracing) : RaytracingMeshDrawer.cs:76-84, Raytracing.compute:105-176 ---------- */
/* K6 -- one hit record per pixel, index y*width + x, row 0 = most negative camera-space y.
 * near = _ProjectionParams.y; tan_half_fov = tan(fovDeg*Deg2Rad/2) (RaytracingMeshDrawer.cs:80);
 * camera_to_world: 16 floats ROW-major (m[r*4+c]); Unity's Matrix4x4 fields m{r}{c} map directly.
 * Rows [y0, y1) only are traced (y0=0,y1=height for a full frame) -- the ray-sharding hook.
 * host_out (may be NULL) receives rows [y0,y1) at their frame positions, i.e. host_out + y0*width;
 * the device copy stays readable through usrt_hits_device. */
int usrt_trace_primary(usrt_context* ctx, int width, int height, float near_plane, float tan_half_fov,
                       const float camera_to_world[16], int y0, int y1, usrt_raycast_result* host_out);
/* Full frame, asynchronous: pinned_host_out must be page-locked; the kernel writes the records straight into it
 * and the call returns once enqueued -- read the frame after usrt_sync. */
int usrt_trace_primary_async(usrt_context* ctx, int width, int height, float near_plane, float tan_half_fov,
                             const float camera_to_world[16], usrt_raycast_result* pinned_host_out);
/* Ray sharding across GPUs against a replicated BVH (north_star (a)): shard s of S traces the row
 * blocks b (block = block_rows consecutive rows) with b % S == s -- interleaved for load balance -- in
 * ONE launch and writes them compactly: local row lr = (b / S) * block_rows + row_in_block, record
 * index lr * width + x. Every shard's buffer has ceil(ceil(H / block_rows) / S) * block_rows rows
 * (equal sizes for an all-gather); rows that fall outside the frame hold MISS records
 * { (float)0x7F7FFFFF, 0, (0,0) }. dev_out may be
 * NULL (the context's hit buffer is used, see usrt_hits_device); host_out may be NULL. */
int usrt_trace_primary_sharded(usrt_context* ctx, int width, int height, float near_plane, float tan_half_fov,
                               const float camera_to_world[16], int block_rows, int shard, int num_shards,
                               void* dev_out, usrt_raycast_result* host_out);
/* Same traversal for caller rays: 8 floats per ray (origin.xyz, pad, dir.xyz, pad); dir is used as
 * given, inv_dir = 1/dir. */
int usrt_trace_rays(usrt_context* ctx, const float* host_rays, uint64_t num_rays, usrt_raycast_result* host_out);
int usrt_trace_rays_device(usrt_context* ctx, const void* dev_rays, uint64_t num_rays, void* dev_out);
/* Diffuse bounce rays (BASELINE config 5, "64 spp random diffuse rays"; the reference itself casts primary rays
 * only, Raytracing.compute:105-176, so the generator is defined by the oracle): for every pixel of the frame and
 * every sample s in [first_sample, first_sample + num_samples), one ray from the pixel's primary hit point along
 * normal + random unit vector (cosine-weighted), seeded by (seed, pixel, s). dev_primary_hits = the W*H records of
 * usrt_trace_primary with the same camera (NULL: the context's own, if that was the last trace). Output: 8 floats
 * per ray in the usrt_trace_rays layout at index (s - first_sample) * W * H + pixel; pixels without a hit give the
 * null ray (all zeros), which hits nothing. The context's triangle buffer must not be re-uploaded in between. */
int usrt_diffuse_rays_device(usrt_context* ctx, int width, int height, float near_plane, float tan_half_fov,
                             const float camera_to_world[16], const void* dev_primary_hits, uint64_t seed,
                             uint32_t first_sample, uint32_t num_samples, void* dev_rays_out);
/* Device pointer of the hit records written by the last trace call (usrt_raycast_result[]). */
int usrt_hits_device(usrt_context* ctx, void** dev_ptr, uint64_t* count);
/* 0 = strict (default): the reference's visiting semantics exactly -- every box the ray line touches
 * is visited, no distance culling. 1 = culled: skips boxes entirely beyond the current closest hit;
 * 2 = culled, and of two internal children that are both hit the nearer one is walked first (SURVEY.md 8f-4
 * "distance-culled + near-first traversal"). Modes 1 and 2 are NOT part of the parity contract (reported
 * separately): equally distant triangles can resolve to a different id than the reference's visiting order gives. */
int usrt_set_trace_mode(usrt_context* ctx, int mode);

/* ---- importing a finished BVH ------------------------------------------------------ */
/* Install the seven scene buffers of MeshBufferContainer.cs:87-94 from HOST arrays in the reference's
 * layouts (n triangles: n keys / indices / triangles / triangle AABBs / leaf nodes, n-1 node AABBs and
 * internal nodes) -- a tree that was built elsewhere: a dump of this library, or the buffers the
 * reference's own kernels produced. The input is treated as UNTRUSTED: child / leaf / triangle indices are range-
 * checked on the device, every node must have exactly one parent, and every leaf must reach node 0 within 64 levels
 * (the depth of the traversal stack, Raytracing.compute:133); otherwise USRT_ERR_ARG and the context holds no scene.
 * The traversal-side arrays are derived on the device; afterwards the context traces exactly as if it had built the
 * tree itself, and usrt_construct_bvh may re-fit it (when every leaf sits in its own slot, as TreeConstructor's do;
 * else USRT_ERR_STATE). Synchronous. */
int usrt_upload_bvh(usrt_context* ctx, uint32_t n, const uint32_t* keys, const uint32_t* triangle_index,
                    const usrt_triangle* triangles, const usrt_aabb* triangle_aabb, const usrt_aabb* bvh_data,
                    const usrt_leaf_node* leaf_nodes, const usrt_internal_node* internal_nodes);

/* ---- shading epilogue : Raytracing.compute:178-184 ------------------------------ */
/* _meshTexture (Raytracing.compute:13) as width x height float4 texels, row 0 at v = 0. Synchronous. */
int usrt_upload_texture(usrt_context* ctx, const float* host_rgba, int width, int height);
/* Shades the hit records of the last trace call (triangle uv/normal interpolation, the scalar
 * lightDir quirk of :181, max(0.4, .), bilinear-clamp texture fetch) into RGBA16F, 4 halfs per record,
 * alpha = hit (the reference's R16G16B16A16_SFloat render target, RaytracingMeshDrawer.cs:56).
 * dev_out may be NULL (an internal buffer is used); host_out may be NULL. */
int usrt_shade(usrt_context* ctx, void* dev_out, uint16_t* host_out);

/* ---- DataBuffer<T>.GetData() : DataBuffer.cs:50-54 ------------------------------------------- */
/* Copy `count` elements of a scene buffer to host memory (synchronous). */
int usrt_download(usrt_context* ctx, int buffer /* usrt_buffer */, void* host_dst, uint64_t count);
/* Raw device pointer of a scene buffer (for zero-copy interop, e.g. torch / NCCL broadcast). The
 * KEYS pointer can change after usrt_sort / usrt_distribute_keys (ping-pong); re-query it. */
int usrt_device_ptr(usrt_context* ctx, int buffer /* usrt_buffer */, void** dev_ptr);
/* Validator of MeshBufferContainer.GetAllGpuData :181-195 on the device: counts leaf [0,n) and
 * internal [0,n-1) entries that still equal NullLeaf. */
int usrt_count_corrupted_nodes(usrt_context* ctx, uint32_t* leaf_corrupted, uint32_t* internal_corrupted);

/* Number of kernels this library has launched on the context since creation (bench bookkeeping). */
uint64_t usrt_kernel_launches(const usrt_context* ctx);

#ifdef __cplusplus
}
#endif
#endif /* USRT_H */
