// api.cu -- the extern "C" boundary of libusrt_b200.so (declared in include/usrt.h): context and
// buffer ownership, stage sequencing, host<->device copies. It replaces the reference's C# dispatch
// layer: MeshBufferContainer.cs (buffers, Morton, DistributeKeys), ComputeBufferSorter.cs (Sort),
// BVHConstructor.cs (ConstructTree / ConstructBVH), DataBuffer.cs (GetData / Sync) and the build +
// trace sequence of RaytracingMeshDrawer.cs:30-54,76-84.

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include "usrt_internal.cuh"

#include <cmath>

using namespace usrt;

enum Stage : uint32_t { ST_TRIS = 1, ST_MORTON = 2, ST_SORTED = 4, ST_DISTRIBUTED = 8, ST_TREE = 16, ST_BVH = 32 };

struct usrt_context {
    int device = 0;
    uint32_t capacity = 0;
    uint32_t n = 0;                       // trianglesLength
    uint32_t dirty_n = 0;                 // slots [0, dirty_n) may differ from their initial fill
    cudaStream_t stream = nullptr;
    cudaStream_t own_stream = nullptr;
    WorldBox whole{{-125.0f, -125.0f, -125.0f}, {125.0f, 125.0f, 125.0f}};   // MeshBufferContainer.cs:9-15
    uint32_t stage = 0;

    // the seven scene buffers of MeshBufferContainer.cs:87-94 (+ ping-pong partners for keys/indices)
    uint32_t *keys = nullptr, *keys_alt = nullptr;
    uint32_t* keys_primary = nullptr;     // the buffer Morton codes are generated into (fixed for the context)
    uint32_t *tri_index = nullptr, *tri_index_alt = nullptr;
    // key variants, usrt_set_key_mode: 64-bit key buffers exist only once mode 2 was selected
    int key_mode = USRT_KEYS_REFERENCE;
    uint64_t *keys64 = nullptr, *keys64_alt = nullptr, *keys64_primary = nullptr;
    void* scan_status64 = nullptr;
    usrt_triangle* triangles = nullptr;
    // compact 48-byte position records {a, b, c} per triangle, filled by usrt_upload_positions only (allocated on first use)
    float4* positions = nullptr;
    bool positions_only = false;          // the last upload was positions only: K1 reads `positions`, `triangles` is stale
    bool graph_positions_only = false;
    usrt_aabb* tri_aabb = nullptr;
    usrt_aabb* bvh = nullptr;
    usrt_leaf_node* leaf = nullptr;
    usrt_internal_node* internal = nullptr;
    // stands in for BVHConstructor.cs:16 _atomics: per-node exchange slots (2 x 16 B), empty = 0xFF..,
    // emptied again by the merging arrival, so ConstructBVH is re-runnable without a memset
    float4* slots = nullptr;
    uint32_t *up_internal = nullptr, *up_leaf = nullptr;   // K4 -> K5 parent links (side + locality bits)
    // traversal-side arrays written by K5
    float4* packed_nodes = nullptr;
    float4* packed_tris = nullptr;
    // scratch
    SortScratch sort;
    void* scan_status = nullptr;
    uint32_t* small = nullptr;            // a few device words (validators)
    float* scene_box = nullptr;           // 6 floats: usrt_fit_world_box
    // trace
    usrt_raycast_result* hits = nullptr;
    uint64_t hits_capacity = 0, hits_count = 0;
    float4* rays = nullptr;
    uint64_t rays_capacity = 0;
    int trace_mode = 0;
    HitMirrors mirrors{};             // usrt_set_hit_mirrors: peer GPUs' frame slots
    // shading epilogue
    float4* texture = nullptr;
    int tex_w = 0, tex_h = 0;
    void* shaded = nullptr;
    uint64_t shaded_capacity = 0;
    // timing
    bool timing = false;
    cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    bool ev_valid = false;
    cudaEvent_t sort_ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    bool sort_ev_valid = false;

    // CUDA graph of the rebuild sequence
    bool use_graph = true;
    cudaGraphExec_t graph_exec = nullptr;
    uint32_t graph_n = 0;
    cudaStream_t graph_stream = nullptr;
    uint64_t graph_sort_generation = 0;   // SortScratch::generation the captured launches point into
    uint64_t graph_launches = 0;

    uint64_t launches = 0;
    char err[512] = {0};
};

namespace {

int fail(usrt_context* ctx, int code, const char* fmt, ...) {
    if (ctx) {
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(ctx->err, sizeof(ctx->err), fmt, ap);
        va_end(ap);
    }
    return code;
}

#define CU(ctx, call)                                                                                        \
    do {                                                                                                     \
        cudaError_t e__ = (call);                                                                            \
        if (e__ != cudaSuccess)                                                                              \
            return fail(ctx, e__ == cudaErrorMemoryAllocation ? USRT_ERR_NOMEM : USRT_ERR_CUDA, "%s: %s", #call, \
                        cudaGetErrorString(e__));                                                            \
    } while (0)

#define NEED_CTX(ctx)                 \
    do {                              \
        if (!(ctx)) return USRT_ERR_ARG; \
    } while (0)

int bind_device(usrt_context* ctx) {
    CU(ctx, cudaSetDevice(ctx->device));
    return USRT_OK;
}

// MeshBufferContainer.cs:108-115: keys/indices = uint.MaxValue, leaf/internal = NullLeaf (all 0xFF).
// Only slots [lo, hi) are touched: a build with n triangles rewrites every slot below n, so after a
// full initialisation only the slots a previous, larger mesh dirtied ([n, previous n)) need restoring.
int reset_scene_buffers(usrt_context* ctx, size_t lo, size_t hi) {
    if (hi <= lo) return USRT_OK;
    const size_t c = hi - lo;
    CU(ctx, cudaMemsetAsync(ctx->keys + lo, 0xFF, c * 4, ctx->stream));
    CU(ctx, cudaMemsetAsync(ctx->keys_alt + lo, 0xFF, c * 4, ctx->stream));
    CU(ctx, cudaMemsetAsync(ctx->tri_index + lo, 0xFF, c * 4, ctx->stream));
    CU(ctx, cudaMemsetAsync(ctx->tri_index_alt + lo, 0xFF, c * 4, ctx->stream));
    if (ctx->keys64) {
        CU(ctx, cudaMemsetAsync(ctx->keys64 + lo, 0xFF, c * 8, ctx->stream));
        CU(ctx, cudaMemsetAsync(ctx->keys64_alt + lo, 0xFF, c * 8, ctx->stream));
    }
    CU(ctx, cudaMemsetAsync(ctx->leaf + lo, 0xFF, c * sizeof(usrt_leaf_node), ctx->stream));
    CU(ctx, cudaMemsetAsync(ctx->tri_aabb + lo, 0, c * sizeof(usrt_aabb), ctx->stream));
    // internal nodes / node boxes are written for [0, n-1): slot n-1 of the new mesh must be restored too
    const size_t ilo = lo > 0 ? lo - 1 : 0;
    CU(ctx, cudaMemsetAsync(ctx->internal + ilo, 0xFF, (hi - ilo) * sizeof(usrt_internal_node), ctx->stream));
    CU(ctx, cudaMemsetAsync(ctx->bvh + ilo, 0, (hi - ilo) * sizeof(usrt_aabb), ctx->stream));
    CU(ctx, cudaMemsetAsync(ctx->slots, 0xFF, (size_t)ctx->capacity * 2 * sizeof(float4), ctx->stream));   // BVHConstructor.cs:41
    return USRT_OK;
}

int ensure_hits(usrt_context* ctx, uint64_t count) {
    if (count > ctx->hits_capacity) {
        if (ctx->hits) CU(ctx, cudaFree(ctx->hits));
        ctx->hits = nullptr; ctx->hits_capacity = 0;
        CU(ctx, cudaMalloc(&ctx->hits, count * sizeof(usrt_raycast_result)));
        ctx->hits_capacity = count;
        // a partial trace (rows [y0,y1)) leaves the other records untouched; usrt_shade / usrt_diffuse_rays walk all of
        // them and index the triangle buffer by triangleIndex, so fresh memory starts as misses (Raytracing.compute:129-131)
        CU(ctx, launch_fill_miss(ctx->hits, count, ctx->stream));
        ctx->launches += 1;
    }
    return USRT_OK;
}

int ensure_rays(usrt_context* ctx, uint64_t count) {
    if (count > ctx->rays_capacity) {
        if (ctx->rays) CU(ctx, cudaFree(ctx->rays));
        ctx->rays = nullptr; ctx->rays_capacity = 0;
        CU(ctx, cudaMalloc(&ctx->rays, count * 2 * sizeof(float4)));
        ctx->rays_capacity = count;
    }
    return USRT_OK;
}

VertexSource vertex_source(const usrt_context* ctx) {
    if (ctx->positions_only) return VertexSource{ctx->positions, 3u};
    return VertexSource{reinterpret_cast<const float4*>(ctx->triangles), 8u};
}

int do_morton(usrt_context* ctx) {
    if (ctx->key_mode == USRT_KEYS_MORTON64)
        CU(ctx, launch_morton64(vertex_source(ctx), ctx->n, ctx->whole, ctx->keys64, ctx->tri_index, ctx->tri_aabb, ctx->stream));
    else
        CU(ctx, launch_morton(vertex_source(ctx), ctx->n, ctx->whole, ctx->keys, ctx->tri_index, ctx->tri_aabb, ctx->stream));
    ctx->launches += 1;
    ctx->stage = ST_TRIS | ST_MORTON;
    return USRT_OK;
}

int do_sort(usrt_context* ctx) {
    ctx->sort_ev_valid = false;
    if (ctx->key_mode == USRT_KEYS_MORTON64) {
        CU(ctx, sort_pairs64(ctx->keys64, ctx->tri_index, ctx->keys64_alt, ctx->tri_index_alt, ctx->n, ctx->sort, ctx->stream,
                             &ctx->launches));
        ctx->stage |= ST_SORTED;
        return USRT_OK;
    }
    CU(ctx, sort_pairs(ctx->keys, ctx->tri_index, ctx->keys_alt, ctx->tri_index_alt, ctx->n, ctx->sort, ctx->stream,
                       &ctx->launches, ctx->timing ? ctx->sort_ev : nullptr));
    ctx->sort_ev_valid = ctx->timing && ctx->n > 0;
    ctx->stage |= ST_SORTED;
    return USRT_OK;
}

int do_distribute(usrt_context* ctx) {
    int l = 0;
    if (ctx->key_mode == USRT_KEYS_MORTON64) {
        CU(ctx, launch_distribute_keys64(ctx->keys64, ctx->keys64_alt, ctx->n, ctx->scan_status64, ctx->stream, &l));
        std::swap(ctx->keys64, ctx->keys64_alt);
    } else {
        CU(ctx, launch_distribute_keys(ctx->keys, ctx->keys_alt, ctx->n, ctx->scan_status, ctx->stream, &l));
        std::swap(ctx->keys, ctx->keys_alt);    // the distributed keys ARE the keys buffer from here on
    }
    ctx->launches += l;
    ctx->stage |= ST_DISTRIBUTED;
    return USRT_OK;
}

int do_tree(usrt_context* ctx) {
    const void* keys = ctx->key_mode == USRT_KEYS_MORTON64 ? static_cast<const void*>(ctx->keys64) : static_cast<const void*>(ctx->keys);
    CU(ctx, launch_construct_tree(keys, ctx->key_mode, ctx->n, ctx->internal, ctx->leaf, ctx->up_internal, ctx->up_leaf, ctx->stream));
    ctx->launches += 1;
    ctx->stage |= ST_TREE;
    return USRT_OK;
}

int do_bvh(usrt_context* ctx) {
    CU(ctx, launch_construct_bvh(ctx->n, ctx->tri_index, ctx->tri_aabb, vertex_source(ctx), ctx->internal, ctx->up_internal,
                                 ctx->up_leaf, ctx->bvh, ctx->slots, ctx->packed_nodes, ctx->packed_tris, ctx->stream));
    ctx->launches += 1;
    ctx->stage |= ST_BVH;
    return USRT_OK;
}

}  // namespace

extern "C" {

const char* usrt_version(void) { return "usrt_b200 0.1 (sm_100a)"; }

int usrt_create(int device, uint32_t capacity, usrt_context** out) {
    if (!out || capacity < 2 || capacity > (1u << 30)) return USRT_ERR_ARG;   // 30-bit node links
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || device < 0 || device >= count) return USRT_ERR_CUDA;   // no CPU fallback, by design
    usrt_context* ctx = new (std::nothrow) usrt_context();
    if (!ctx) return USRT_ERR_NOMEM;
    ctx->device = device;
    ctx->capacity = capacity;
    int rc = USRT_OK;
    auto init = [&]() -> int {
        CU(ctx, cudaSetDevice(device));
        CU(ctx, cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking));
        ctx->stream = ctx->own_stream;
        const size_t c = capacity;
        CU(ctx, cudaMalloc(&ctx->keys, c * 4));
        ctx->keys_primary = ctx->keys;
        ctx->use_graph = getenv("USRT_NO_GRAPH") == nullptr;
        CU(ctx, cudaMalloc(&ctx->keys_alt, c * 4));
        CU(ctx, cudaMalloc(&ctx->tri_index, c * 4));
        CU(ctx, cudaMalloc(&ctx->tri_index_alt, c * 4));
        CU(ctx, cudaMalloc(&ctx->triangles, c * sizeof(usrt_triangle)));
        CU(ctx, cudaMalloc(&ctx->tri_aabb, c * sizeof(usrt_aabb)));
        CU(ctx, cudaMalloc(&ctx->bvh, c * sizeof(usrt_aabb)));
        CU(ctx, cudaMalloc(&ctx->leaf, c * sizeof(usrt_leaf_node)));
        CU(ctx, cudaMalloc(&ctx->internal, c * sizeof(usrt_internal_node)));
        CU(ctx, cudaMalloc(&ctx->slots, c * 2 * sizeof(float4)));
        CU(ctx, cudaMalloc(&ctx->up_internal, c * 4));
        CU(ctx, cudaMalloc(&ctx->up_leaf, c * 4));
        CU(ctx, cudaMalloc(&ctx->packed_nodes, c * 4 * sizeof(float4)));
        CU(ctx, cudaMalloc(&ctx->packed_tris, c * 3 * sizeof(float4)));
        CU(ctx, cudaMalloc(&ctx->scan_status, distribute_status_bytes(capacity)));
        CU(ctx, cudaMalloc(&ctx->small, 64));
        CU(ctx, sort_scratch_reserve(ctx->sort, capacity, false));
        CU(ctx, cudaMemsetAsync(ctx->triangles, 0, c * sizeof(usrt_triangle), ctx->stream));
        for (auto& ev : ctx->ev) CU(ctx, cudaEventCreate(&ev));
        for (auto& ev : ctx->sort_ev) CU(ctx, cudaEventCreate(&ev));
        int r = reset_scene_buffers(ctx, 0, capacity);
        if (r != USRT_OK) return r;
        CU(ctx, cudaStreamSynchronize(ctx->stream));
        return USRT_OK;
    };
    rc = init();
    if (rc != USRT_OK) {
        fprintf(stderr, "usrt_create failed: %s\n", ctx->err);
        usrt_destroy(ctx);
        return rc;
    }
    *out = ctx;
    return USRT_OK;
}

int usrt_destroy(usrt_context* ctx) {
    NEED_CTX(ctx);
    cudaSetDevice(ctx->device);
    if (ctx->own_stream) cudaStreamSynchronize(ctx->own_stream);
    if (ctx->graph_exec) cudaGraphExecDestroy(ctx->graph_exec);
    void* ptrs[] = {ctx->keys, ctx->keys_alt, ctx->tri_index, ctx->tri_index_alt, ctx->triangles, ctx->tri_aabb,
                    ctx->positions, ctx->bvh, ctx->leaf, ctx->internal, ctx->slots, ctx->up_internal, ctx->up_leaf, ctx->packed_nodes, ctx->packed_tris,
                    ctx->scan_status, ctx->scan_status64, ctx->keys64, ctx->keys64_alt, ctx->small, ctx->scene_box, ctx->hits, ctx->rays, ctx->texture, ctx->shaded};
    for (void* p : ptrs)
        if (p) cudaFree(p);
    sort_scratch_free(ctx->sort);
    for (auto& ev : ctx->ev)
        if (ev) cudaEventDestroy(ev);
    for (auto& ev : ctx->sort_ev)
        if (ev) cudaEventDestroy(ev);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
    return USRT_OK;
}

const char* usrt_last_error(const usrt_context* ctx) { return ctx ? ctx->err : "null context"; }

int usrt_sync(usrt_context* ctx) {
    NEED_CTX(ctx);
    if (int r = bind_device(ctx)) return r;
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return USRT_OK;
}

int usrt_set_stream(usrt_context* ctx, void* cuda_stream) {
    NEED_CTX(ctx);
    if (int r = bind_device(ctx)) return r;
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : ctx->own_stream;
    return USRT_OK;
}

int usrt_set_world_bounds(usrt_context* ctx, float whole_min, float whole_max) {
    NEED_CTX(ctx);
    if (!(whole_max > whole_min)) return fail(ctx, USRT_ERR_ARG, "world bounds: max must exceed min");
    for (int k = 0; k < 3; ++k) { ctx->whole.min[k] = whole_min; ctx->whole.max[k] = whole_max; }
    if (ctx->graph_exec) { cudaGraphExecDestroy(ctx->graph_exec); ctx->graph_exec = nullptr; }   // kernel arguments changed
    return USRT_OK;
}

int usrt_set_world_box(usrt_context* ctx, const float box_min[3], const float box_max[3]) {
    NEED_CTX(ctx);
    if (!box_min || !box_max) return fail(ctx, USRT_ERR_ARG, "world box: null pointer");
    for (int k = 0; k < 3; ++k)
        if (!(box_max[k] > box_min[k])) return fail(ctx, USRT_ERR_ARG, "world box: max must exceed min on every axis");
    for (int k = 0; k < 3; ++k) { ctx->whole.min[k] = box_min[k]; ctx->whole.max[k] = box_max[k]; }
    if (ctx->graph_exec) { cudaGraphExecDestroy(ctx->graph_exec); ctx->graph_exec = nullptr; }
    return USRT_OK;
}

int usrt_fit_world_box(usrt_context* ctx, float out_min[3], float out_max[3]) {
    NEED_CTX(ctx);
    if (!(ctx->stage & ST_TRIS) || ctx->n == 0) return fail(ctx, USRT_ERR_STATE, "fit_world_box: no triangles uploaded");
    if (int r = bind_device(ctx)) return r;
    if (!ctx->scene_box) CU(ctx, cudaMalloc(&ctx->scene_box, 6 * sizeof(float)));
    CU(ctx, launch_scene_box(vertex_source(ctx), ctx->n, ctx->scene_box, ctx->stream));
    ctx->launches += 1;
    float b[6];
    CU(ctx, cudaMemcpyAsync(b, ctx->scene_box, sizeof(b), cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    for (int k = 0; k < 3; ++k)
        if (!(b[3 + k] > b[k])) b[3 + k] = b[k] + 1.0f;      // flat on this axis: any positive extent, every centroid maps to 0
    for (int k = 0; k < 3; ++k) {
        if (!std::isfinite(b[k]) || !std::isfinite(b[3 + k])) return fail(ctx, USRT_ERR_ARG, "fit_world_box: non-finite vertex");
        ctx->whole.min[k] = b[k]; ctx->whole.max[k] = b[3 + k];
        if (out_min) out_min[k] = b[k];
        if (out_max) out_max[k] = b[3 + k];
    }
    if (ctx->graph_exec) { cudaGraphExecDestroy(ctx->graph_exec); ctx->graph_exec = nullptr; }
    return USRT_OK;
}

uint32_t usrt_capacity(const usrt_context* ctx) { return ctx ? ctx->capacity : 0; }
uint32_t usrt_triangles_length(const usrt_context* ctx) { return ctx ? ctx->n : 0; }
uint64_t usrt_kernel_launches(const usrt_context* ctx) { return ctx ? ctx->launches : 0; }

int usrt_upload_triangles(usrt_context* ctx, const usrt_triangle* host_triangles, uint32_t n) {
    NEED_CTX(ctx);
    if (!host_triangles || n > ctx->capacity) return fail(ctx, USRT_ERR_ARG, "upload_triangles: n=%u capacity=%u", n, ctx->capacity);
    if (int r = bind_device(ctx)) return r;
    if (int r = reset_scene_buffers(ctx, n, ctx->dirty_n)) return r;
    ctx->dirty_n = n;
    CU(ctx, cudaMemcpyAsync(ctx->triangles, host_triangles, (size_t)n * sizeof(usrt_triangle), cudaMemcpyHostToDevice, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->positions_only = false;
    ctx->n = n;
    ctx->stage = ST_TRIS;
    return USRT_OK;
}

int usrt_upload_triangles_async(usrt_context* ctx, const usrt_triangle* pinned_host_triangles, uint32_t n) {
    NEED_CTX(ctx);
    if (!pinned_host_triangles || n > ctx->capacity)
        return fail(ctx, USRT_ERR_ARG, "upload_triangles_async: n=%u capacity=%u", n, ctx->capacity);
    if (int r = bind_device(ctx)) return r;
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, pinned_host_triangles) != cudaSuccess || attr.type != cudaMemoryTypeHost) {
        cudaGetLastError();
        return fail(ctx, USRT_ERR_ARG, "upload_triangles_async: host memory must be page-locked");
    }
    if (int r = reset_scene_buffers(ctx, n, ctx->dirty_n)) return r;
    ctx->dirty_n = n;
    CU(ctx, cudaMemcpyAsync(ctx->triangles, pinned_host_triangles, (size_t)n * sizeof(usrt_triangle), cudaMemcpyHostToDevice, ctx->stream));
    ctx->positions_only = false;
    ctx->n = n;
    ctx->stage = ST_TRIS;
    return USRT_OK;
}

int usrt_host_alloc(usrt_context* ctx, uint64_t bytes, void** host_ptr) {
    NEED_CTX(ctx);
    if (!host_ptr || bytes == 0) return fail(ctx, USRT_ERR_ARG, "host_alloc: null pointer or zero size");
    *host_ptr = nullptr;
    if (int r = bind_device(ctx)) return r;
    const cudaError_t e = cudaHostAlloc(host_ptr, (size_t)bytes, cudaHostAllocPortable);
    if (e != cudaSuccess) {
        cudaGetLastError();
        *host_ptr = nullptr;
        return fail(ctx, e == cudaErrorMemoryAllocation ? USRT_ERR_NOMEM : USRT_ERR_CUDA, "host_alloc(%llu bytes): %s",
                    (unsigned long long)bytes, cudaGetErrorString(e));
    }
    return USRT_OK;
}

int usrt_host_free(usrt_context* ctx, void* host_ptr) {
    NEED_CTX(ctx);
    if (!host_ptr) return USRT_OK;
    CU(ctx, cudaFreeHost(host_ptr));
    return USRT_OK;
}

int usrt_set_triangles_device(usrt_context* ctx, const void* dev_triangles, uint32_t n) {
    NEED_CTX(ctx);
    if (!dev_triangles || n > ctx->capacity) return fail(ctx, USRT_ERR_ARG, "set_triangles_device: n=%u capacity=%u", n, ctx->capacity);
    if (int r = bind_device(ctx)) return r;
    if (int r = reset_scene_buffers(ctx, n, ctx->dirty_n)) return r;
    ctx->dirty_n = n;
    CU(ctx, cudaMemcpyAsync(ctx->triangles, dev_triangles, (size_t)n * sizeof(usrt_triangle), cudaMemcpyDeviceToDevice, ctx->stream));
    ctx->positions_only = false;
    ctx->n = n;
    ctx->stage = ST_TRIS;
    return USRT_OK;
}

// Positions only: 48 bytes per triangle instead of 128. Everything the build and the traversal read of a Triangle is its
// first 48 bytes (a, b, c); uv / normals matter to the shading epilogue alone.
static int upload_positions(usrt_context* ctx, const void* host_positions, uint32_t n, bool async) {
    if (!host_positions || n > ctx->capacity) return fail(ctx, USRT_ERR_ARG, "upload_positions: n=%u capacity=%u", n, ctx->capacity);
    if (int r = bind_device(ctx)) return r;
    if (async) {
        cudaPointerAttributes attr;
        if (cudaPointerGetAttributes(&attr, host_positions) != cudaSuccess || attr.type != cudaMemoryTypeHost) {
            cudaGetLastError();
            return fail(ctx, USRT_ERR_ARG, "upload_positions_async: host memory must be page-locked");
        }
    }
    if (!ctx->positions) CU(ctx, cudaMalloc(&ctx->positions, (size_t)ctx->capacity * 3 * sizeof(float4)));
    if (int r = reset_scene_buffers(ctx, n, ctx->dirty_n)) return r;
    ctx->dirty_n = n;
    CU(ctx, cudaMemcpyAsync(ctx->positions, host_positions, (size_t)n * 48, cudaMemcpyHostToDevice, ctx->stream));
    if (!async) CU(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->positions_only = true;
    ctx->n = n;
    ctx->stage = ST_TRIS;
    return USRT_OK;
}

int usrt_upload_positions(usrt_context* ctx, const float* host_positions, uint32_t n) {
    NEED_CTX(ctx);
    return upload_positions(ctx, host_positions, n, false);
}

int usrt_upload_positions_async(usrt_context* ctx, const float* pinned_host_positions, uint32_t n) {
    NEED_CTX(ctx);
    return upload_positions(ctx, pinned_host_positions, n, true);
}

int usrt_upload_bvh(usrt_context* ctx, uint32_t n, const uint32_t* keys, const uint32_t* triangle_index,
                    const usrt_triangle* triangles, const usrt_aabb* triangle_aabb, const usrt_aabb* bvh_data,
                    const usrt_leaf_node* leaf_nodes, const usrt_internal_node* internal_nodes) {
    NEED_CTX(ctx);
    if (n < 2 || n > ctx->capacity || !keys || !triangle_index || !triangles || !triangle_aabb || !bvh_data || !leaf_nodes || !internal_nodes)
        return fail(ctx, USRT_ERR_ARG, "upload_bvh: n=%u capacity=%u or null buffer", n, ctx->capacity);
    if (int r = bind_device(ctx)) return r;
    if (int r = reset_scene_buffers(ctx, n, ctx->dirty_n)) return r;
    ctx->dirty_n = n;
    const cudaMemcpyKind k = cudaMemcpyHostToDevice;
    CU(ctx, cudaMemcpyAsync(ctx->keys, keys, (size_t)n * 4, k, ctx->stream));
    CU(ctx, cudaMemcpyAsync(ctx->tri_index, triangle_index, (size_t)n * 4, k, ctx->stream));
    CU(ctx, cudaMemcpyAsync(ctx->triangles, triangles, (size_t)n * sizeof(usrt_triangle), k, ctx->stream));
    CU(ctx, cudaMemcpyAsync(ctx->tri_aabb, triangle_aabb, (size_t)n * sizeof(usrt_aabb), k, ctx->stream));
    CU(ctx, cudaMemcpyAsync(ctx->bvh, bvh_data, (size_t)(n - 1) * sizeof(usrt_aabb), k, ctx->stream));
    CU(ctx, cudaMemcpyAsync(ctx->leaf, leaf_nodes, (size_t)n * sizeof(usrt_leaf_node), k, ctx->stream));
    CU(ctx, cudaMemcpyAsync(ctx->internal, internal_nodes, (size_t)(n - 1) * sizeof(usrt_internal_node), k, ctx->stream));
    // The file is untrusted: every index the traversal and the refit will follow is checked on the device first
    // (ranges, exactly one parent per node, one tree of depth <= 64 -- the trace stack of Raytracing.compute:133).
    ctx->n = 0;
    ctx->stage = 0;                                        // nothing usable until the tree is accepted
    uint32_t err[2] = {0, 0};
    CU(ctx, launch_import_validate(n, ctx->tri_index, ctx->internal, ctx->leaf, ctx->up_internal, ctx->up_leaf, ctx->small, ctx->stream));
    ctx->launches += 2;
    CU(ctx, cudaMemcpyAsync(err, ctx->small, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    if (err[0]) return fail(ctx, USRT_ERR_ARG, "upload_bvh: %u invalid node links / indices (out of range, or not one parent per node)", err[0]);
    CU(ctx, launch_import_links(n, ctx->internal, ctx->up_internal, ctx->up_leaf, ctx->small, ctx->stream));
    ctx->launches += 2;
    CU(ctx, cudaMemcpyAsync(err, ctx->small, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    if (err[0]) return fail(ctx, USRT_ERR_ARG, "upload_bvh: %u leaves do not reach node 0 within 64 levels (cycle, forest, or too deep)", err[0]);
    CU(ctx, launch_pack_traversal(n, ctx->tri_index, ctx->tri_aabb, ctx->triangles, ctx->internal, ctx->leaf, ctx->bvh,
                                  ctx->packed_nodes, ctx->packed_tris, ctx->stream));
    ctx->launches += 1;
    ctx->positions_only = false;
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->n = n;
    // ConstructBVH refits by leaf slot (BVH.compute:199-208): allowed again on this tree only if every leaf sits in its
    // own slot, as every tree of TreeConstructor does (:114-118); the K4 -> K5 links were rebuilt above.
    ctx->stage = ST_TRIS | ST_MORTON | ST_SORTED | ST_DISTRIBUTED | ST_BVH | (err[1] == 0 ? ST_TREE : 0u);
    return USRT_OK;
}

int usrt_morton(usrt_context* ctx) {
    NEED_CTX(ctx);
    if (!(ctx->stage & ST_TRIS)) return fail(ctx, USRT_ERR_STATE, "morton: no triangles uploaded");
    if (int r = bind_device(ctx)) return r;
    return do_morton(ctx);
}

int usrt_sort(usrt_context* ctx) {
    NEED_CTX(ctx);
    if (!(ctx->stage & ST_MORTON)) return fail(ctx, USRT_ERR_STATE, "sort: keys not generated (call usrt_morton)");
    if (int r = bind_device(ctx)) return r;
    return do_sort(ctx);
}

int usrt_sort_pairs_device(usrt_context* ctx, uint32_t* dev_keys, uint32_t* dev_values, uint64_t count) {
    NEED_CTX(ctx);
    if (count && !dev_keys) return fail(ctx, USRT_ERR_ARG, "sort_pairs: null keys");
    if (count >= (1ull << 32)) return fail(ctx, USRT_ERR_ARG, "sort_pairs: count must be < 2^32");
    if (int r = bind_device(ctx)) return r;
    CU(ctx, sort_scratch_reserve(ctx->sort, std::max<uint64_t>(count, 1), true));
    ctx->sort_ev_valid = false;
    CU(ctx, sort_pairs(dev_keys, dev_values, ctx->sort.keys_alt, dev_values ? ctx->sort.vals_alt : nullptr, count,
                       ctx->sort, ctx->stream, &ctx->launches, ctx->timing ? ctx->sort_ev : nullptr));
    ctx->sort_ev_valid = ctx->timing && count > 0;
    return USRT_OK;
}

int usrt_sort_pairs_host(usrt_context* ctx, uint32_t* host_keys, uint32_t* host_values, uint64_t count) {
    NEED_CTX(ctx);
    if (count == 0) return USRT_OK;
    if (!host_keys) return fail(ctx, USRT_ERR_ARG, "sort_pairs_host: null keys");
    if (count >= (1ull << 32)) return fail(ctx, USRT_ERR_ARG, "sort_pairs: count must be < 2^32");
    if (int r = bind_device(ctx)) return r;
    uint32_t *dk = nullptr, *dv = nullptr;
    CU(ctx, cudaMalloc(&dk, count * 4));
    if (host_values) {
        cudaError_t e = cudaMalloc(&dv, count * 4);
        if (e != cudaSuccess) { cudaFree(dk); return fail(ctx, USRT_ERR_NOMEM, "sort_pairs_host: %s", cudaGetErrorString(e)); }
    }
    int rc = USRT_OK;
    auto run = [&]() -> int {
        CU(ctx, cudaMemcpyAsync(dk, host_keys, count * 4, cudaMemcpyHostToDevice, ctx->stream));
        if (dv) CU(ctx, cudaMemcpyAsync(dv, host_values, count * 4, cudaMemcpyHostToDevice, ctx->stream));
        if (int r = usrt_sort_pairs_device(ctx, dk, dv, count)) return r;
        CU(ctx, cudaMemcpyAsync(host_keys, dk, count * 4, cudaMemcpyDeviceToHost, ctx->stream));
        if (dv) CU(ctx, cudaMemcpyAsync(host_values, dv, count * 4, cudaMemcpyDeviceToHost, ctx->stream));
        CU(ctx, cudaStreamSynchronize(ctx->stream));
        return USRT_OK;
    };
    rc = run();
    cudaFree(dk);
    if (dv) cudaFree(dv);
    return rc;
}

int usrt_partition_pass_device(usrt_context* ctx, const uint32_t* src_keys, const uint32_t* src_values,
                               uint32_t* dst_keys, uint32_t* dst_values, uint64_t count, int bit_offset,
                               uint32_t* histogram_out) {
    NEED_CTX(ctx);
    if (bit_offset < 0 || bit_offset > 24 || (bit_offset & 7)) return fail(ctx, USRT_ERR_ARG, "partition_pass: bit_offset must be 0, 8, 16 or 24");
    if (count && (!src_keys || !dst_keys)) return fail(ctx, USRT_ERR_ARG, "partition_pass: null keys");
    if (count >= (1ull << 32)) return fail(ctx, USRT_ERR_ARG, "partition_pass: count must be < 2^32");
    if (int r = bind_device(ctx)) return r;
    CU(ctx, partition_pass(src_keys, src_values, dst_keys, dst_values, count, bit_offset, histogram_out, ctx->sort,
                           ctx->stream, &ctx->launches));
    return USRT_OK;
}

int usrt_digit_histogram_device(usrt_context* ctx, const uint32_t* dev_keys, uint64_t count, int bit_offset, uint32_t* dev_hist_out) {
    NEED_CTX(ctx);
    if (bit_offset < 0 || bit_offset > 24 || (bit_offset & 7) || !dev_hist_out || (count && !dev_keys) || count >= (1ull << 32))
        return fail(ctx, USRT_ERR_ARG, "digit_histogram: bad arguments");
    if (int r = bind_device(ctx)) return r;
    CU(ctx, digit_histogram(dev_keys, count, bit_offset, dev_hist_out, ctx->sort, ctx->stream, &ctx->launches));
    return USRT_OK;
}

int usrt_partition_scatter_device(usrt_context* ctx, const uint32_t* src_keys, const uint32_t* src_values, uint64_t count,
                                  int bit_offset, const uint64_t* dev_key_base, const uint64_t* dev_value_base) {
    NEED_CTX(ctx);
    if (bit_offset < 0 || bit_offset > 24 || (bit_offset & 7) || !dev_key_base || !dev_value_base ||
        (count && (!src_keys || !src_values)) || count >= (1ull << 30))
        return fail(ctx, USRT_ERR_ARG, "partition_scatter: bad arguments");
    if (int r = bind_device(ctx)) return r;
    CU(ctx, partition_scatter(src_keys, src_values, count, bit_offset, reinterpret_cast<const unsigned long long*>(dev_key_base),
                              reinterpret_cast<const unsigned long long*>(dev_value_base), ctx->sort, ctx->stream, &ctx->launches));
    return USRT_OK;
}

int usrt_peer_scatter_plan_device(usrt_context* ctx, const uint32_t* dev_all_hist, int world, int rank, const uint64_t* dev_peer_base,
                                  uint64_t capacity, uint64_t* dev_key_base, uint64_t* dev_value_base, uint64_t* dev_recv_total,
                                  uint32_t* dev_bounds) {
    NEED_CTX(ctx);
    if (!dev_all_hist || !dev_peer_base || !dev_key_base || !dev_value_base || !dev_recv_total || world < 1 || world > 16 || rank < 0 ||
        rank >= world)
        return fail(ctx, USRT_ERR_ARG, "peer_scatter_plan: bad arguments (world 1..16)");
    if (int r = bind_device(ctx)) return r;
    CU(ctx, peer_scatter_plan(dev_all_hist, world, rank, reinterpret_cast<const unsigned long long*>(dev_peer_base), capacity,
                              reinterpret_cast<unsigned long long*>(dev_key_base), reinterpret_cast<unsigned long long*>(dev_value_base),
                              reinterpret_cast<unsigned long long*>(dev_recv_total), dev_bounds, ctx->stream, &ctx->launches));
    return USRT_OK;
}

int usrt_peer_buffer_create(usrt_context* ctx, uint64_t bytes, void** dev_ptr, unsigned char handle_out[64]) {
    NEED_CTX(ctx);
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    if (!dev_ptr || !handle_out || bytes == 0) return fail(ctx, USRT_ERR_ARG, "peer_buffer_create: bad arguments");
    if (int r = bind_device(ctx)) return r;
    void* p = nullptr;
    CU(ctx, cudaMalloc(&p, bytes));
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) { cudaFree(p); return fail(ctx, USRT_ERR_CUDA, "cudaIpcGetMemHandle: %s", cudaGetErrorString(e)); }
    memcpy(handle_out, &h, 64);
    *dev_ptr = p;
    return USRT_OK;
}

int usrt_peer_buffer_open(usrt_context* ctx, const unsigned char handle[64], void** dev_ptr) {
    NEED_CTX(ctx);
    if (!handle || !dev_ptr) return fail(ctx, USRT_ERR_ARG, "peer_buffer_open: bad arguments");
    if (int r = bind_device(ctx)) return r;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    CU(ctx, cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return USRT_OK;
}

int usrt_peer_buffer_close(usrt_context* ctx, void* dev_ptr, int opened) {
    NEED_CTX(ctx);
    if (!dev_ptr) return USRT_OK;
    if (int r = bind_device(ctx)) return r;
    if (opened) CU(ctx, cudaIpcCloseMemHandle(dev_ptr));
    else CU(ctx, cudaFree(dev_ptr));
    return USRT_OK;
}

int usrt_set_key_mode(usrt_context* ctx, int mode) {
    NEED_CTX(ctx);
    if (mode != USRT_KEYS_REFERENCE && mode != USRT_KEYS_INDEX_TIEBREAK && mode != USRT_KEYS_MORTON64)
        return fail(ctx, USRT_ERR_ARG, "set_key_mode: unknown mode %d", mode);
    if (int r = bind_device(ctx)) return r;
    if (mode == USRT_KEYS_MORTON64 && !ctx->keys64) {
        const size_t c = ctx->capacity;
        CU(ctx, cudaMalloc(&ctx->keys64, c * 8));
        CU(ctx, cudaMalloc(&ctx->keys64_alt, c * 8));
        CU(ctx, cudaMalloc(&ctx->scan_status64, distribute_status_bytes64(ctx->capacity)));
        ctx->keys64_primary = ctx->keys64;
        CU(ctx, cudaMemsetAsync(ctx->keys64, 0xFF, c * 8, ctx->stream));
        CU(ctx, cudaMemsetAsync(ctx->keys64_alt, 0xFF, c * 8, ctx->stream));
        CU(ctx, sort_scratch_reserve64(ctx->sort, ctx->capacity, false));
    }
    if (mode != ctx->key_mode) {
        ctx->key_mode = mode;
        ctx->stage &= ST_TRIS;                            // keys, tree and boxes of the other mode are void
        if (ctx->graph_exec) { cudaGraphExecDestroy(ctx->graph_exec); ctx->graph_exec = nullptr; }
    }
    return USRT_OK;
}

int usrt_sort_pairs64_device(usrt_context* ctx, uint64_t* dev_keys, uint32_t* dev_values, uint64_t count) {
    NEED_CTX(ctx);
    if (count && !dev_keys) return fail(ctx, USRT_ERR_ARG, "sort_pairs64: null keys");
    if (count >= (1ull << 32)) return fail(ctx, USRT_ERR_ARG, "sort_pairs64: count must be < 2^32");
    if (int r = bind_device(ctx)) return r;
    CU(ctx, sort_scratch_reserve64(ctx->sort, std::max<uint64_t>(count, 1), true));
    ctx->sort_ev_valid = false;
    CU(ctx, sort_pairs64(dev_keys, dev_values, ctx->sort.keys64_alt, dev_values ? ctx->sort.vals64_alt : nullptr, count, ctx->sort,
                         ctx->stream, &ctx->launches));
    return USRT_OK;
}

int usrt_sort_pairs64_host(usrt_context* ctx, uint64_t* host_keys, uint32_t* host_values, uint64_t count) {
    NEED_CTX(ctx);
    if (count == 0) return USRT_OK;
    if (!host_keys) return fail(ctx, USRT_ERR_ARG, "sort_pairs64_host: null keys");
    if (count >= (1ull << 32)) return fail(ctx, USRT_ERR_ARG, "sort_pairs64: count must be < 2^32");
    if (int r = bind_device(ctx)) return r;
    uint64_t* dk = nullptr;
    uint32_t* dv = nullptr;
    CU(ctx, cudaMalloc(&dk, count * 8));
    if (host_values) {
        cudaError_t e = cudaMalloc(&dv, count * 4);
        if (e != cudaSuccess) { cudaFree(dk); return fail(ctx, USRT_ERR_NOMEM, "sort_pairs64_host: %s", cudaGetErrorString(e)); }
    }
    auto run = [&]() -> int {
        CU(ctx, cudaMemcpyAsync(dk, host_keys, count * 8, cudaMemcpyHostToDevice, ctx->stream));
        if (dv) CU(ctx, cudaMemcpyAsync(dv, host_values, count * 4, cudaMemcpyHostToDevice, ctx->stream));
        if (int r = usrt_sort_pairs64_device(ctx, dk, dv, count)) return r;
        CU(ctx, cudaMemcpyAsync(host_keys, dk, count * 8, cudaMemcpyDeviceToHost, ctx->stream));
        if (dv) CU(ctx, cudaMemcpyAsync(host_values, dv, count * 4, cudaMemcpyDeviceToHost, ctx->stream));
        CU(ctx, cudaStreamSynchronize(ctx->stream));
        return USRT_OK;
    };
    const int rc = run();
    cudaFree(dk);
    if (dv) cudaFree(dv);
    return rc;
}

int usrt_distribute_keys(usrt_context* ctx) {
    NEED_CTX(ctx);
    if (!(ctx->stage & ST_SORTED)) return fail(ctx, USRT_ERR_STATE, "distribute_keys: keys not sorted (call usrt_sort)");
    if (ctx->stage & ST_DISTRIBUTED) return fail(ctx, USRT_ERR_STATE, "distribute_keys: already applied to these keys");
    if (ctx->key_mode == USRT_KEYS_INDEX_TIEBREAK)
        return fail(ctx, USRT_ERR_STATE, "distribute_keys: key mode USRT_KEYS_INDEX_TIEBREAK builds the tree on the raw sorted codes");
    if (int r = bind_device(ctx)) return r;
    return do_distribute(ctx);
}

int usrt_construct_tree(usrt_context* ctx) {
    NEED_CTX(ctx);
    if (!(ctx->stage & ST_SORTED)) return fail(ctx, USRT_ERR_STATE, "construct_tree: keys not sorted");
    if (ctx->n < 2) return fail(ctx, USRT_ERR_ARG, "construct_tree: trianglesCount=%u; the reference needs >= 2 (BVH.compute:101)", ctx->n);
    if (int r = bind_device(ctx)) return r;
    return do_tree(ctx);
}

int usrt_construct_bvh(usrt_context* ctx) {
    NEED_CTX(ctx);
    if (!(ctx->stage & ST_TREE)) return fail(ctx, USRT_ERR_STATE, "construct_bvh: tree not built (call usrt_construct_tree)");
    if (int r = bind_device(ctx)) return r;
    return do_bvh(ctx);
}

namespace {

// The five stages enqueued back to back (RaytracingMeshDrawer.cs:34-51 without its readbacks).
int enqueue_rebuild(usrt_context* ctx, bool timed) {
    if (timed) CU(ctx, cudaEventRecord(ctx->ev[0], ctx->stream));
    if (int r = do_morton(ctx)) return r;
    if (timed) CU(ctx, cudaEventRecord(ctx->ev[1], ctx->stream));
    if (int r = do_sort(ctx)) return r;
    if (timed) CU(ctx, cudaEventRecord(ctx->ev[2], ctx->stream));
    if (ctx->key_mode != USRT_KEYS_INDEX_TIEBREAK)
        if (int r = do_distribute(ctx)) return r;
    if (timed) CU(ctx, cudaEventRecord(ctx->ev[3], ctx->stream));
    if (int r = do_tree(ctx)) return r;
    if (timed) CU(ctx, cudaEventRecord(ctx->ev[4], ctx->stream));
    if (int r = do_bvh(ctx)) return r;
    if (timed) CU(ctx, cudaEventRecord(ctx->ev[5], ctx->stream));
    return USRT_OK;
}

void drop_rebuild_graph(usrt_context* ctx) {
    if (ctx->graph_exec) cudaGraphExecDestroy(ctx->graph_exec);
    ctx->graph_exec = nullptr;
    ctx->graph_n = 0;
}

}  // namespace

int usrt_rebuild(usrt_context* ctx) {
    NEED_CTX(ctx);
    if (!(ctx->stage & ST_TRIS)) return fail(ctx, USRT_ERR_STATE, "rebuild: no triangles uploaded");
    if (ctx->n < 2) return fail(ctx, USRT_ERR_ARG, "rebuild: trianglesCount=%u; the reference needs >= 2 (BVH.compute:101)", ctx->n);
    if (int r = bind_device(ctx)) return r;
    ctx->ev_valid = false;
    // DistributeKeys leaves keys/keys_alt swapped; start every rebuild from the same orientation so the
    // launch sequence (and therefore a captured graph) is identical from one rebuild to the next.
    if (ctx->keys != ctx->keys_primary) std::swap(ctx->keys, ctx->keys_alt);
    if (ctx->keys64 && ctx->keys64 != ctx->keys64_primary) std::swap(ctx->keys64, ctx->keys64_alt);
    if (ctx->key_mode != USRT_KEYS_REFERENCE) {          // the variants are enqueued launch by launch (no graph, no stage events)
        if (int r = enqueue_rebuild(ctx, false)) return r;
        return USRT_OK;
    }

    if (ctx->timing || !ctx->use_graph) {            // per-stage events are recorded between launches: no graph
        if (int r = enqueue_rebuild(ctx, ctx->timing)) return r;
        ctx->ev_valid = ctx->timing;
        return USRT_OK;
    }
    // The 12 launches + 4 memsets of a rebuild are short (the whole 1M-triangle rebuild is ~0.28 ms): replay them
    // as one CUDA graph so the gaps between them are not paid on every rebuild. Re-captured when n, the
    // stream, or the world box changes.
    // (the captured launches hold raw pointers into the sort scratch: a standalone sort of more pairs than this
    // context ever saw re-allocates it, which bumps sort.generation and forces a re-capture here)
    CU(ctx, sort_scratch_reserve(ctx->sort, ctx->n, false));              // no allocation while capturing
    if (!ctx->graph_exec || ctx->graph_n != ctx->n || ctx->graph_stream != ctx->stream ||
        ctx->graph_sort_generation != ctx->sort.generation || ctx->graph_positions_only != ctx->positions_only) {
        drop_rebuild_graph(ctx);
        const uint64_t before = ctx->launches;
        cudaGraph_t graph = nullptr;
        if (cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeRelaxed) != cudaSuccess) {
            cudaGetLastError();                                            // e.g. the legacy default stream cannot be captured
            ctx->use_graph = false;
            return enqueue_rebuild(ctx, false);
        }
        const int rc = enqueue_rebuild(ctx, false);
        const cudaError_t ce = cudaStreamEndCapture(ctx->stream, &graph);
        if (rc != USRT_OK || ce != cudaSuccess || !graph) {
            if (graph) cudaGraphDestroy(graph);
            cudaGetLastError();
            ctx->use_graph = false;                                        // fall back to plain launches for good
            if (ctx->keys != ctx->keys_primary) std::swap(ctx->keys, ctx->keys_alt);
            ctx->launches = before;
            if (int r = enqueue_rebuild(ctx, false)) return r;
            return USRT_OK;
        }
        ctx->graph_launches = ctx->launches - before;
        ctx->launches = before;
        const cudaError_t ie = cudaGraphInstantiate(&ctx->graph_exec, graph, 0);
        cudaGraphDestroy(graph);
        if (ie != cudaSuccess) return fail(ctx, USRT_ERR_CUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(ie));
        ctx->graph_n = ctx->n;
        ctx->graph_stream = ctx->stream;
        ctx->graph_sort_generation = ctx->sort.generation;
        ctx->graph_positions_only = ctx->positions_only;
        if (ctx->keys != ctx->keys_primary) std::swap(ctx->keys, ctx->keys_alt);   // capture advanced the host-side state
    }
    CU(ctx, cudaGraphLaunch(ctx->graph_exec, ctx->stream));
    ctx->launches += ctx->graph_launches;
    std::swap(ctx->keys, ctx->keys_alt);                                  // as do_distribute does
    ctx->sort_ev_valid = false;
    ctx->stage = ST_TRIS | ST_MORTON | ST_SORTED | ST_DISTRIBUTED | ST_TREE | ST_BVH;
    return USRT_OK;
}

int usrt_enable_stage_timing(usrt_context* ctx, int enabled) {
    NEED_CTX(ctx);
    ctx->timing = enabled != 0;
    return USRT_OK;
}

int usrt_last_rebuild_ms(usrt_context* ctx, float out_ms[6]) {
    NEED_CTX(ctx);
    if (!out_ms) return USRT_ERR_ARG;
    if (!ctx->ev_valid) return fail(ctx, USRT_ERR_STATE, "last_rebuild_ms: no timed rebuild");
    if (int r = bind_device(ctx)) return r;
    CU(ctx, cudaEventSynchronize(ctx->ev[5]));
    for (int i = 0; i < 5; ++i) CU(ctx, cudaEventElapsedTime(&out_ms[i], ctx->ev[i], ctx->ev[i + 1]));
    CU(ctx, cudaEventElapsedTime(&out_ms[5], ctx->ev[0], ctx->ev[5]));
    return USRT_OK;
}

int usrt_last_sort_ms(usrt_context* ctx, float out_ms[6]) {
    NEED_CTX(ctx);
    if (!out_ms) return USRT_ERR_ARG;
    if (!ctx->sort_ev_valid) return fail(ctx, USRT_ERR_STATE, "last_sort_ms: no timed sort");
    if (int r = bind_device(ctx)) return r;
    CU(ctx, cudaEventSynchronize(ctx->sort_ev[5]));
    for (int i = 0; i < 5; ++i) CU(ctx, cudaEventElapsedTime(&out_ms[i], ctx->sort_ev[i], ctx->sort_ev[i + 1]));
    CU(ctx, cudaEventElapsedTime(&out_ms[5], ctx->sort_ev[0], ctx->sort_ev[5]));
    return USRT_OK;
}

int usrt_set_trace_mode(usrt_context* ctx, int mode) {
    NEED_CTX(ctx);
    if (mode < 0 || mode > 2) return fail(ctx, USRT_ERR_ARG, "trace mode must be 0 (strict), 1 (culled) or 2 (culled, near child first)");
    ctx->trace_mode = mode;
    return USRT_OK;
}

int usrt_set_hit_mirrors(usrt_context* ctx, int count, void* const* dev_ptrs) {
    NEED_CTX(ctx);
    if (count < 0 || count > kMaxHitMirrors - 1 || (count && !dev_ptrs))
        return fail(ctx, USRT_ERR_ARG, "set_hit_mirrors: count must be 0..%d", kMaxHitMirrors - 1);
    for (int i = 0; i < count; ++i)
        if (!dev_ptrs[i]) return fail(ctx, USRT_ERR_ARG, "set_hit_mirrors: null mirror %d", i);
    ctx->mirrors = HitMirrors{};
    for (int i = 0; i < count; ++i) ctx->mirrors.ptr[i] = static_cast<usrt_raycast_result*>(dev_ptrs[i]);
    ctx->mirrors.count = count;
    return USRT_OK;
}

int usrt_trace_primary(usrt_context* ctx, int width, int height, float near_plane, float tan_half_fov,
                       const float camera_to_world[16], int y0, int y1, usrt_raycast_result* host_out) {
    NEED_CTX(ctx);
    if (!(ctx->stage & ST_BVH)) return fail(ctx, USRT_ERR_STATE, "trace: BVH not built");
    if (width <= 0 || height <= 0 || !camera_to_world || y0 < 0 || y1 > height || y0 > y1)
        return fail(ctx, USRT_ERR_ARG, "trace_primary: bad frame %dx%d rows [%d,%d)", width, height, y0, y1);
    if (int r = bind_device(ctx)) return r;
    const uint64_t count = (uint64_t)width * (uint64_t)height;
    if (int r = ensure_hits(ctx, count)) return r;
    ctx->hits_count = count;
    PrimaryParams p;
    p.width = width; p.height = height; p.near_plane = near_plane; p.tan_half_fov = tan_half_fov;
    memcpy(p.m, camera_to_world, sizeof(p.m));
    p.y0 = y0; p.y1 = y1;
    p.block_rows = 1; p.shard = 0; p.num_shards = 0; p.local_rows = 0;
    TraceScene s{ctx->packed_nodes, ctx->packed_tris, ctx->bvh};
    const int rows = y1 - y0;
    if (!host_out || rows <= 0) {
        CU(ctx, launch_trace_primary(s, p, ctx->hits, ctx->trace_mode, ctx->stream, ctx->mirrors));
        ctx->launches += rows > 0 ? 1 : 0;
        return USRT_OK;
    }
    // Host readback. If host_out is page-locked (cudaHostAlloc / cudaHostRegister / torch pin_memory) the
    // kernel writes each hit record to it directly as well (posted PCIe writes, 128-B runs per warp
    // row), so the 16 B/ray transfer overlaps the traversal instead of following it. Pageable
    // memory takes a staged copy after the kernel.
    cudaPointerAttributes attr;
    usrt_raycast_result* alias = nullptr;
    if (cudaPointerGetAttributes(&attr, host_out) == cudaSuccess && attr.type == cudaMemoryTypeHost && attr.devicePointer)
        alias = static_cast<usrt_raycast_result*>(attr.devicePointer);
    else
        cudaGetLastError();                               // clear the "not a CUDA pointer" status
    HitMirrors mirrors = ctx->mirrors;
    if (alias) mirrors.ptr[mirrors.count++] = alias;      // count <= kMaxHitMirrors - 1 by usrt_set_hit_mirrors
    CU(ctx, launch_trace_primary(s, p, ctx->hits, ctx->trace_mode, ctx->stream, mirrors));
    ctx->launches += 1;
    if (!alias) {
        const size_t off = (size_t)y0 * width;
        CU(ctx, cudaMemcpyAsync(host_out + off, ctx->hits + off, (size_t)rows * width * sizeof(usrt_raycast_result),
                                cudaMemcpyDeviceToHost, ctx->stream));
    }
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return USRT_OK;
}

int usrt_trace_primary_async(usrt_context* ctx, int width, int height, float near_plane, float tan_half_fov,
                             const float camera_to_world[16], usrt_raycast_result* pinned_host_out) {
    NEED_CTX(ctx);
    if (!(ctx->stage & ST_BVH)) return fail(ctx, USRT_ERR_STATE, "trace: BVH not built");
    if (width <= 0 || height <= 0 || !camera_to_world || !pinned_host_out)
        return fail(ctx, USRT_ERR_ARG, "trace_primary_async: bad frame %dx%d or null output", width, height);
    if (int r = bind_device(ctx)) return r;
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, pinned_host_out) != cudaSuccess || attr.type != cudaMemoryTypeHost || !attr.devicePointer) {
        cudaGetLastError();
        return fail(ctx, USRT_ERR_ARG, "trace_primary_async: host memory must be page-locked");
    }
    const uint64_t count = (uint64_t)width * (uint64_t)height;
    if (int r = ensure_hits(ctx, count)) return r;
    ctx->hits_count = count;
    PrimaryParams p;
    p.width = width; p.height = height; p.near_plane = near_plane; p.tan_half_fov = tan_half_fov;
    memcpy(p.m, camera_to_world, sizeof(p.m));
    p.y0 = 0; p.y1 = height;
    p.block_rows = 1; p.shard = 0; p.num_shards = 0; p.local_rows = 0;
    TraceScene s{ctx->packed_nodes, ctx->packed_tris, ctx->bvh};
    HitMirrors mirrors = ctx->mirrors;
    mirrors.ptr[mirrors.count++] = static_cast<usrt_raycast_result*>(attr.devicePointer);
    CU(ctx, launch_trace_primary(s, p, ctx->hits, ctx->trace_mode, ctx->stream, mirrors));
    ctx->launches += 1;
    return USRT_OK;
}

int usrt_trace_primary_sharded(usrt_context* ctx, int width, int height, float near_plane, float tan_half_fov,
                               const float camera_to_world[16], int block_rows, int shard, int num_shards,
                               void* dev_out, usrt_raycast_result* host_out) {
    NEED_CTX(ctx);
    if (!(ctx->stage & ST_BVH)) return fail(ctx, USRT_ERR_STATE, "trace: BVH not built");
    if (width <= 0 || height <= 0 || !camera_to_world || block_rows <= 0 || num_shards <= 0 || shard < 0 || shard >= num_shards)
        return fail(ctx, USRT_ERR_ARG, "trace_primary_sharded: bad frame %dx%d block_rows=%d shard %d/%d", width, height,
                    block_rows, shard, num_shards);
    if (int r = bind_device(ctx)) return r;
    const int blocks_total = (height + block_rows - 1) / block_rows;
    const int blocks_per_shard = (blocks_total + num_shards - 1) / num_shards;      // padded: equal on every shard
    const int local_rows = blocks_per_shard * block_rows;
    const uint64_t count = (uint64_t)local_rows * (uint64_t)width;
    usrt_raycast_result* out = static_cast<usrt_raycast_result*>(dev_out);
    if (!out) {
        if (int r = ensure_hits(ctx, count)) return r;
        out = ctx->hits;
        ctx->hits_count = count;
    }
    // rows past the frame (padding of the last block / last shard) are never traced: pre-fill them with MISS records
    // {MAX_FLOAT, 0, (0,0)} (Raytracing.compute:129-131) -- only the last local block can hold such rows, y grows with
    // the local row
    const uint64_t last_block = (uint64_t)(local_rows - block_rows) * (uint64_t)width;
    CU(ctx, launch_fill_miss(out + last_block, count - last_block, ctx->stream));
    ctx->launches += 1;
    for (int i = 0; i < ctx->mirrors.count; ++i) {
        CU(ctx, launch_fill_miss(ctx->mirrors.ptr[i] + last_block, count - last_block, ctx->stream));
        ctx->launches += 1;
    }
    PrimaryParams p;
    p.width = width; p.height = height; p.near_plane = near_plane; p.tan_half_fov = tan_half_fov;
    memcpy(p.m, camera_to_world, sizeof(p.m));
    p.y0 = 0; p.y1 = 0;
    p.block_rows = block_rows; p.shard = shard; p.num_shards = num_shards; p.local_rows = local_rows;
    TraceScene s{ctx->packed_nodes, ctx->packed_tris, ctx->bvh};
    CU(ctx, launch_trace_primary(s, p, out, ctx->trace_mode, ctx->stream, ctx->mirrors));
    ctx->launches += 1;
    if (host_out) {
        CU(ctx, cudaMemcpyAsync(host_out, out, count * sizeof(usrt_raycast_result), cudaMemcpyDeviceToHost, ctx->stream));
        CU(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return USRT_OK;
}

int usrt_diffuse_rays_device(usrt_context* ctx, int width, int height, float near_plane, float tan_half_fov,
                             const float camera_to_world[16], const void* dev_primary_hits, uint64_t seed,
                             uint32_t first_sample, uint32_t num_samples, void* dev_rays_out) {
    NEED_CTX(ctx);
    if (!(ctx->stage & ST_TRIS)) return fail(ctx, USRT_ERR_STATE, "diffuse_rays: no triangles uploaded");
    if (width <= 0 || height <= 0 || !camera_to_world || !dev_rays_out || (uint64_t)width * height > 0xFFFFFFFFull ||
        (uint64_t)first_sample + num_samples > 65536)
        return fail(ctx, USRT_ERR_ARG, "diffuse_rays: bad frame %dx%d, samples [%u,+%u) (at most 65536) or null output", width, height,
                    first_sample, num_samples);
    const usrt_raycast_result* hits = static_cast<const usrt_raycast_result*>(dev_primary_hits);
    if (!hits) {
        if (ctx->hits_count != (uint64_t)width * height) return fail(ctx, USRT_ERR_STATE, "diffuse_rays: the last trace was not this frame");
        hits = ctx->hits;
    }
    if (int r = bind_device(ctx)) return r;
    PrimaryParams p;
    p.width = width; p.height = height; p.near_plane = near_plane; p.tan_half_fov = tan_half_fov;
    memcpy(p.m, camera_to_world, sizeof(p.m));
    p.y0 = 0; p.y1 = height;
    p.block_rows = 1; p.shard = 0; p.num_shards = 0; p.local_rows = 0;
    CU(ctx, launch_diffuse_rays(p, hits, vertex_source(ctx), seed, first_sample, num_samples, static_cast<float4*>(dev_rays_out), ctx->stream));
    ctx->launches += num_samples ? 1 : 0;
    return USRT_OK;
}

int usrt_trace_rays_device(usrt_context* ctx, const void* dev_rays, uint64_t num_rays, void* dev_out) {
    NEED_CTX(ctx);
    if (!(ctx->stage & ST_BVH)) return fail(ctx, USRT_ERR_STATE, "trace: BVH not built");
    if (num_rays && !dev_rays) return fail(ctx, USRT_ERR_ARG, "trace_rays: null rays");
    if (int r = bind_device(ctx)) return r;
    usrt_raycast_result* out = static_cast<usrt_raycast_result*>(dev_out);
    if (!out) {
        if (int r = ensure_hits(ctx, num_rays)) return r;
        out = ctx->hits;
        ctx->hits_count = num_rays;
    }
    TraceScene s{ctx->packed_nodes, ctx->packed_tris, ctx->bvh};
    CU(ctx, launch_trace_rays(s, static_cast<const float4*>(dev_rays), num_rays, out, ctx->trace_mode, ctx->stream));
    ctx->launches += num_rays ? 1 : 0;
    return USRT_OK;
}

int usrt_trace_rays(usrt_context* ctx, const float* host_rays, uint64_t num_rays, usrt_raycast_result* host_out) {
    NEED_CTX(ctx);
    if (num_rays == 0) return USRT_OK;
    if (!host_rays) return fail(ctx, USRT_ERR_ARG, "trace_rays: null rays");
    if (!(ctx->stage & ST_BVH)) return fail(ctx, USRT_ERR_STATE, "trace: BVH not built");
    if (int r = bind_device(ctx)) return r;
    if (int r = ensure_rays(ctx, num_rays)) return r;
    CU(ctx, cudaMemcpyAsync(ctx->rays, host_rays, num_rays * 32, cudaMemcpyHostToDevice, ctx->stream));
    if (int r = usrt_trace_rays_device(ctx, ctx->rays, num_rays, nullptr)) return r;
    if (host_out) {
        CU(ctx, cudaMemcpyAsync(host_out, ctx->hits, num_rays * sizeof(usrt_raycast_result), cudaMemcpyDeviceToHost, ctx->stream));
    }
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return USRT_OK;
}

int usrt_upload_texture(usrt_context* ctx, const float* host_rgba, int width, int height) {
    NEED_CTX(ctx);
    if (!host_rgba || width <= 0 || height <= 0) return fail(ctx, USRT_ERR_ARG, "upload_texture: bad texture %dx%d", width, height);
    if (int r = bind_device(ctx)) return r;
    if (ctx->texture) CU(ctx, cudaFree(ctx->texture));
    ctx->texture = nullptr;
    const size_t bytes = (size_t)width * height * sizeof(float4);
    CU(ctx, cudaMalloc(&ctx->texture, bytes));
    CU(ctx, cudaMemcpyAsync(ctx->texture, host_rgba, bytes, cudaMemcpyHostToDevice, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->tex_w = width; ctx->tex_h = height;
    return USRT_OK;
}

int usrt_shade(usrt_context* ctx, void* dev_out, uint16_t* host_out) {
    NEED_CTX(ctx);
    if (!ctx->texture) return fail(ctx, USRT_ERR_STATE, "shade: no texture uploaded");
    if (!ctx->hits || ctx->hits_count == 0) return fail(ctx, USRT_ERR_STATE, "shade: no hit records (trace first)");
    if (int r = bind_device(ctx)) return r;
    const uint64_t count = ctx->hits_count;
    void* out = dev_out;
    if (!out) {
        if (count > ctx->shaded_capacity) {
            if (ctx->shaded) CU(ctx, cudaFree(ctx->shaded));
            ctx->shaded = nullptr; ctx->shaded_capacity = 0;
            CU(ctx, cudaMalloc(&ctx->shaded, count * 8));
            ctx->shaded_capacity = count;
        }
        out = ctx->shaded;
    }
    CU(ctx, launch_shade(ctx->hits, count, ctx->triangles, ctx->texture, ctx->tex_w, ctx->tex_h, out, ctx->stream));
    ctx->launches += 1;
    if (host_out) {
        CU(ctx, cudaMemcpyAsync(host_out, out, count * 8, cudaMemcpyDeviceToHost, ctx->stream));
        CU(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return USRT_OK;
}

int usrt_hits_device(usrt_context* ctx, void** dev_ptr, uint64_t* count) {
    NEED_CTX(ctx);
    if (dev_ptr) *dev_ptr = ctx->hits;
    if (count) *count = ctx->hits_count;
    return USRT_OK;
}

static int buffer_info(usrt_context* ctx, int buffer, void** ptr, size_t* elem) {
    switch (buffer) {
        case USRT_BUF_KEYS:
            if (ctx->key_mode == USRT_KEYS_MORTON64) return fail(ctx, USRT_ERR_STATE, "key mode USRT_KEYS_MORTON64: the keys are USRT_BUF_KEYS64");
            *ptr = ctx->keys; *elem = 4; return USRT_OK;
        case USRT_BUF_KEYS64:
            if (ctx->key_mode != USRT_KEYS_MORTON64) return fail(ctx, USRT_ERR_STATE, "USRT_BUF_KEYS64 exists in key mode USRT_KEYS_MORTON64 only");
            *ptr = ctx->keys64; *elem = 8; return USRT_OK;
        case USRT_BUF_TRIANGLE_INDEX: *ptr = ctx->tri_index; *elem = 4; return USRT_OK;
        case USRT_BUF_TRIANGLE_DATA: *ptr = ctx->triangles; *elem = sizeof(usrt_triangle); return USRT_OK;
        case USRT_BUF_TRIANGLE_AABB: *ptr = ctx->tri_aabb; *elem = sizeof(usrt_aabb); return USRT_OK;
        case USRT_BUF_BVH_DATA: *ptr = ctx->bvh; *elem = sizeof(usrt_aabb); return USRT_OK;
        case USRT_BUF_LEAF_NODES: *ptr = ctx->leaf; *elem = sizeof(usrt_leaf_node); return USRT_OK;
        case USRT_BUF_INTERNAL_NODES: *ptr = ctx->internal; *elem = sizeof(usrt_internal_node); return USRT_OK;
        default: return fail(ctx, USRT_ERR_ARG, "unknown buffer id %d", buffer);
    }
}

int usrt_download(usrt_context* ctx, int buffer, void* host_dst, uint64_t count) {
    NEED_CTX(ctx);
    void* p; size_t elem;
    if (int r = buffer_info(ctx, buffer, &p, &elem)) return r;
    if (!host_dst || count > ctx->capacity) return fail(ctx, USRT_ERR_ARG, "download: count=%llu capacity=%u", (unsigned long long)count, ctx->capacity);
    if (int r = bind_device(ctx)) return r;
    CU(ctx, cudaMemcpyAsync(host_dst, p, count * elem, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return USRT_OK;
}

int usrt_device_ptr(usrt_context* ctx, int buffer, void** dev_ptr) {
    NEED_CTX(ctx);
    if (!dev_ptr) return USRT_ERR_ARG;
    size_t elem;
    return buffer_info(ctx, buffer, dev_ptr, &elem);
}

int usrt_count_corrupted_nodes(usrt_context* ctx, uint32_t* leaf_corrupted, uint32_t* internal_corrupted) {
    NEED_CTX(ctx);
    if (int r = bind_device(ctx)) return r;
    CU(ctx, launch_count_corrupted(ctx->leaf, ctx->internal, ctx->n, ctx->small, ctx->stream));
    ctx->launches += 1;
    uint32_t h[2] = {0, 0};
    CU(ctx, cudaMemcpyAsync(h, ctx->small, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    if (leaf_corrupted) *leaf_corrupted = h[0];
    if (internal_corrupted) *internal_corrupted = h[1];
    return USRT_OK;
}

}  // extern "C"
