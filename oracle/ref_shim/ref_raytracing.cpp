// ref_raytracing.cpp -- TEST INFRASTRUCTURE ONLY. Runs the reference's Raytracing kernel (text generated from
// Raytracing.compute by build_ref.sh) for every pixel and captures the RaycastResult it holds just before the
// shading epilogue (:178), plus the float4 the epilogue writes to _outputTexture.
#include "hlsl_shim.hpp"

#include <thread>
#include <vector>

static float4 _ProjectionParams;             // UnityShaderVariables.cginc: .y = near plane of the camera

struct RefHit { float distance; uint triangleIndex; float u, v; };
static RefHit* g_hits = nullptr;
static int g_hits_width = 0;
// the one line build_ref.sh inserts ahead of Raytracing.compute:178
template <typename R> static inline void usrt_ref_emit(const uint3& id, const R& r) {
    RefHit h; h.distance = r.distance; h.triangleIndex = r.triangleIndex; h.u = r.uv.x; h.v = r.uv.y;
    g_hits[(size_t)id.y * g_hits_width + id.x] = h;
}
#define USRT_REF_HIT(id, result) usrt_ref_emit(id, result)

namespace ref_raytracing_compute {   // each .compute / .cs file keeps its own globals
#include "../_ref/gen_Raytracing.inc"
}
using namespace ref_raytracing_compute;

static_assert(sizeof(Triangle) == 128 && sizeof(AABB) == 32, "Constants.cginc layouts");

extern "C" {

// RaytracingMeshDrawer.cs:63-70,76-83: bind the six buffers + texture, set the uniforms, dispatch. Only on-screen
// thread ids are run (the reference's (W/32+1) x (H/32+1) groups also trace off-screen pixels, with no effect).
// texture may be NULL (a 1x1 white texel is bound). out_rgba: W*H float4 (what :184 stores, before the RGBA16F
// render target rounds it). camera_to_world: 16 floats row-major.
void usrt_ref_raytracing(const uint* sorted_indices, const AABB* tri_aabb, const InternalNode* internal, const LeafNode* leaf,
                         const AABB* bvh, const Triangle* triangles, const float* texture, int tex_w, int tex_h,
                         int width, int height, float near_plane, float camera_fov, const float* camera_to_world,
                         RefHit* out_hits, float4* out_rgba, int threads) {
    static const float white[4] = {1.0f, 1.0f, 1.0f, 1.0f};
    sortedTriangleIndices.data = sorted_indices;
    triangleAABB.data = tri_aabb;
    internalNodes.data = internal;
    leafNodes.data = leaf;
    bvhData.data = bvh;
    triangleData.data = triangles;
    _meshTexture.texels = texture ? texture : white;
    _meshTexture.width = texture ? tex_w : 1;
    _meshTexture.height = texture ? tex_h : 1;
    std::vector<float4> scratch;
    if (!out_rgba) { scratch.resize((size_t)width * height); out_rgba = scratch.data(); }
    _outputTexture.pixels = out_rgba;
    _outputTexture.width = width;
    screenWidth = width;
    screenHeight = height;
    cameraFov = camera_fov;
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) cameraToWorldMatrix.m[r][c] = camera_to_world[r * 4 + c];
    _ProjectionParams = float4(0.0f, near_plane, 0.0f, 0.0f);
    g_hits = out_hits;
    g_hits_width = width;
    if (threads < 1) threads = 1;
    std::vector<std::thread> pool;                 // the kernel only reads shared state; every pixel writes its own slots
    for (int w = 0; w < threads; ++w)
        pool.emplace_back([=]() {
            for (int y = w; y < height; y += threads)
                for (int x = 0; x < width; ++x) Raytracing(uint3((uint)x, (uint)y, 0));
        });
    for (auto& t : pool) t.join();
}

}  // extern "C"
