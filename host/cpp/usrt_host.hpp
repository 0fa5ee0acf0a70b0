// usrt_host.hpp -- C++ mirror of the reference's C# dispatch layer over the C ABI (include/usrt.h).
//
// The reference host is Unity C# (Assets/_Scripts/*.cs); no .NET toolchain exists in this image, so the
// compiled-language host that actually builds and runs here is this header (the P/Invoke source lives in
// host/csharp/UsrtNative.cs). Same class names, constructor arguments, call order and error behaviour as
//   MeshBufferContainer.cs, ComputeBufferSorter.cs, BVHConstructor.cs, DataBuffer.cs and the
//   Awake()/Update() sequence of RaytracingMeshDrawer.cs:30-54,76-84.
// Errors that the reference reports with Debug.LogError surface as usrt::Error exceptions carrying
// usrt_last_error().
#pragma once

#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/usrt.h"

namespace usrt {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& what) : std::runtime_error(what), code(c) {}
};

inline void check(usrt_context* ctx, int rc) {
    if (rc != USRT_OK) throw Error(rc, std::string("usrt error ") + std::to_string(rc) + ": " + usrt_last_error(ctx));
}

// Stands in for UnityEngine.ComputeBuffer: names one scene buffer of a context (DataBuffer.cs:7 DeviceBuffer).
struct DeviceBuffer {
    usrt_context* ctx;
    usrt_buffer which;
    template <class T> void GetData(std::vector<T>& dst) const {        // DataBuffer.cs:50-54
        check(ctx, usrt_download(ctx, which, dst.data(), dst.size()));
    }
};

class MeshBufferContainer {                                            // MeshBufferContainer.cs
  public:
    explicit MeshBufferContainer(const std::vector<usrt_triangle>& mesh, uint32_t capacity = 0, int device = 0) {
        static_assert(sizeof(usrt_triangle) == 128 && sizeof(usrt_aabb) == 32, "struct layout (:98-106)");
        const uint32_t n = (uint32_t)mesh.size();
        const uint32_t cap = capacity > n ? capacity : (n > 2 ? n : 2);
        const int rc = usrt_create(device, cap, &ctx_);
        if (rc != USRT_OK) throw Error(rc, "usrt_create failed (a CUDA device is required; there is no CPU path)");
        check(ctx_, usrt_upload_triangles(ctx_, mesh.data(), n));       // :148-151 Sync()
        check(ctx_, usrt_morton(ctx_));                                 // :123-146, on the GPU
    }
    ~MeshBufferContainer() { Dispose(); }
    MeshBufferContainer(const MeshBufferContainer&) = delete;
    MeshBufferContainer& operator=(const MeshBufferContainer&) = delete;

    DeviceBuffer Keys() const { return {ctx_, USRT_BUF_KEYS}; }                       // :17
    DeviceBuffer TriangleIndex() const { return {ctx_, USRT_BUF_TRIANGLE_INDEX}; }    // :19
    DeviceBuffer TriangleData() const { return {ctx_, USRT_BUF_TRIANGLE_DATA}; }      // :20
    DeviceBuffer TriangleAABB() const { return {ctx_, USRT_BUF_TRIANGLE_AABB}; }      // :21
    DeviceBuffer BvhData() const { return {ctx_, USRT_BUF_BVH_DATA}; }                // :22
    DeviceBuffer BvhLeafNode() const { return {ctx_, USRT_BUF_LEAF_NODES}; }          // :23
    DeviceBuffer BvhInternalNode() const { return {ctx_, USRT_BUF_INTERNAL_NODES}; }  // :24
    uint32_t TrianglesLength() const { return usrt_triangles_length(ctx_); }          // :30
    void DistributeKeys() { check(ctx_, usrt_distribute_keys(ctx_)); }                // :154-169
    void GetAllGpuData() {                                                             // :171-196
        uint32_t leaf = 0, inner = 0;
        check(ctx_, usrt_count_corrupted_nodes(ctx_, &leaf, &inner));
        if (leaf || inner) throw Error(USRT_ERR_STATE, "LEAF/INTERNAL CORRUPTED " + std::to_string(leaf) + "/" + std::to_string(inner));
    }
    void Dispose() { if (ctx_) { usrt_destroy(ctx_); ctx_ = nullptr; } }              // :207-216
    usrt_context* context() const { return ctx_; }

  private:
    usrt_context* ctx_ = nullptr;
};

class ComputeBufferSorter {                                            // ComputeBufferSorter.cs (TKey = TValue = uint)
  public:
    ComputeBufferSorter(uint32_t dataLength, DeviceBuffer keys, DeviceBuffer values) : ctx_(keys.ctx) {
        if (keys.ctx != values.ctx || keys.which != USRT_BUF_KEYS || values.which != USRT_BUF_TRIANGLE_INDEX)
            throw Error(USRT_ERR_ARG, "Sort() is bound to the container's Keys / TriangleIndex buffers");
        if (dataLength != usrt_triangles_length(ctx_)) throw Error(USRT_ERR_ARG, "dataLength != TrianglesLength");
    }
    void Sort() { check(ctx_, usrt_sort(ctx_)); }                      // :100-126
    // caller-owned arrays (the generic constructor of :44)
    static void Sort(usrt_context* ctx, std::vector<uint32_t>& keys, std::vector<uint32_t>& values) {
        check(ctx, usrt_sort_pairs_host(ctx, keys.data(), values.data(), keys.size()));
    }
    // ComputeBufferSorter<ulong, uint>: GetRadix is generic over uint / ulong keys (:179-191)
    static void Sort(usrt_context* ctx, std::vector<uint64_t>& keys, std::vector<uint32_t>& values) {
        check(ctx, usrt_sort_pairs64_host(ctx, keys.data(), values.data(), keys.size()));
    }

  private:
    usrt_context* ctx_;
};

class BVHConstructor {                                                 // BVHConstructor.cs:24-69
  public:
    BVHConstructor(uint32_t trianglesCount, DeviceBuffer sortedMortonCodes, DeviceBuffer, DeviceBuffer, DeviceBuffer,
                   DeviceBuffer, DeviceBuffer)
        : ctx_(sortedMortonCodes.ctx) {
        if (trianglesCount != usrt_triangles_length(ctx_)) throw Error(USRT_ERR_ARG, "trianglesCount != TrianglesLength");
    }
    void ConstructTree() { check(ctx_, usrt_construct_tree(ctx_)); }   // :61-64
    void ConstructBVH() { check(ctx_, usrt_construct_bvh(ctx_)); }     // :66-69

  private:
    usrt_context* ctx_;
};

class RaytracingMeshDrawer {                                           // RaytracingMeshDrawer.cs:30-54,76-84
  public:
    void Awake(const std::vector<usrt_triangle>& mesh) {
        container_.reset(new MeshBufferContainer(mesh));
        ComputeBufferSorter sorter(container_->TrianglesLength(), container_->Keys(), container_->TriangleIndex());
        sorter.Sort();
        container_->DistributeKeys();
        BVHConstructor bvh(container_->TrianglesLength(), container_->Keys(), container_->TriangleIndex(),
                           container_->TriangleAABB(), container_->BvhInternalNode(), container_->BvhLeafNode(),
                           container_->BvhData());
        bvh.ConstructTree();
        bvh.ConstructBVH();
        container_->GetAllGpuData();
    }
    void Rebuild() { check(container_->context(), usrt_rebuild(container_->context())); }
    // cameraFov = tan(fieldOfView * Deg2Rad / 2) (:80); near = camera.nearClipPlane; matrix row-major
    std::vector<usrt_raycast_result> Update(int screenWidth, int screenHeight, float nearPlane, float cameraFov,
                                            const float cameraToWorld[16]) {
        std::vector<usrt_raycast_result> hits((size_t)screenWidth * screenHeight);
        check(container_->context(), usrt_trace_primary(container_->context(), screenWidth, screenHeight, nearPlane,
                                                        cameraFov, cameraToWorld, 0, screenHeight, hits.data()));
        return hits;
    }
    MeshBufferContainer& container() { return *container_; }

  private:
    std::unique_ptr<MeshBufferContainer> container_;
};

}  // namespace usrt
