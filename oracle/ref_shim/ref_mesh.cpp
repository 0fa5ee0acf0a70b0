// ref_mesh.cpp -- TEST INFRASTRUCTURE ONLY. Drives the reference's own Morton / AABB functions and DistributeKeys
// (text generated from MeshBufferContainer.cs by build_ref.sh) the way the constructor loop does.
#include "unity_shim.hpp"
namespace ref_mesh_container {   // each .compute / .cs file keeps its own globals
#include "../_ref/gen_MeshBufferContainer.inc"
}
using namespace ref_mesh_container;

extern "C" {

// MeshBufferContainer.cs:123-146: for every triangle, GetCentroidAndAABB -> NormalizeCentroid -> Morton3D;
// keys[i] = code, triangleIndex[i] = i, triangleAABB[i] = aabb. Vertices come from the packed 128-byte Triangle
// (a, b, c at float offsets 0, 4, 8), which is what :133-137 stores.
void usrt_ref_morton(const float* triangles, uint n, uint* keys, uint* values, AABB* aabbs) {
    for (uint i = 0; i < n; i++) {
        const float* t = triangles + (size_t)i * 32;
        Vector3 a(t[0], t[1], t[2]), b(t[4], t[5], t[6]), c(t[8], t[9], t[10]);
        Vector3 centroid;
        AABB aabb;
        GetCentroidAndAABB(a, b, c, centroid, aabb);
        centroid = NormalizeCentroid(centroid);
        uint mortonCode = Morton3D(centroid.x, centroid.y, centroid.z);
        keys[i] = mortonCode;
        values[i] = i;
        aabbs[i] = aabb;
    }
}

void usrt_ref_distribute_keys(uint* keys, uint trianglesLength) {
    _keysBuffer.LocalBuffer = keys;
    _trianglesLength = trianglesLength;
    DistributeKeys();
}

}  // extern "C"
