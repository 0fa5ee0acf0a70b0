// ref_bvh.cpp -- TEST INFRASTRUCTURE ONLY. Runs the reference's TreeConstructor and BVHConstructor kernels (text
// generated from BVH.compute by build_ref.sh, Constants.cginc included from the reference checkout) one thread id
// after the other.
#include "hlsl_shim.hpp"

// BVH.compute:99,177 sit at the top of each kernel before any work: no cross-thread effect to emulate
inline void AllMemoryBarrierWithGroupSync() {}

namespace ref_bvh_compute {   // each .compute / .cs file keeps its own globals
#include "../_ref/gen_BVH.inc"
}
using namespace ref_bvh_compute;

static_assert(sizeof(AABB) == 32 && sizeof(InternalNode) == 24 && sizeof(LeafNode) == 8, "Constants.cginc layouts");

extern "C" {

// BVHConstructor.cs:61-64: Dispatch(TreeConstructor, BLOCK_SIZE groups of THREADS_PER_BLOCK). `threads` = how many
// thread ids to run (the reference always runs 524,288; ids >= trianglesCount - 1 return at BVH.compute:101).
// internalNodes / leafNodes must arrive NullLeaf-filled (MeshBufferContainer.cs:114-115).
void usrt_ref_construct_tree(const uint* codes, uint n, InternalNode* internal, LeafNode* leaf, uint threads) {
    trianglesCount = (int)n;
    sortedMortonCodes.data = codes;
    internalNodes.data = internal;
    leafNodes.data = leaf;
    for (uint t = 0; t < threads; ++t) TreeConstructor(uint3(t, 0, 0));
}

// BVHConstructor.cs:66-69. atomics must arrive zeroed (BVHConstructor.cs:41). Serial execution is one legal
// interleaving of the kernel: a thread that finds the counter at 0 leaves, the later one merges boxes that are complete.
void usrt_ref_construct_bvh(uint n, const uint* sorted_indices, const AABB* tri_aabb, InternalNode* internal, LeafNode* leaf,
                            uint* atomics, AABB* bvh, uint threads) {
    trianglesCount = (int)n;
    sortedTriangleIndices.data = sorted_indices;
    triangleAABB.data = tri_aabb;
    internalNodes.data = internal;
    leafNodes.data = leaf;
    atomicsData.data = atomics;
    BVHData.data = bvh;
    for (uint t = 0; t < threads; ++t) BVHConstructor(uint3(t, 0, 0));
}

}  // extern "C"
