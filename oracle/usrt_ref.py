"""ctypes loader for oracle/_ref/libusrt_ref.so -- TEST INFRASTRUCTURE ONLY.

That library is the REFERENCE'S OWN TEXT (BVH.compute, Raytracing.compute, Sorting/*.compute, the static functions and
DistributeKeys of MeshBufferContainer.cs), compiled with g++ through the syntactic recipe in oracle/build_ref.sh; the
sort kernels run under a lock-step wave emulator (oracle/ref_shim/wave_emulator.hpp). It pins the
hand-written restatement oracle/usrt_oracle.cpp (tests/test_ref_pin.py) and generates the golden digests under
tests/golden/ (tests/golden/make_ref_golden.py). /root/reference exists only in the build container: on the GPU box the
prebuilt .so is used if it travelled with the snapshot, and nothing here is imported by the product.
"""
import ctypes
import os
import subprocess

import numpy as np

from .usrt_oracle import AABB, INTERNAL_NODE, LEAF_NODE, RAYCAST_RESULT, TRIANGLE, null_internal, null_leaf

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_ref", "libusrt_ref.so")
REFERENCE = os.environ.get("REF", "/root/reference")


def reference_present():
    return os.path.exists(os.path.join(REFERENCE, "Assets", "_Shaders", "BVH", "BVH.compute"))


def build(force=False):
    """(Re)build from the reference checkout when it is present; otherwise keep whatever .so travelled here."""
    if reference_present():
        srcs = [os.path.join(_HERE, "build_ref.sh")] + [os.path.join(_HERE, "ref_shim", f) for f in os.listdir(os.path.join(_HERE, "ref_shim"))]
        stale = not os.path.exists(_LIB_PATH) or any(os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs)
        if force or stale:
            subprocess.check_call([os.path.join(_HERE, "build_ref.sh")], stdout=subprocess.DEVNULL)
    return _LIB_PATH if os.path.exists(_LIB_PATH) else None


def available():
    return build() is not None


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError("oracle/_ref/libusrt_ref.so is not built and no reference checkout is present")
        _lib = ctypes.CDLL(_LIB_PATH)
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def morton(tris):
    """MeshBufferContainer.cs:123-146 -> (keys, triangleIndex, triangleAABB)."""
    tris = np.ascontiguousarray(tris, TRIANGLE)
    n = len(tris)
    keys = np.empty(n, np.uint32); values = np.empty(n, np.uint32); aabbs = np.zeros(n, AABB)
    lib().usrt_ref_morton(_p(tris), ctypes.c_uint32(n), _p(keys), _p(values), _p(aabbs))
    return keys, values, aabbs


def distribute_keys(keys):
    k = np.array(keys, np.uint32)
    lib().usrt_ref_distribute_keys(_p(k), ctypes.c_uint32(len(k)))
    return k


def _dispatch_threads(n):
    return -(-n // 1024) * 1024          # whole THREADS_PER_BLOCK groups, like Dispatch(); ids past the guard do nothing


def construct_tree(keys, n):
    internal = null_internal(n); leaf = null_leaf(n)
    keys = np.ascontiguousarray(keys, np.uint32)
    lib().usrt_ref_construct_tree(_p(keys), ctypes.c_uint32(n), _p(internal), _p(leaf), ctypes.c_uint32(_dispatch_threads(n)))
    return internal, leaf


def construct_bvh(n, sorted_indices, tri_aabb, internal, leaf):
    bvh = np.zeros(n, AABB)
    atomics = np.zeros(n, np.uint32)                                  # BVHConstructor.cs:41
    lib().usrt_ref_construct_bvh(ctypes.c_uint32(n), _p(np.ascontiguousarray(sorted_indices, np.uint32)),
                                 _p(np.ascontiguousarray(tri_aabb, AABB)), _p(internal), _p(leaf), _p(atomics), _p(bvh),
                                 ctypes.c_uint32(_dispatch_threads(n)))
    return bvh


def raytracing(sorted_indices, tri_aabb, internal, leaf, bvh, tris, width, height, near, tan_half_fov, cam_to_world,
               texture=None, threads=None):
    """The Raytracing kernel over a width x height frame -> (hit records as the kernel holds them at :176,
    float32 RGBA the epilogue :178-184 writes)."""
    hits = np.zeros(width * height, RAYCAST_RESULT)
    rgba = np.zeros((width * height, 4), np.float32)
    m = np.ascontiguousarray(cam_to_world, np.float32).reshape(16)
    tex = None if texture is None else np.ascontiguousarray(texture, np.float32)
    th, tw = (0, 0) if tex is None else tex.shape[:2]
    lib().usrt_ref_raytracing(_p(np.ascontiguousarray(sorted_indices, np.uint32)), _p(np.ascontiguousarray(tri_aabb, AABB)),
                              _p(np.ascontiguousarray(internal, INTERNAL_NODE)), _p(np.ascontiguousarray(leaf, LEAF_NODE)),
                              _p(np.ascontiguousarray(bvh, AABB)), _p(np.ascontiguousarray(tris, TRIANGLE)), _p(tex),
                              ctypes.c_int(tw), ctypes.c_int(th), ctypes.c_int(width), ctypes.c_int(height),
                              ctypes.c_float(near), ctypes.c_float(tan_half_fov), _p(m), _p(hits), _p(rgba),
                              ctypes.c_int(threads or os.cpu_count() or 1))
    return hits, rgba


class Scene:
    """RaytracingMeshDrawer.Awake() (:30-54) with the reference's own code for every stage: MeshBufferContainer's Morton /
    AABB functions, the five Sorting kernels under the wave emulator, DistributeKeys, TreeConstructor, BVHConstructor.
    The sort kernels are hard-wired to 512 blocks of 1024 elements (Scan.compute:50,64); meshes beyond that capacity
    (BASELINE configs[1] and up) take the sort's contract instead -- a stable ascending sort by the 32-bit key
    (ComputeBufferSorter.cs:150-177) -- applied with numpy, and say so in `sorted_by`."""

    def __init__(self, tris):
        self.triangleData = np.ascontiguousarray(tris, TRIANGLE)
        self.n = n = len(tris)
        self.mortonCodes, idx, self.triangleAABB = morton(self.triangleData)
        if n <= REF_BLOCKS * 1024:
            self.sortedMortonRaw, self.sortedTriangleIndices = sort(self.mortonCodes, idx)
            self.sorted_by = "reference kernels (Sorting/*.compute under the wave emulator)"
        else:
            order = np.argsort(self.mortonCodes, kind="stable")
            self.sortedMortonRaw = self.mortonCodes[order]
            self.sortedTriangleIndices = idx[order]
            self.sorted_by = "numpy stable sort (mesh exceeds the reference's fixed 524,288-element sort capacity)"
        self.sortedMortonCodes = distribute_keys(self.sortedMortonRaw)
        self.internalNodes, self.leafNodes = construct_tree(self.sortedMortonCodes, n)
        self.bvhData = construct_bvh(n, self.sortedTriangleIndices, self.triangleAABB, self.internalNodes, self.leafNodes)

    def trace_primary(self, width, height, near, tan_half_fov, cam_to_world, texture=None, threads=None):
        return raytracing(self.sortedTriangleIndices, self.triangleAABB, self.internalNodes, self.leafNodes, self.bvhData,
                          self.triangleData, width, height, near, tan_half_fov, cam_to_world, texture, threads)


# ---- the reference's five sort kernels under the lock-step wave emulator (oracle/ref_shim/wave_emulator.hpp) ----------
REF_BLOCKS = 512          # Constants.cginc:3 BLOCK_SIZE: the count tables are laid out for 512 blocks whatever is dispatched


def sort_pass(keys, values, bit_offset):
    """One pass of ComputeBufferSorter.Sort() (:104-116) through LocalRadixSort, PreScan, BlockSum, GlobalScan and
    GlobalRadixSort as written. len(keys) must be a multiple of 1024 and at most 512 * 1024. Returns every
    intermediate in the reference's layouts: offsets[block * 256 + digit], sizes[digit * 512 + block]."""
    n = len(keys)
    assert n % 1024 == 0 and 0 < n <= REF_BLOCKS * 1024
    k = np.array(keys, np.uint32); v = np.array(values, np.uint32)
    out = dict(sortedBlocksKeys=np.empty(n, np.uint32), sortedBlocksValues=np.empty(n, np.uint32),
               offsets=np.zeros(REF_BLOCKS * 256, np.uint32), sizesBefore=np.zeros(REF_BLOCKS * 256, np.uint32),
               sizesAfter=np.zeros(REF_BLOCKS * 256, np.uint32))
    lib().usrt_ref_sort_pass(_p(k), _p(v), ctypes.c_int(n // 1024), ctypes.c_int(bit_offset), _p(out["sortedBlocksKeys"]),
                             _p(out["sortedBlocksValues"]), _p(out["offsets"]), _p(out["sizesBefore"]), _p(out["sizesAfter"]))
    out["keys"] = k; out["values"] = v
    return out


def sort(keys, values):
    """ComputeBufferSorter.Sort() (:100-126): four passes of the kernels above. Padded like the reference's buffers
    (MeshBufferContainer.cs:108-109: unused slots hold 0xFFFFFFFF, which sink to the end) and cut back to len(keys)."""
    n = len(keys)
    padded = -(-n // 1024) * 1024
    assert padded <= REF_BLOCKS * 1024
    k = np.full(padded, 0xFFFFFFFF, np.uint32); v = np.full(padded, 0xFFFFFFFF, np.uint32)
    k[:n] = keys; v[:n] = values
    lib().usrt_ref_sort(_p(k), _p(v), ctypes.c_int(padded // 1024))
    return k[:n].copy(), v[:n].copy()
