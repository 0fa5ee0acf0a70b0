"""ctypes loader for the CPU oracle (oracle/usrt_oracle.cpp) -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this module. Pinned against the reference's own text compiled by oracle/build_ref.sh (oracle/_ref,
tests/test_ref_pin.py) for Morton/AABB, the sort (all five kernels under a wave emulator), DistributeKeys,
tree, refit, traversal and shading; further cross-checked by the reference's runtime self-checks, an
independent numpy restatement (np_oracle.py) and brute force. See the header of usrt_oracle.cpp.

The struct dtypes are declared here independently of the product package on purpose, so that a
layout bug in the product's header shows up as a byte mismatch in the parity tests.
Reference layouts: Assets/_Shaders/Constants.cginc:9-54, Assets/_Scripts/SceneDataTypes.cs:4-90,
Assets/_Shaders/Raytracing/Raytracing.compute:30-35.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libusrt_oracle.so")

AABB = np.dtype([("min", "<f4", 3), ("_dummy0", "<f4"), ("max", "<f4", 3), ("_dummy1", "<f4")])
INTERNAL_NODE = np.dtype([("leftNode", "<u4"), ("leftNodeType", "<u4"), ("rightNode", "<u4"),
                          ("rightNodeType", "<u4"), ("parent", "<u4"), ("index", "<u4")])
LEAF_NODE = np.dtype([("parent", "<u4"), ("index", "<u4")])
TRIANGLE = np.dtype([("a", "<f4", 3), ("_d0", "<f4"), ("b", "<f4", 3), ("_d1", "<f4"), ("c", "<f4", 3), ("_d2", "<f4"),
                     ("a_uv", "<f4", 2), ("b_uv", "<f4", 2), ("c_uv", "<f4", 2), ("_d3", "<f4", 2),
                     ("a_normal", "<f4", 3), ("_d4", "<f4"), ("b_normal", "<f4", 3), ("_d5", "<f4"),
                     ("c_normal", "<f4", 3), ("_d6", "<f4")])
RAYCAST_RESULT = np.dtype([("distance", "<f4"), ("triangleIndex", "<u4"), ("uv", "<f4", 2)])
assert AABB.itemsize == 32 and INTERNAL_NODE.itemsize == 24 and LEAF_NODE.itemsize == 8
assert TRIANGLE.itemsize == 128 and RAYCAST_RESULT.itemsize == 16

WHOLE_MIN, WHOLE_MAX = -125.0, 125.0   # MeshBufferContainer.cs:9-15


def build(force=False):
    """Compile the oracle with the committed Makefile (g++ only; seconds)."""
    src = os.path.join(_HERE, "usrt_oracle.cpp")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "libusrt_oracle.so"],
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.usrt_oracle_max_float.restype = ctypes.c_float
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def max_float():
    return np.float32(lib().usrt_oracle_max_float())


def scene_box(tris):
    """Per-axis min / max of all vertices (the MeshBufferContainer.cs:7 TODO; flat axes get max = min + 1)."""
    tris = np.ascontiguousarray(tris, dtype=TRIANGLE)
    lo, hi = np.zeros(3, np.float32), np.zeros(3, np.float32)
    lib().usrt_oracle_scene_box(_p(tris), ctypes.c_uint32(len(tris)), _p(lo), _p(hi))
    return lo, hi


def morton(tris, whole_min=WHOLE_MIN, whole_max=WHOLE_MAX):
    """whole_min / whole_max: scalars (the reference's cube) or 3-vectors (a per-axis box)."""
    n = len(tris)
    tris = np.ascontiguousarray(tris, dtype=TRIANGLE)
    keys = np.empty(n, np.uint32); values = np.empty(n, np.uint32); aabbs = np.zeros(n, AABB)
    if np.ndim(whole_min) or np.ndim(whole_max):
        lo = np.ascontiguousarray(np.broadcast_to(np.asarray(whole_min, np.float32), 3))
        hi = np.ascontiguousarray(np.broadcast_to(np.asarray(whole_max, np.float32), 3))
        lib().usrt_oracle_morton_box(_p(tris), ctypes.c_uint32(n), _p(lo), _p(hi), _p(keys), _p(values), _p(aabbs))
        return keys, values, aabbs
    lib().usrt_oracle_morton(_p(tris), ctypes.c_uint32(n), ctypes.c_float(whole_min), ctypes.c_float(whole_max),
                             _p(keys), _p(values), _p(aabbs))
    return keys, values, aabbs


def sort_pass(keys, values, bit_offset):
    """One reference pass; returns dict of every intermediate (in-place on copies)."""
    n = len(keys)
    assert n % 1024 == 0
    nb = n // 1024
    k = np.array(keys, np.uint32); v = np.array(values, np.uint32)
    out = dict(sortedBlocksKeys=np.empty(n, np.uint32), sortedBlocksValues=np.empty(n, np.uint32),
               offsets=np.empty(nb * 256, np.uint32), sizesBefore=np.empty(nb * 256, np.uint32),
               sizesAfter=np.empty(nb * 256, np.uint32))
    lib().usrt_oracle_sort_pass(_p(k), _p(v), ctypes.c_uint32(n), ctypes.c_int(bit_offset),
                                _p(out["sortedBlocksKeys"]), _p(out["sortedBlocksValues"]), _p(out["offsets"]),
                                _p(out["sizesBefore"]), _p(out["sizesAfter"]))
    out["keys"] = k; out["values"] = v
    return out


def sort(keys, values):
    k = np.array(keys, np.uint32); v = np.array(values, np.uint32)
    lib().usrt_oracle_sort(_p(k), _p(v), ctypes.c_uint64(len(k)))
    return k, v


def stable_sort(keys, values):
    k = np.array(keys, np.uint32); v = np.array(values, np.uint32)
    lib().usrt_oracle_stable_sort(_p(k), _p(v), ctypes.c_uint64(len(k)))
    return k, v


def distribute_keys(keys, n=None):
    k = np.array(keys, np.uint32)
    lib().usrt_oracle_distribute_keys(_p(k), ctypes.c_uint32(len(k) if n is None else n))
    return k


def null_internal(n):
    return np.full(n * 6, 0xFFFFFFFF, np.uint32).view(INTERNAL_NODE)


def null_leaf(n):
    return np.full(n * 2, 0xFFFFFFFF, np.uint32).view(LEAF_NODE)


def construct_tree(keys, n, capacity=None):
    capacity = n if capacity is None else capacity
    internal = null_internal(capacity); leaf = null_leaf(capacity)
    keys = np.ascontiguousarray(keys, np.uint32)
    lib().usrt_oracle_construct_tree(_p(keys), ctypes.c_uint32(n), _p(internal), _p(leaf))
    return internal, leaf


def construct_bvh(n, sorted_indices, tri_aabb, internal, leaf, capacity=None):
    capacity = n if capacity is None else capacity
    bvh = np.zeros(capacity, AABB)
    lib().usrt_oracle_construct_bvh(ctypes.c_uint32(n), _p(np.ascontiguousarray(sorted_indices, np.uint32)),
                                    _p(np.ascontiguousarray(tri_aabb, AABB)), _p(internal), _p(leaf), _p(bvh))
    return bvh


# ---- SURVEY 8(f)-4 key variants (defined by the oracle, see usrt_oracle.cpp) ----------------------------------------
def morton64(tris, whole_min=WHOLE_MIN, whole_max=WHOLE_MAX):
    n = len(tris)
    tris = np.ascontiguousarray(tris, dtype=TRIANGLE)
    keys = np.empty(n, np.uint64); values = np.empty(n, np.uint32); aabbs = np.zeros(n, AABB)
    lib().usrt_oracle_morton64(_p(tris), ctypes.c_uint32(n), ctypes.c_float(whole_min), ctypes.c_float(whole_max),
                               _p(keys), _p(values), _p(aabbs))
    return keys, values, aabbs


def sort64(keys, values):
    k = np.array(keys, np.uint64); v = np.array(values, np.uint32)
    lib().usrt_oracle_sort64(_p(k), _p(v), ctypes.c_uint64(len(k)))
    return k, v


def stable_sort64(keys, values):
    k = np.array(keys, np.uint64); v = np.array(values, np.uint32)
    lib().usrt_oracle_stable_sort64(_p(k), _p(v), ctypes.c_uint64(len(k)))
    return k, v


def distribute_keys64(keys):
    k = np.array(keys, np.uint64)
    lib().usrt_oracle_distribute_keys64(_p(k), ctypes.c_uint32(len(k)))
    return k


def construct_tree64(keys64, n):
    internal = null_internal(n); leaf = null_leaf(n)
    keys64 = np.ascontiguousarray(keys64, np.uint64)
    lib().usrt_oracle_construct_tree64(_p(keys64), ctypes.c_uint32(n), _p(internal), _p(leaf))
    return internal, leaf


class VariantScene:
    """The build with one of the key variants; traversal is the Scene's (it never looks at keys).
    mode 1: 32-bit Morton codes, NO DistributeKeys, tree on (code << 32 | sorted position).
    mode 2: 63-bit Morton codes, 8-pass sort, 64-bit DistributeKeys, tree on the distributed 64-bit keys."""

    def __init__(self, tris, mode):
        self.triangleData = np.ascontiguousarray(tris, TRIANGLE)
        self.n = n = len(tris)
        if mode == 1:
            codes, idx, self.triangleAABB = morton(self.triangleData)
            self.sortedMortonCodes, self.sortedTriangleIndices = sort(codes, idx)
            tree_keys = (self.sortedMortonCodes.astype(np.uint64) << np.uint64(32)) | np.arange(n, dtype=np.uint64)
        elif mode == 2:
            codes, idx, self.triangleAABB = morton64(self.triangleData)
            self.sortedMortonRaw, self.sortedTriangleIndices = sort64(codes, idx)
            self.sortedMortonCodes = tree_keys = distribute_keys64(self.sortedMortonRaw)
        else:
            raise ValueError(mode)
        self.internalNodes, self.leafNodes = construct_tree64(tree_keys, n)
        self.bvhData = construct_bvh(n, self.sortedTriangleIndices, self.triangleAABB, self.internalNodes, self.leafNodes)

    _scene_args = None


class Scene:
    """Everything the build produces, with the reference's buffer names."""

    def __init__(self, tris, whole_min=WHOLE_MIN, whole_max=WHOLE_MAX, timings=None):
        import time
        t0 = time.perf_counter()
        self.triangleData = np.ascontiguousarray(tris, TRIANGLE)
        self.n = len(tris)
        self.mortonCodes, idx, self.triangleAABB = morton(self.triangleData, whole_min, whole_max)
        t1 = time.perf_counter()
        self.sortedMortonRaw, self.sortedTriangleIndices = sort(self.mortonCodes, idx)
        t2 = time.perf_counter()
        self.sortedMortonCodes = distribute_keys(self.sortedMortonRaw)
        t3 = time.perf_counter()
        self.internalNodes, self.leafNodes = construct_tree(self.sortedMortonCodes, self.n)
        t4 = time.perf_counter()
        self.bvhData = construct_bvh(self.n, self.sortedTriangleIndices, self.triangleAABB, self.internalNodes,
                                     self.leafNodes)
        t5 = time.perf_counter()
        if timings is not None:
            timings.update(morton=t1 - t0, sort=t2 - t1, distribute=t3 - t2, tree=t4 - t3, refit=t5 - t4,
                           build=t5 - t0)

    def _scene_args(self):
        return (_p(self.sortedTriangleIndices), _p(self.triangleAABB), _p(self.internalNodes), _p(self.leafNodes),
                _p(self.bvhData), _p(self.triangleData))

    def trace_primary(self, width, height, near, tan_half_fov, cam_to_world, y0=0, y1=None, threads=1, counters=False):
        y1 = height if y1 is None else y1
        out = np.zeros(width * height, RAYCAST_RESULT)
        m = np.ascontiguousarray(cam_to_world, np.float32).reshape(16)
        cnt = np.zeros(4, np.uint64) if counters else None
        lib().usrt_oracle_trace_primary(*self._scene_args(), ctypes.c_int(width), ctypes.c_int(height),
                                        ctypes.c_float(near), ctypes.c_float(tan_half_fov), _p(m),
                                        ctypes.c_uint32(y0), ctypes.c_uint32(y1), _p(out), ctypes.c_int(threads),
                                        _p(cnt))
        return (out, cnt) if counters else out

    def trace_rays(self, rays, threads=1, counters=False):
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 8)
        out = np.zeros(len(rays), RAYCAST_RESULT)
        cnt = np.zeros(4, np.uint64) if counters else None
        lib().usrt_oracle_trace_rays(*self._scene_args(), _p(rays), ctypes.c_uint64(len(rays)), _p(out),
                                     ctypes.c_int(threads), _p(cnt))
        return (out, cnt) if counters else out

    def visit_order(self):
        order = np.empty(self.n, np.uint32)
        lib().usrt_oracle_visit_order(_p(self.sortedTriangleIndices), _p(self.internalNodes), _p(self.leafNodes),
                                      ctypes.c_uint32(self.n), _p(order))
        return order

    def brute_force(self, rays, order=None, threads=1):
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 8)
        out = np.zeros(len(rays), RAYCAST_RESULT)
        o = None if order is None else np.ascontiguousarray(order, np.uint32)
        lib().usrt_oracle_brute_force(_p(self.triangleAABB), _p(self.triangleData), _p(o), ctypes.c_uint32(self.n),
                                      _p(rays), ctypes.c_uint64(len(rays)), _p(out), ctypes.c_int(threads))
        return out


for _m in ("_scene_args", "trace_primary", "trace_rays", "visit_order", "brute_force"):
    setattr(VariantScene, _m, getattr(Scene, _m))


def primary_rays(width, height, near, tan_half_fov, cam_to_world):
    m = np.ascontiguousarray(cam_to_world, np.float32).reshape(16)
    rays = np.zeros((width * height, 8), np.float32)
    lib().usrt_oracle_primary_rays(ctypes.c_int(width), ctypes.c_int(height), ctypes.c_float(near),
                                   ctypes.c_float(tan_half_fov), _p(m), _p(rays))
    return rays


def shade(hits, triangle_data, texture_rgba):
    """Raytracing.compute:178-184 on host: returns (count, 4) float16 (RGBA16F)."""
    hits = np.ascontiguousarray(hits, RAYCAST_RESULT)
    tex = np.ascontiguousarray(texture_rgba, np.float32)
    th, tw = tex.shape[0], tex.shape[1]
    out = np.zeros((len(hits), 4), np.uint16)
    lib().usrt_oracle_shade(_p(hits), ctypes.c_uint64(len(hits)), _p(np.ascontiguousarray(triangle_data, TRIANGLE)),
                            _p(tex), ctypes.c_int(tw), ctypes.c_int(th), _p(out))
    return out.view(np.float16)


def diffuse_rays(primary_hits, triangle_data, width, height, near, tan_half_fov, cam_to_world, seed, first_sample,
                 num_samples):
    """BASELINE configs[4] bounce rays (defined by this oracle; the reference has primary rays only):
    (num_samples * width * height, 8) float32, sample-major."""
    hits = np.ascontiguousarray(primary_hits, RAYCAST_RESULT)
    assert len(hits) == width * height
    m = np.ascontiguousarray(cam_to_world, np.float32).reshape(16)
    out = np.zeros((num_samples * width * height, 8), np.float32)
    lib().usrt_oracle_diffuse_rays(_p(hits), _p(np.ascontiguousarray(triangle_data, TRIANGLE)), ctypes.c_int(width),
                                   ctypes.c_int(height), ctypes.c_float(near), ctypes.c_float(tan_half_fov), _p(m),
                                   ctypes.c_uint64(seed), ctypes.c_uint32(first_sample), ctypes.c_uint32(num_samples), _p(out))
    return out


def float_to_half_bits(f):
    lib().usrt_oracle_float_to_half.restype = ctypes.c_uint16
    return int(lib().usrt_oracle_float_to_half(ctypes.c_float(f)))
