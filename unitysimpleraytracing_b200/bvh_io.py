"""an on-disk format for a built BVH -- the reference has none (the tree is rebuilt
from the mesh in every Awake(), RaytracingMeshDrawer.cs:30-54). The file is the seven scene buffers of
MeshBufferContainer.cs:87-94 verbatim, little-endian, in the reference's struct layouts, behind a small
header, so it can be handed to the P/Invoke host or reloaded into a context without rebuilding.

    header : magic 'USRTBVH1' | uint32 n | uint32 flags (0) | 8 x uint32 section byte sizes
    body   : keys[n] u32 | triangleIndex[n] u32 | triangleData[n] 128 B | triangleAABB[n] 32 B |
             bvhData[n-1] 32 B | leafNodes[n] 8 B | internalNodes[n-1] 24 B
"""
import struct

import numpy as np

from . import _lib
from .scene_types import AABB, InternalNode, LeafNode, Triangle

MAGIC = b"USRTBVH1"
_SECTIONS = (("keys", np.dtype("<u4"), 0), ("triangleIndex", np.dtype("<u4"), 0), ("triangleData", Triangle, 0),
             ("triangleAABB", AABB, 0), ("bvhData", AABB, 1), ("leafNodes", LeafNode, 0), ("internalNodes", InternalNode, 1))
_BUFS = (_lib.BUF_KEYS, _lib.BUF_TRIANGLE_INDEX, _lib.BUF_TRIANGLE_DATA, _lib.BUF_TRIANGLE_AABB, _lib.BUF_BVH_DATA,
         _lib.BUF_LEAF_NODES, _lib.BUF_INTERNAL_NODES)


def download_bvh(ctx):
    """dict of the seven buffers of a built context (host copies)."""
    n = ctx.triangles_length
    return {name: ctx.download(buf, n - short) for (name, _, short), buf in zip(_SECTIONS, _BUFS)}


def save_bvh(ctx, path):
    bufs = download_bvh(ctx)
    n = ctx.triangles_length
    with open(path, "wb") as f:
        sizes = [bufs[name].nbytes for name, _, _ in _SECTIONS] + [0]
        f.write(MAGIC + struct.pack("<II8I", n, 0, *sizes))
        for name, _, _ in _SECTIONS:
            f.write(np.ascontiguousarray(bufs[name]).tobytes())
    return bufs


def read_bvh(path):
    with open(path, "rb") as f:
        head = f.read(8 + 4 + 4 + 32)
        if head[:8] != MAGIC:
            raise ValueError("not a USRTBVH1 file")
        n, flags, *sizes = struct.unpack("<II8I", head[8:])
        out = {}
        for (name, dt, short), size in zip(_SECTIONS, sizes):
            count = n - short
            if size != count * dt.itemsize:
                raise ValueError("section %s: %d bytes, expected %d" % (name, size, count * dt.itemsize))
            out[name] = np.frombuffer(f.read(size), dtype=dt, count=count).copy()
    return n, out


def upload_bvh(ctx, n, bufs):
    """Install a finished BVH (usrt_upload_bvh): keys, triangleIndex, triangleData, triangleAABB, bvhData,
    leafNodes, internalNodes as numpy arrays in the reference layouts."""
    arrs = [np.ascontiguousarray(bufs[name], dt) for name, dt, _ in _SECTIONS]
    for a, (name, _, short) in zip(arrs, _SECTIONS):
        if len(a) < n - short:
            raise ValueError("%s has %d entries, need %d" % (name, len(a), n - short))
    ptr = lambda a: a.ctypes.data_as(__import__("ctypes").c_void_p)
    ctx._check(ctx._lib.usrt_upload_bvh(ctx._h, n, *[ptr(a) for a in arrs]))


def load_bvh(ctx, path):
    n, bufs = read_bvh(path)
    upload_bvh(ctx, n, bufs)
    return n
