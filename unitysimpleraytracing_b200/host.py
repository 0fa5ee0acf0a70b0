"""Host side above the C ABI: a thin `Context` over include/usrt.h plus mirrors of the reference's
C# dispatch classes (same names, argument meaning and call order), so the parity tests read like
the reference's own driver, Assets/_Scripts/RaytracingMeshDrawer.cs:30-54,76-84:

    container = MeshBufferContainer(mesh)                       # RaytracingMeshDrawer.cs:34
    sorter = ComputeBufferSorter(container.TrianglesLength, container.Keys, container.TriangleIndex)
    sorter.Sort()                                               # :36-37
    container.DistributeKeys()                                  # :39
    bvh = BVHConstructor(container.TrianglesLength, container.Keys, container.TriangleIndex,
                         container.TriangleAABB, container.BvhInternalNode, container.BvhLeafNode,
                         container.BvhData)                     # :41-48
    bvh.ConstructTree(); bvh.ConstructBVH()                     # :50-51
    container.GetAllGpuData()                                   # :53

Everything computes on the GPU through libusrt_b200.so; nothing here has a CPU fallback.
"""
import ctypes

import numpy as np

from . import _lib
from .scene_types import AABB, InternalNode, LeafNode, RaycastResult, Triangle

_BUF_DTYPES = {
    _lib.BUF_KEYS: np.dtype("<u4"), _lib.BUF_TRIANGLE_INDEX: np.dtype("<u4"), _lib.BUF_TRIANGLE_DATA: Triangle,
    _lib.BUF_TRIANGLE_AABB: AABB, _lib.BUF_BVH_DATA: AABB, _lib.BUF_LEAF_NODES: LeafNode,
    _lib.BUF_INTERNAL_NODES: InternalNode, _lib.BUF_KEYS64: np.dtype("<u8"),
}


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


class Context:
    """Owns one usrt_context (one GPU, one stream)."""

    def __init__(self, capacity, device=0):
        self._lib = _lib.load()
        self._h = ctypes.c_void_p()
        rc = self._lib.usrt_create(int(device), int(capacity), ctypes.byref(self._h))
        if rc != 0:
            self._h = None
            raise _lib.UsrtError(rc, "usrt_create(device=%d, capacity=%d) failed -- a CUDA device is required"
                                 % (device, capacity))
        self.device = device
        self.capacity = int(capacity)

    # -- plumbing ------------------------------------------------------------------------------
    def _check(self, rc):
        if rc != 0:
            raise _lib.UsrtError(rc, self._lib.usrt_last_error(self._h).decode())

    def close(self):
        if getattr(self, "_h", None):
            for p in list(getattr(self, "_pinned", {}).values()):        # buffers of host_alloc nobody gave back
                self._lib.usrt_host_free(self._h, ctypes.c_void_p(p))
            self._pinned = {}
            self._lib.usrt_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def sync(self):
        self._check(self._lib.usrt_sync(self._h))

    def set_stream(self, cuda_stream):
        self._check(self._lib.usrt_set_stream(self._h, ctypes.c_void_p(cuda_stream)))

    def use_torch_stream(self, stream=None):
        """Enqueue on a torch CUDA stream (default: torch's current stream on this device). torch's
        default stream has handle 0, which the ABI reads as "back to the context's own stream", so it is
        passed as cudaStreamLegacy (0x1) instead."""
        import torch
        if stream is None:
            stream = torch.cuda.current_stream(self.device)
        self.set_stream(stream.cuda_stream or 1)

    def set_key_mode(self, mode):
        """_lib.KEYS_REFERENCE (default) / KEYS_INDEX_TIEBREAK / KEYS_MORTON64 (include/usrt.h usrt_key_mode)."""
        self._check(self._lib.usrt_set_key_mode(self._h, int(mode)))

    def set_world_bounds(self, whole_min, whole_max):
        self._check(self._lib.usrt_set_world_bounds(self._h, whole_min, whole_max))

    def set_world_box(self, box_min, box_max):
        lo = np.ascontiguousarray(box_min, np.float32).reshape(3); hi = np.ascontiguousarray(box_max, np.float32).reshape(3)
        self._check(self._lib.usrt_set_world_box(self._h, _ptr(lo), _ptr(hi)))

    def fit_world_box(self):
        """Opt-in (MeshBufferContainer.cs:7 TODO): NormalizeCentroid's box becomes the per-axis min/max of the
        uploaded vertices, reduced on the device. Returns (min[3], max[3])."""
        lo, hi = np.zeros(3, np.float32), np.zeros(3, np.float32)
        self._check(self._lib.usrt_fit_world_box(self._h, _ptr(lo), _ptr(hi)))
        return lo, hi

    @property
    def triangles_length(self):
        return int(self._lib.usrt_triangles_length(self._h))

    @property
    def kernel_launches(self):
        return int(self._lib.usrt_kernel_launches(self._h))

    # -- stages --------------------------------------------------------------------------------
    def upload_triangles(self, tris):
        tris = np.ascontiguousarray(tris, dtype=Triangle)
        self._check(self._lib.usrt_upload_triangles(self._h, _ptr(tris), len(tris)))

    def host_alloc(self, nbytes):
        """Page-locked host memory from the library (usrt_host_alloc = cudaHostAlloc) as a uint8 numpy array; give it
        back with host_free(array) once nothing enqueued still uses it. Faster to upload from than pages pinned after the
        fact, and the only way to get pinned memory for a host that has no CUDA binding of its own."""
        p = ctypes.c_void_p()
        self._check(self._lib.usrt_host_alloc(self._h, int(nbytes), ctypes.byref(p)))
        buf = (ctypes.c_uint8 * int(nbytes)).from_address(p.value)
        arr = np.frombuffer(buf, dtype=np.uint8)
        self._pinned = getattr(self, "_pinned", {})
        self._pinned[arr.ctypes.data] = p.value
        return arr

    def host_free(self, arr):
        base = arr if isinstance(arr, int) else arr.ctypes.data
        p = getattr(self, "_pinned", {}).pop(base, None)
        if p is None:
            raise ValueError("host_free: not a buffer of host_alloc")
        self._check(self._lib.usrt_host_free(self._h, ctypes.c_void_p(p)))

    def upload_triangles_async(self, pinned_tris):
        """pinned_tris: a Triangle array over page-locked memory (e.g. a view of a torch pin_memory() tensor);
        it must stay untouched until sync()."""
        if pinned_tris.dtype != Triangle or not pinned_tris.flags["C_CONTIGUOUS"]:
            raise ValueError("upload_triangles_async needs a contiguous Triangle array")
        self._check(self._lib.usrt_upload_triangles_async(self._h, _ptr(pinned_tris), len(pinned_tris)))

    def upload_positions(self, positions, pinned=False):
        """(n, 12) float32: the first 48 bytes of every Triangle (a.xyz, pad, b.xyz, pad, c.xyz, pad). pinned=True:
        page-locked memory, asynchronous (untouched until sync())."""
        pos = positions if pinned else np.ascontiguousarray(positions, np.float32)
        assert pos.dtype == np.float32 and pos.flags["C_CONTIGUOUS"] and pos.size % 12 == 0
        fn = self._lib.usrt_upload_positions_async if pinned else self._lib.usrt_upload_positions
        self._check(fn(self._h, _ptr(pos), pos.size // 12))

    def set_triangles_device(self, dev_ptr, n):
        self._check(self._lib.usrt_set_triangles_device(self._h, ctypes.c_void_p(dev_ptr), n))

    def morton(self):
        self._check(self._lib.usrt_morton(self._h))

    def sort(self):
        self._check(self._lib.usrt_sort(self._h))

    def sort_pairs_host(self, keys, values=None):
        """In place on contiguous uint32 numpy arrays."""
        assert keys.dtype == np.uint32 and keys.flags.c_contiguous
        if values is not None:
            assert values.dtype == np.uint32 and values.flags.c_contiguous and len(values) == len(keys)
        self._check(self._lib.usrt_sort_pairs_host(self._h, _ptr(keys), _ptr(values), len(keys)))

    def sort_pairs64_host(self, keys, values=None):
        """ComputeBufferSorter<ulong, uint>: in place on contiguous uint64 keys / uint32 values."""
        assert keys.dtype == np.uint64 and keys.flags.c_contiguous
        if values is not None:
            assert values.dtype == np.uint32 and values.flags.c_contiguous and len(values) == len(keys)
        self._check(self._lib.usrt_sort_pairs64_host(self._h, _ptr(keys), _ptr(values), len(keys)))

    def sort_pairs64_device(self, keys_ptr, values_ptr, count):
        self._check(self._lib.usrt_sort_pairs64_device(self._h, ctypes.c_void_p(keys_ptr),
                                                       ctypes.c_void_p(values_ptr) if values_ptr else None, count))

    def sort_pairs_device(self, keys_ptr, values_ptr, count):
        self._check(self._lib.usrt_sort_pairs_device(self._h, ctypes.c_void_p(keys_ptr),
                                                     ctypes.c_void_p(values_ptr) if values_ptr else None, count))

    def partition_pass_device(self, src_keys, src_vals, dst_keys, dst_vals, count, bit_offset, hist_ptr=None):
        vp = ctypes.c_void_p
        self._check(self._lib.usrt_partition_pass_device(self._h, vp(src_keys), vp(src_vals) if src_vals else None,
                                                         vp(dst_keys), vp(dst_vals) if dst_vals else None, count,
                                                         bit_offset, vp(hist_ptr) if hist_ptr else None))

    def digit_histogram_device(self, keys_ptr, count, bit_offset, hist_ptr):
        vp = ctypes.c_void_p
        self._check(self._lib.usrt_digit_histogram_device(self._h, vp(keys_ptr), count, bit_offset, vp(hist_ptr)))

    def partition_scatter_device(self, src_keys, src_vals, count, bit_offset, key_base_ptr, value_base_ptr):
        vp = ctypes.c_void_p
        self._check(self._lib.usrt_partition_scatter_device(self._h, vp(src_keys), vp(src_vals), count, bit_offset,
                                                            vp(key_base_ptr), vp(value_base_ptr)))

    def peer_scatter_plan_device(self, all_hist_ptr, world, rank, peer_base_ptr, capacity, key_base_ptr, value_base_ptr,
                                 recv_total_ptr, bounds_ptr=None):
        """The bucket-exchange landing plan from the all-gathered histograms, device in / device out."""
        vp = ctypes.c_void_p
        self._check(self._lib.usrt_peer_scatter_plan_device(self._h, vp(all_hist_ptr), int(world), int(rank), vp(peer_base_ptr),
                                                            int(capacity), vp(key_base_ptr), vp(value_base_ptr),
                                                            vp(recv_total_ptr), vp(bounds_ptr) if bounds_ptr else None))

    def peer_buffer_create(self, nbytes):
        """-> (device pointer, 64-byte IPC handle) of a buffer other processes on the node can map."""
        ptr = ctypes.c_void_p()
        handle = (ctypes.c_ubyte * 64)()
        self._check(self._lib.usrt_peer_buffer_create(self._h, nbytes, ctypes.byref(ptr), handle))
        return ptr.value, bytes(handle)

    def peer_buffer_open(self, handle):
        ptr = ctypes.c_void_p()
        buf = (ctypes.c_ubyte * 64).from_buffer_copy(handle)
        self._check(self._lib.usrt_peer_buffer_open(self._h, buf, ctypes.byref(ptr)))
        return ptr.value

    def peer_buffer_close(self, ptr, opened):
        self._check(self._lib.usrt_peer_buffer_close(self._h, ctypes.c_void_p(ptr), 1 if opened else 0))

    def set_hit_mirrors(self, dev_ptrs):
        """Peer frame slots every traced hit record is also stored to (usrt_set_hit_mirrors); [] clears."""
        arr = (ctypes.c_void_p * max(len(dev_ptrs), 1))(*[int(p) for p in dev_ptrs])
        self._check(self._lib.usrt_set_hit_mirrors(self._h, len(dev_ptrs), arr))

    def distribute_keys(self):
        self._check(self._lib.usrt_distribute_keys(self._h))

    def construct_tree(self):
        self._check(self._lib.usrt_construct_tree(self._h))

    def construct_bvh(self):
        self._check(self._lib.usrt_construct_bvh(self._h))

    def rebuild(self):
        self._check(self._lib.usrt_rebuild(self._h))

    def enable_stage_timing(self, enabled=True):
        self._check(self._lib.usrt_enable_stage_timing(self._h, int(enabled)))

    def last_rebuild_ms(self):
        out = (ctypes.c_float * 6)()
        self._check(self._lib.usrt_last_rebuild_ms(self._h, out))
        return dict(zip(("morton", "sort", "distribute", "tree", "bvh", "total"), [float(x) for x in out]))

    def last_sort_ms(self):
        out = (ctypes.c_float * 6)()
        self._check(self._lib.usrt_last_sort_ms(self._h, out))
        return dict(zip(("histogram", "pass0", "pass8", "pass16", "pass24", "total"), [float(x) for x in out]))

    def set_trace_mode(self, mode):
        self._check(self._lib.usrt_set_trace_mode(self._h, int(mode)))

    def trace_primary(self, width, height, near, tan_half_fov, cam_to_world, y0=0, y1=None, download=True, out=None):
        y1 = height if y1 is None else y1
        m = np.ascontiguousarray(cam_to_world, np.float32).reshape(16)
        if download and out is None:
            out = np.zeros(width * height, RaycastResult)
        self._check(self._lib.usrt_trace_primary(self._h, width, height, float(near), float(tan_half_fov), _ptr(m),
                                                 y0, y1, _ptr(out) if download else None))
        return out

    def trace_primary_async(self, width, height, near, tan_half_fov, cam_to_world, pinned_out):
        """Enqueue a full-frame trace whose records land in `pinned_out` (page-locked RaycastResult array);
        readable after sync()."""
        m = np.ascontiguousarray(cam_to_world, np.float32).reshape(16)
        self._check(self._lib.usrt_trace_primary_async(self._h, width, height, float(near), float(tan_half_fov),
                                                       _ptr(m), _ptr(pinned_out)))

    def diffuse_rays_device(self, width, height, near, tan_half_fov, cam_to_world, seed, first_sample, num_samples,
                            rays_out_ptr, primary_hits_ptr=None):
        """Bounce rays of samples [first_sample, +num_samples) from the primary hit records (usrt_diffuse_rays_device);
        rays_out_ptr: device memory for num_samples * width * height rays of 32 bytes."""
        m = np.ascontiguousarray(cam_to_world, np.float32).reshape(16)
        self._check(self._lib.usrt_diffuse_rays_device(self._h, width, height, float(near), float(tan_half_fov), _ptr(m),
                                                       ctypes.c_void_p(primary_hits_ptr) if primary_hits_ptr else None,
                                                       int(seed), int(first_sample), int(num_samples),
                                                       ctypes.c_void_p(rays_out_ptr)))

    def trace_primary_sharded(self, width, height, near, tan_half_fov, cam_to_world, block_rows, shard, num_shards,
                              dev_out=None, download=False):
        """One launch over this shard's interleaved row blocks; compact output (see include/usrt.h)."""
        m = np.ascontiguousarray(cam_to_world, np.float32).reshape(16)
        blocks = -(-height // block_rows)
        local_rows = -(-blocks // num_shards) * block_rows
        out = np.zeros(local_rows * width, RaycastResult) if download else None
        self._check(self._lib.usrt_trace_primary_sharded(self._h, width, height, float(near), float(tan_half_fov),
                                                         _ptr(m), block_rows, shard, num_shards,
                                                         ctypes.c_void_p(dev_out) if dev_out else None, _ptr(out)))
        return out

    def upload_texture(self, texture_rgba):
        """(H, W, 4) float32 texels, row 0 at v = 0 (the reference's _meshTexture)."""
        tex = np.ascontiguousarray(texture_rgba, np.float32)
        assert tex.ndim == 3 and tex.shape[2] == 4
        self._check(self._lib.usrt_upload_texture(self._h, _ptr(tex), tex.shape[1], tex.shape[0]))

    def shade(self, count=None):
        """Raytracing.compute:178-184 over the last trace's hit records -> (count, 4) float16 (RGBA16F)."""
        _, n = self.hits_device()
        out = np.zeros((n, 4), np.float16)
        self._check(self._lib.usrt_shade(self._h, None, _ptr(out)))
        return out

    def trace_rays(self, rays, out=None):
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 8)
        if out is None:
            out = np.zeros(len(rays), RaycastResult)
        self._check(self._lib.usrt_trace_rays(self._h, _ptr(rays), len(rays), _ptr(out)))
        return out

    def trace_rays_device(self, rays_ptr, num_rays, out_ptr=None):
        self._check(self._lib.usrt_trace_rays_device(self._h, ctypes.c_void_p(rays_ptr), num_rays,
                                                     ctypes.c_void_p(out_ptr) if out_ptr else None))

    def hits_device(self):
        p = ctypes.c_void_p(); c = ctypes.c_uint64()
        self._check(self._lib.usrt_hits_device(self._h, ctypes.byref(p), ctypes.byref(c)))
        return p.value, int(c.value)

    def download(self, buffer, count=None):
        count = self.triangles_length if count is None else int(count)
        out = np.zeros(count, _BUF_DTYPES[buffer])
        if count:
            self._check(self._lib.usrt_download(self._h, buffer, _ptr(out), count))
        return out

    def device_ptr(self, buffer):
        p = ctypes.c_void_p()
        self._check(self._lib.usrt_device_ptr(self._h, buffer, ctypes.byref(p)))
        return p.value

    def count_corrupted_nodes(self):
        a = ctypes.c_uint32(); b = ctypes.c_uint32()
        self._check(self._lib.usrt_count_corrupted_nodes(self._h, ctypes.byref(a), ctypes.byref(b)))
        return int(a.value), int(b.value)


# ==================================================================================================
# Mirrors of the reference's C# classes
# ==================================================================================================
class DeviceBuffer:
    """Stands in for a Unity ComputeBuffer handle: names one scene buffer of a context
    (Assets/_Scripts/DataBuffer.cs:7 DeviceBuffer). GetData() = DataBuffer.GetData (:50-54)."""

    def __init__(self, ctx, buffer):
        self.ctx, self.buffer = ctx, buffer

    def GetData(self, count=None):
        return self.ctx.download(self.buffer, count)


def merge_meshes(meshes):
    """Several Triangle arrays -> (one Triangle array, offsets[len(meshes) + 1]): mesh k owns the triangle indices
    [offsets[k], offsets[k + 1]) of the merged scene (the "TODO multiple meshes" of MeshBufferContainer.cs:96)."""
    parts = [np.ascontiguousarray(m, dtype=Triangle).reshape(-1) for m in meshes]
    offsets = np.concatenate([[0], np.cumsum([len(p) for p in parts])]).astype(np.int64)
    merged = np.concatenate(parts) if parts else np.zeros(0, Triangle)
    return merged, offsets


def mesh_of_triangle(offsets, triangle_index):
    """(mesh id, triangle index inside that mesh) of merged-scene triangle indices (e.g. hit records'
    triangleIndex); vectorised."""
    idx = np.asarray(triangle_index, np.int64)
    mesh = np.searchsorted(np.asarray(offsets, np.int64), idx, side="right") - 1
    return mesh, idx - np.asarray(offsets, np.int64)[mesh]


class MeshBufferContainer:
    """Assets/_Scripts/MeshBufferContainer.cs. `mesh` is Triangle[n] (the packing loop :117-146 is mesh
    ingest) or a list of such arrays (:96 "TODO multiple meshes": they are merged into one scene and
    `MeshOffsets` / `MeshOfTriangle` map hit records back); Morton codes, indices and triangle AABBs are
    computed on the GPU (K1) instead of the reference's CPU loop."""

    def __init__(self, mesh, capacity=None, device=0, ctx=None, fitWorldBox=False):
        if isinstance(mesh, (list, tuple)):
            mesh, self.MeshOffsets = merge_meshes(mesh)
        else:
            self.MeshOffsets = np.array([0, len(mesh)], np.int64)
        mesh = np.ascontiguousarray(mesh, dtype=Triangle)
        self._ctx = ctx if ctx is not None else Context(capacity if capacity else max(len(mesh), 2), device)
        self._ctx.upload_triangles(mesh)           # :148-151 Sync()
        if fitWorldBox:                            # :7 TODO, opt-in: keys no longer match the fixed +-125 cube
            self.Whole = self._ctx.fit_world_box()
        self._ctx.morton()                         # :123-146
        self.local = {}

    @property
    def ctx(self):
        return self._ctx

    Keys = property(lambda s: DeviceBuffer(s._ctx, _lib.BUF_KEYS))                     # :17
    TriangleIndex = property(lambda s: DeviceBuffer(s._ctx, _lib.BUF_TRIANGLE_INDEX))  # :19
    TriangleData = property(lambda s: DeviceBuffer(s._ctx, _lib.BUF_TRIANGLE_DATA))    # :20
    TriangleAABB = property(lambda s: DeviceBuffer(s._ctx, _lib.BUF_TRIANGLE_AABB))    # :21
    BvhData = property(lambda s: DeviceBuffer(s._ctx, _lib.BUF_BVH_DATA))              # :22
    BvhLeafNode = property(lambda s: DeviceBuffer(s._ctx, _lib.BUF_LEAF_NODES))        # :23
    BvhInternalNode = property(lambda s: DeviceBuffer(s._ctx, _lib.BUF_INTERNAL_NODES))  # :24

    def MeshOfTriangle(self, triangleIndex):
        return mesh_of_triangle(self.MeshOffsets, triangleIndex)

    @property
    def TrianglesLength(self):                     # :30
        return self._ctx.triangles_length

    def DistributeKeys(self):                      # :154-169, on the GPU (K3), no readback
        self._ctx.distribute_keys()

    def GetAllGpuData(self):
        """:171-196 -- read everything back and run the NullLeaf corruption check. Returns the number
        of (leaf, internal) entries still equal to NullLeaf (the reference logs an error per entry)."""
        n = self.TrianglesLength
        self.local = dict(
            keys=self.Keys.GetData(n), triangleIndex=self.TriangleIndex.GetData(n),
            triangleData=self.TriangleData.GetData(n), triangleAABB=self.TriangleAABB.GetData(n),
            bvhData=self.BvhData.GetData(n), leafNodes=self.BvhLeafNode.GetData(n),
            internalNodes=self.BvhInternalNode.GetData(n))
        leaf, internal = self.local["leafNodes"], self.local["internalNodes"][:n - 1]
        bad_leaf = int(((leaf["index"] == 0xFFFFFFFF) & (leaf["parent"] == 0xFFFFFFFF)).sum())
        bad_int = int(((internal["index"] == 0xFFFFFFFF) & (internal["parent"] == 0xFFFFFFFF)).sum())
        return bad_leaf, bad_int

    def Dispose(self):                             # :207-216
        self._ctx.close()


class ComputeBufferSorter:
    """Assets/_Scripts/ComputeBufferSorter.cs: ComputeBufferSorter<uint,uint>(dataLength, keys, values,
    shaders) + Sort(). keys/values are either the container's DeviceBuffers (the reference's use,
    RaytracingMeshDrawer.cs:36) or caller-owned uint32 numpy arrays (sorted in place)."""

    def __init__(self, dataLength, keys, values, ctx=None):
        self._n = int(dataLength)
        self._keys, self._values = keys, values
        if isinstance(keys, DeviceBuffer):
            if not (isinstance(values, DeviceBuffer) and values.ctx is keys.ctx):
                raise ValueError("keys and values must belong to the same container")
            if keys.buffer != _lib.BUF_KEYS or values.buffer != _lib.BUF_TRIANGLE_INDEX:
                raise ValueError("device-side Sort() is bound to the container's Keys / TriangleIndex buffers")
            self._ctx = keys.ctx
        else:
            self._ctx = ctx if ctx is not None else Context(max(self._n, 2))
            if self._n != len(keys):
                raise ValueError("dataLength != len(keys)")

    def Sort(self):                                # :100-126 (4 passes, bitOffset 0,8,16,24)
        if isinstance(self._keys, DeviceBuffer):
            if self._n != self._ctx.triangles_length:
                raise ValueError("dataLength != container.TrianglesLength")
            self._ctx.sort()
        else:
            self._ctx.sort_pairs_host(self._keys, self._values)

    def Dispose(self):
        pass


class BVHConstructor:
    """Assets/_Scripts/BVHConstructor.cs:24-69. The buffers must be the container's (the reference binds
    exactly those, RaytracingMeshDrawer.cs:41-48)."""

    def __init__(self, trianglesCount, sortedMortonCodes, sortedTriangleIndices, triangleAABB, internalNodes,
                 leafNodes, BVHData):
        bufs = (sortedMortonCodes, sortedTriangleIndices, triangleAABB, internalNodes, leafNodes, BVHData)
        want = (_lib.BUF_KEYS, _lib.BUF_TRIANGLE_INDEX, _lib.BUF_TRIANGLE_AABB, _lib.BUF_INTERNAL_NODES,
                _lib.BUF_LEAF_NODES, _lib.BUF_BVH_DATA)
        if any(not isinstance(b, DeviceBuffer) for b in bufs) or any(b.ctx is not bufs[0].ctx for b in bufs):
            raise ValueError("BVHConstructor takes the DeviceBuffers of one MeshBufferContainer")
        if tuple(b.buffer for b in bufs) != want:
            raise ValueError("BVHConstructor buffers are out of order")
        self._ctx = bufs[0].ctx
        if int(trianglesCount) != self._ctx.triangles_length:
            raise ValueError("trianglesCount != container.TrianglesLength")

    def ConstructTree(self):                       # :61-64
        self._ctx.construct_tree()

    def ConstructBVH(self):                        # :66-69
        self._ctx.construct_bvh()

    def Dispose(self):
        pass


class RaytracingMeshDrawer:
    """The build-once / trace-per-frame sequence of Assets/_Scripts/RaytracingMeshDrawer.cs.
    Awake() = :30-54 step by step (as the reference dispatches it); Rebuild() = the same stages as one
    fused enqueue; Update() = :76-84 returning the hit records instead of shading a texture."""

    def __init__(self, mesh, capacity=None, device=0):
        self._mesh, self._capacity, self._device = mesh, capacity, device
        self.container = None

    def Awake(self):
        self.container = MeshBufferContainer(self._mesh, self._capacity, self._device)
        c = self.container
        self.sorter = ComputeBufferSorter(c.TrianglesLength, c.Keys, c.TriangleIndex)
        self.sorter.Sort()
        c.DistributeKeys()
        self.bvhConstructor = BVHConstructor(c.TrianglesLength, c.Keys, c.TriangleIndex, c.TriangleAABB,
                                             c.BvhInternalNode, c.BvhLeafNode, c.BvhData)
        self.bvhConstructor.ConstructTree()
        self.bvhConstructor.ConstructBVH()
        return self

    def Rebuild(self):
        self.container.ctx.rebuild()

    def Update(self, screenWidth, screenHeight, near, cameraFov, cameraToWorldMatrix, y0=0, y1=None, download=True):
        """cameraFov = tan(fieldOfView * Deg2Rad / 2) (:80); near = _ProjectionParams.y."""
        return self.container.ctx.trace_primary(screenWidth, screenHeight, near, cameraFov, cameraToWorldMatrix,
                                                y0, y1, download)

    def OnDestroy(self):                           # :118-123
        if self.container:
            self.container.Dispose()
