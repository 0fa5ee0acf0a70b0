"""ctypes binding of libusrt_b200.so (include/usrt.h). No fallback: if the CUDA library is missing
or cannot be loaded this raises, and every compute call needs a CUDA device."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libusrt_b200.so")

USRT_OK = 0
ERRORS = {-1: "USRT_ERR_ARG", -2: "USRT_ERR_CUDA", -3: "USRT_ERR_STATE", -4: "USRT_ERR_NOMEM"}

# usrt_buffer
BUF_KEYS, BUF_TRIANGLE_INDEX, BUF_TRIANGLE_DATA, BUF_TRIANGLE_AABB, BUF_BVH_DATA, BUF_LEAF_NODES, \
    BUF_INTERNAL_NODES, BUF_KEYS64 = range(8)
# usrt_key_mode
KEYS_REFERENCE, KEYS_INDEX_TIEBREAK, KEYS_MORTON64 = range(3)

_c = ctypes
_P = _c.c_void_p
# name -> (restype, argtypes); must list every symbol include/usrt.h declares (tests check this).
SIGNATURES = {
    "usrt_create": (_c.c_int, [_c.c_int, _c.c_uint32, _c.POINTER(_P)]),
    "usrt_destroy": (_c.c_int, [_P]),
    "usrt_last_error": (_c.c_char_p, [_P]),
    "usrt_version": (_c.c_char_p, []),
    "usrt_sync": (_c.c_int, [_P]),
    "usrt_set_stream": (_c.c_int, [_P, _P]),
    "usrt_set_key_mode": (_c.c_int, [_P, _c.c_int]),
    "usrt_sort_pairs64_device": (_c.c_int, [_P, _P, _P, _c.c_uint64]),
    "usrt_sort_pairs64_host": (_c.c_int, [_P, _P, _P, _c.c_uint64]),
    "usrt_set_world_bounds": (_c.c_int, [_P, _c.c_float, _c.c_float]),
    "usrt_capacity": (_c.c_uint32, [_P]),
    "usrt_triangles_length": (_c.c_uint32, [_P]),
    "usrt_upload_triangles": (_c.c_int, [_P, _P, _c.c_uint32]),
    "usrt_upload_triangles_async": (_c.c_int, [_P, _P, _c.c_uint32]),
    "usrt_host_alloc": (_c.c_int, [_P, _c.c_uint64, _c.POINTER(_P)]),
    "usrt_host_free": (_c.c_int, [_P, _P]),
    "usrt_upload_positions": (_c.c_int, [_P, _P, _c.c_uint32]),
    "usrt_upload_positions_async": (_c.c_int, [_P, _P, _c.c_uint32]),
    "usrt_set_triangles_device": (_c.c_int, [_P, _P, _c.c_uint32]),
    "usrt_upload_bvh": (_c.c_int, [_P, _c.c_uint32, _P, _P, _P, _P, _P, _P, _P]),
    "usrt_morton": (_c.c_int, [_P]),
    "usrt_sort": (_c.c_int, [_P]),
    "usrt_sort_pairs_device": (_c.c_int, [_P, _P, _P, _c.c_uint64]),
    "usrt_sort_pairs_host": (_c.c_int, [_P, _P, _P, _c.c_uint64]),
    "usrt_partition_pass_device": (_c.c_int, [_P, _P, _P, _P, _P, _c.c_uint64, _c.c_int, _P]),
    "usrt_digit_histogram_device": (_c.c_int, [_P, _P, _c.c_uint64, _c.c_int, _P]),
    "usrt_partition_scatter_device": (_c.c_int, [_P, _P, _P, _c.c_uint64, _c.c_int, _P, _P]),
    "usrt_peer_scatter_plan_device": (_c.c_int, [_P, _P, _c.c_int, _c.c_int, _P, _c.c_uint64, _P, _P, _P, _P]),
    "usrt_peer_buffer_create": (_c.c_int, [_P, _c.c_uint64, _c.POINTER(_P), _P]),
    "usrt_peer_buffer_open": (_c.c_int, [_P, _P, _c.POINTER(_P)]),
    "usrt_peer_buffer_close": (_c.c_int, [_P, _P, _c.c_int]),
    "usrt_set_hit_mirrors": (_c.c_int, [_P, _c.c_int, _c.POINTER(_P)]),
    "usrt_set_world_box": (_c.c_int, [_P, _P, _P]),
    "usrt_fit_world_box": (_c.c_int, [_P, _P, _P]),
    "usrt_distribute_keys": (_c.c_int, [_P]),
    "usrt_construct_tree": (_c.c_int, [_P]),
    "usrt_construct_bvh": (_c.c_int, [_P]),
    "usrt_rebuild": (_c.c_int, [_P]),
    "usrt_enable_stage_timing": (_c.c_int, [_P, _c.c_int]),
    "usrt_last_rebuild_ms": (_c.c_int, [_P, _c.POINTER(_c.c_float)]),
    "usrt_last_sort_ms": (_c.c_int, [_P, _c.POINTER(_c.c_float)]),
    "usrt_diffuse_rays_device": (_c.c_int, [_P, _c.c_int, _c.c_int, _c.c_float, _c.c_float, _P, _P, _c.c_uint64, _c.c_uint32,
                                            _c.c_uint32, _P]),
    "usrt_trace_primary_async": (_c.c_int, [_P, _c.c_int, _c.c_int, _c.c_float, _c.c_float, _P, _P]),
    "usrt_trace_primary": (_c.c_int, [_P, _c.c_int, _c.c_int, _c.c_float, _c.c_float, _P, _c.c_int, _c.c_int, _P]),
    "usrt_trace_primary_sharded": (_c.c_int, [_P, _c.c_int, _c.c_int, _c.c_float, _c.c_float, _P, _c.c_int, _c.c_int,
                                              _c.c_int, _P, _P]),
    "usrt_trace_rays": (_c.c_int, [_P, _P, _c.c_uint64, _P]),
    "usrt_trace_rays_device": (_c.c_int, [_P, _P, _c.c_uint64, _P]),
    "usrt_hits_device": (_c.c_int, [_P, _c.POINTER(_P), _c.POINTER(_c.c_uint64)]),
    "usrt_set_trace_mode": (_c.c_int, [_P, _c.c_int]),
    "usrt_upload_texture": (_c.c_int, [_P, _P, _c.c_int, _c.c_int]),
    "usrt_shade": (_c.c_int, [_P, _P, _P]),
    "usrt_download": (_c.c_int, [_P, _c.c_int, _P, _c.c_uint64]),
    "usrt_device_ptr": (_c.c_int, [_P, _c.c_int, _c.POINTER(_P)]),
    "usrt_count_corrupted_nodes": (_c.c_int, [_P, _c.POINTER(_c.c_uint32), _c.POINTER(_c.c_uint32)]),
    "usrt_kernel_launches": (_c.c_uint64, [_P]),
}

_lib = None


class UsrtError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("%s (%d): %s" % (ERRORS.get(code, "USRT_ERR"), code, message))
        self.code = code


def load():
    """Load the shared library and bind every declared entry point. Raises if the library was not
    built (run `python -m unitysimpleraytracing_b200.build`) -- there is deliberately no CPU path."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("libusrt_b200.so not built: run `python -m unitysimpleraytracing_b200.build` "
                              "(nvcc, sm_100a). There is no CPU fallback for this path.")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)      # AttributeError if the .so lacks a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib
