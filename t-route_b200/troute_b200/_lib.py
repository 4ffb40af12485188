"""ctypes binding of libtroute_b200.so (C ABI: include/troute_b200.h).

The product path has NO CPU fallback: if the CUDA library is missing or cannot be loaded this module
raises, and every entry point of the package fails with it.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TROUTE_B200_LIB") or os.path.join(_HERE, "lib", "libtroute_b200.so")

TRT_KIND_MC = 0
TRT_KIND_LEVELPOOL = 1
TRT_KIND_BOUNDARY = 2

TRT_ERR_INVALID = -1
TRT_ERR_CYCLE = -2
TRT_ERR_CUDA = -3
TRT_ERR_NOMEM = -4
TRT_ERR_STATE = -5


class TrouteB200Error(RuntimeError):
    """CUDA / state failure reported by libtroute_b200."""


_lib = None

# every symbol include/troute_b200.h declares: (name, restype, argtypes)
_i64p = C.POINTER(C.c_int64)
_i32p = C.POINTER(C.c_int32)
_f32p = C.POINTER(C.c_float)
_f64p = C.POINTER(C.c_double)
_u8p = C.POINTER(C.c_uint8)
_net = C.c_void_p

SYMBOLS = [
    ("trt_last_error", C.c_char_p, []),
    ("trt_version", C.c_int, []),
    ("trt_device_count", C.c_int, []),
    ("trt_network_create", C.c_int, [C.c_int, C.c_int64, _i64p, _i64p, _u8p, _f32p, C.c_int32, _i32p, C.POINTER(_net)]),
    ("trt_network_create_ex", C.c_int, [C.c_int, C.c_int64, _i64p, _i64p, _u8p, _f32p, C.c_int32, _i32p, _i32p, C.POINTER(_net)]),
    ("trt_network_create_ordered", C.c_int, [C.c_int, C.c_int64, _i64p, _i64p, _u8p, _f32p, C.c_int32, _i32p, _i32p, _i32p, C.POINTER(_net)]),
    ("trt_trip_counts", C.c_int, [_net, _i32p]),
    ("trt_trip_counts_bucketed", C.c_int, [_net, C.c_int32, _i32p]),
    ("trt_overbank_counts", C.c_int, [_net, _i32p]),
    ("trt_network_destroy", C.c_int, [_net]),
    ("trt_network_num_levels", C.c_int, [_net, _i32p]),
    ("trt_network_get_levels", C.c_int, [_net, _i32p]),
    ("trt_network_get_positions", C.c_int, [_net, _i32p]),
    ("trt_network_set_levelpools", C.c_int, [_net, C.c_int64, _i64p, _f64p, C.c_float]),
    ("trt_network_set_gages", C.c_int, [_net, C.c_int64, _i64p, _u8p, _f32p, C.c_int32, _f32p, _f32p, C.c_float, C.c_float]),
    ("trt_download_gages", C.c_int, [_net, _f32p, _f32p, _f32p]),
    ("trt_upload_forcing", C.c_int, [_net, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_int64, _i64p, C.c_void_p]),
    ("trt_continue", C.c_int, [_net, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_int64, _i64p, C.c_void_p]),
    ("trt_network_update_gage_observations", C.c_int, [_net, C.c_void_p, C.c_int32]),
    ("trt_result_hash", C.c_int, [_net, C.c_int64, _i64p, _i64p, C.POINTER(C.c_uint64)]),
    ("trt_download_rows", C.c_int, [_net, C.c_int64, _i64p, C.c_void_p]),
    ("trt_run", C.c_int, [_net, C.c_int32]),
    ("trt_run_async", C.c_int, [_net, C.c_int32]),
    ("trt_sync", C.c_int, [_net]),
    ("trt_download_results", C.c_int, [_net, C.c_void_p, C.c_void_p]),
    ("trt_download_levelpool_inflow", C.c_int, [_net, C.c_void_p]),
    ("trt_download_last_step", C.c_int, [_net, C.c_void_p]),
    ("trt_run_download", C.c_int, [_net, C.c_int32, C.c_void_p, C.c_void_p]),
    ("trt_route", C.c_int, [_net, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_int64, _i64p, C.c_void_p, C.c_void_p, C.c_void_p]),
    ("trt_export_flow_series", C.c_int, [_net, C.c_int64, _i64p, C.c_void_p]),
    ("trt_import_boundary_flow", C.c_int, [_net, C.c_int64, _i64p, C.c_void_p]),
    ("trt_device_results", C.c_int, [_net, C.POINTER(C.c_void_p)]),
    ("trt_network_state_ptr", C.c_int, [_net, C.POINTER(C.c_void_p)]),
    ("trt_ipc_get_handle", C.c_int, [C.c_void_p, C.c_void_p]),
    ("trt_ipc_open_handle", C.c_int, [C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]),
    ("trt_ipc_close_handle", C.c_int, [C.c_void_p]),
    ("trt_network_set_peer", C.c_int, [_net, C.c_int32, C.c_void_p, C.c_int64]),
    ("trt_network_set_exports", C.c_int, [_net, C.c_int64, _i64p, _i32p, _i64p]),
    ("trt_network_set_imports", C.c_int, [_net, C.c_int64, _i64p]),
    ("trt_prepare", C.c_int, [_net]),
    ("trt_set_option", C.c_int, [_net, C.c_char_p, C.c_int64]),
    ("trt_stage_profile", C.c_int, [_net, C.c_int64, _f32p, _i64p, _i64p]),
    ("trt_last_run_phases", C.c_int, [_net, _f64p, _f64p, _i32p]),
    ("trt_march_profile", C.c_int, [_net, C.c_int64, C.POINTER(C.c_uint64), _i64p]),
    ("trt_last_run_stats", C.c_int, [_net, _f64p, _i64p, _i64p, _i64p]),
    ("trt_mc_segment_batch", C.c_int, [C.c_int, C.c_int64, _f32p, _f32p, _i32p]),
    ("trt_levelpool_series", C.c_int, [C.c_int, _f64p, C.c_int64, _f32p, C.c_float, C.c_float, _f32p, _f32p]),
    ("trt_powf_batch", C.c_int, [C.c_int, C.c_int64, _f32p, _f32p, _f32p]),
    ("trt_fdiv_batch", C.c_int, [C.c_int, C.c_int64, _f32p, _f32p, _f32p, C.POINTER(C.c_uint8)]),
    ("trt_host_alloc", C.c_int, [C.POINTER(C.c_void_p), C.c_uint64]),
    ("trt_host_free", C.c_int, [C.c_void_p]),
    ("trt_c_diffnw", C.c_int, [C.c_void_p] * 42),
    ("trt_diffnw_batch", C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    ("trt_diffusive_set_device", C.c_int, [C.c_int]),
    ("trt_diffusive_last_run", C.c_int, [_f64p, _f64p, C.POINTER(C.c_longlong)]),
]


def lib():
    """Load libtroute_b200.so (once).  Raises ImportError when the CUDA extension is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: the CUDA extension is not built (run `python -c 'import __graft_entry__ as g; "
            f"g.build()'` or `make -C t-route_b200/csrc`).  There is no CPU fallback for the routing path."
        )
    handle = C.CDLL(LIB_PATH)
    for name, restype, argtypes in SYMBOLS:
        fn = getattr(handle, name)
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = handle
    return _lib


def check(rc):
    """Translate a trt_status into the exception the reference raises for the same condition."""
    if rc >= 0:
        return rc
    msg = lib().trt_last_error().decode("utf-8", "replace")
    if rc in (TRT_ERR_INVALID, TRT_ERR_CYCLE):
        raise ValueError(msg)            # mc_reach.pyx:243-250 raises ValueError on shape mismatch
    if rc == TRT_ERR_NOMEM:
        raise MemoryError(msg)
    raise TrouteB200Error(msg)


def ptr(arr, ctype):
    return arr.ctypes.data_as(C.POINTER(ctype))


def as_c(arr, dtype):
    return np.ascontiguousarray(arr, dtype=dtype)
