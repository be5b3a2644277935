"""ctypes binding of the C-ABI in include/xpcs_b200.h (libxpcs_b200.so).

There is no fallback: if the CUDA library is missing or cannot be loaded this module
raises, and every compute call raises XpcsError when the library reports a failure
(including "no CUDA device").
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libxpcs_b200.so")

XPCS_COMPAT_STALE_TAIL = 1
XPCS_COMPAT_LATE_WINDOW = 2
XPCS_FLAG_LANE_MULTITAU = 0x100
XPCS_FLAG_SCALAR_DENSE = 0x200


class XpcsError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("xpcs_b200 error %d: %s" % (code, msg))
        self.code = code


class XpcsParams(C.Structure):
    _fields_ = [
        ("struct_size", C.c_int32),
        ("width", C.c_int32), ("height", C.c_int32),
        ("frames", C.c_int32),
        ("delays_per_level", C.c_int32),
        ("stride_frames", C.c_int32), ("avg_frames", C.c_int32),
        ("static_window", C.c_int32),
        ("normalize_by_framesum", C.c_int32),
        ("compat_flags", C.c_uint32),
        ("lld", C.c_float), ("sigma", C.c_float),
        ("dqmap", C.c_void_p), ("sqmap", C.c_void_p), ("flatfield", C.c_void_p),
        ("shard_index", C.c_int32), ("shard_count", C.c_int32),
        ("reserve_events", C.c_int64),
    ]


class XpcsInfo(C.Structure):
    _fields_ = [
        ("n_delays", C.c_int32), ("max_level", C.c_int32),
        ("n_static", C.c_int32), ("n_dynamic", C.c_int32),
        ("n_segments", C.c_int32), ("n_rows", C.c_int32), ("n_rows_total", C.c_int32),
        ("raw_frames_seen", C.c_int32),
        ("events_pushed", C.c_int64), ("events_stored", C.c_int64), ("store_words", C.c_int64),
        ("value_kind", C.c_int32), ("max_row_events", C.c_int32),
    ]


class XpcsShardPlan(C.Structure):
    _fields_ = [("n_static", C.c_int32), ("n_dynamic", C.c_int32), ("n_segments", C.c_int32),
                ("seg_first", C.c_int32), ("seg_last", C.c_int32), ("n_rows", C.c_int32),
                ("n_rows_total", C.c_int32), ("n_delays", C.c_int32)]


# every symbol include/xpcs_b200.h declares: name -> (restype, argtypes)
_vp, _i, _i64 = C.c_void_p, C.c_int, C.c_int64
SYMBOLS = {
    "xpcs_level_max": (_i, [_i, _i]),
    "xpcs_delay_schedule": (_i, [_i, _i, _vp, _vp, _i]),
    "xpcs_plan_shard": (_i, [C.POINTER(XpcsParams), C.POINTER(XpcsShardPlan), _vp, _i64]),
    "xpcs_create": (_i, [C.POINTER(XpcsParams), _i, C.POINTER(_vp)]),
    "xpcs_destroy": (None, [_vp]),
    "xpcs_last_error": (C.c_char_p, [_vp]),
    "xpcs_get_info": (_i, [_vp, C.POINTER(XpcsInfo)]),
    "xpcs_get_row_pixels": (_i, [_vp, _vp]),
    "xpcs_set_stream": (_i, [_vp, _vp]),
    "xpcs_reset": (_i, [_vp]),
    "xpcs_set_dark": (_i, [_vp, _vp, _i]),
    "xpcs_get_dark": (_i, [_vp, _vp, _vp]),
    "xpcs_push_sparse": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i]),
    "xpcs_push_sparse_device": (_i, [_vp, _vp, _vp, _vp, _i64, _i]),
    "xpcs_push_dense": (_i, [_vp, _vp, _vp, _vp, _i]),
    "xpcs_push_dense_device": (_i, [_vp, _vp, _i]),
    "xpcs_finish_ingest": (_i, [_vp, _vp, _vp, _vp, _vp]),
    "xpcs_stream_begin": (_i, [_vp, _i]),
    "xpcs_stream_push_sparse": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i]),
    "xpcs_stream_push_sparse_device": (_i, [_vp, _vp, _vp, _vp, _i64, _i]),
    "xpcs_stream_finish": (_i, [_vp, _vp, _vp, _vp, _vp]),
    "xpcs_get_timestamps": (_i, [_vp, _vp, _vp]),
    "xpcs_get_frames": (_i, [_vp, C.c_int, _vp]),
    "xpcs_multitau": (_i, [_vp, _vp, _vp, _vp]),
    "xpcs_get_correlators": (_i, [_vp, _vp, _i, _vp, _vp, _vp]),
    "xpcs_normalize": (_i, [_vp, _vp, _vp]),
    "xpcs_normalize_partials": (_i, [_vp, C.POINTER(_vp), C.POINTER(_i64)]),
    "xpcs_normalize_finish": (_i, [_vp, _vp, _vp]),
    "xpcs_comm_unique_id": (_i, [_vp]),
    "xpcs_comm_init": (_i, [_vp, _i, _i, _vp]),
    "xpcs_comm_nccl_version": (_i, []),
    "xpcs_comm_transport": (_i, [_vp]),
    "xpcs_push_sparse_slab": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _vp, _i]),
    "xpcs_push_sparse_slab_device": (_i, [_vp, _i, _vp, _vp, _vp, _i64, _i]),
    "xpcs_twotime": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "xpcs_twotime_sg": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, C.POINTER(C.c_int)]),
    "xpcs_host_alloc": (_vp, [C.c_size_t]),
    "xpcs_host_free": (None, [_vp]),
    "xpcs_kernel_timing": (_i, [_vp, _i]),
    "xpcs_launch_count": (_i64, [_vp]),
    "xpcs_multitau_fallback_slices": (_i64, [_vp]),
    "xpcs_kernel_report": (_i, [_vp, _vp, _vp, _vp, _i]),
    "xpcs_kernel_report_reset": (_i, [_vp]),
    "xpcs_abi_version": (_i, []),
    "xpcs_compiled_arch": (_i, []),
}

_lib = None


def _pick_nccl():
    """The library binds NCCL at run time (csrc/comm.cu: an NCCL already in the process, else $XPCS_NCCL_LIB, else
    the system libnccl.so.2).  In a Python process PyTorch may be imported LATER and would then be handed whatever
    libnccl.so.2 is already loaded, so point the library at the copy PyTorch itself ships (no torch import here)."""
    if os.environ.get("XPCS_NCCL_LIB"):
        return
    try:
        import importlib.util
        spec = importlib.util.find_spec("nvidia")
        for base in (spec.submodule_search_locations if spec else []):
            cand = os.path.join(base, "nccl", "lib", "libnccl.so.2")
            if os.path.exists(cand):
                os.environ["XPCS_NCCL_LIB"] = cand
                return
    except Exception:
        pass


def load():
    """Load libxpcs_b200.so; raises (never falls back) when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C xpcs-eigen_b200 lib`). There is no CPU fallback." % LIB_PATH)
    _pick_nccl()
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError = header/library drift
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
