"""ctypes binding of libsodso_pr.so (include/sodso_pr.h).

The shared library is built in-tree by `make -C so_dso_place_recognition_b200/csrc` (or
`__graft_entry__.build()`).  There is no CPU fallback: a missing library or a missing sm_100
device raises, loudly.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libsodso_pr.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "sodso_pr.h")
DEBUG_HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "sodso_pr_debug.h")

SODSO_TYPE_SC = 0
SODSO_TYPE_M2DP = 1
SODSO_ALGO_TC = 0      # include/sodso_pr_debug.h
SODSO_ALGO_SIMT = 1
STATS_W = 6            # sodso_db_partial_stats row width
COMM_ID_BYTES = 128
SC_SIZE = 1200
M2DP_SIZE = 192

_lib = None

_vp = C.c_void_p
_i = C.c_int
_i64 = C.c_int64
_d = C.c_double

_PROTOS = {
    "sodso_ctx_create": (_i, [_i, C.POINTER(_vp)]),
    "sodso_ctx_destroy": (None, [_vp]),
    "sodso_ctx_stream": (_vp, [_vp]),
    "sodso_ctx_set_stream": (_i, [_vp, _vp]),
    "sodso_ctx_sync": (_i, [_vp]),
    "sodso_ctx_set_stream_threshold": (_i, [_vp, _i]),
    "sodso_last_error": (C.c_char_p, []),
    "sodso_version": (C.c_char_p, []),
    "sodso_ctx_launch_count": (_i64, [_vp]),
    "sodso_ctx_last_kernel_ms": (_d, [_vp]),
    "sodso_ctx_last_kernel_name": (C.c_char_p, [_vp]),
    "sodso_sc_signature_size": (_i, []),
    "sodso_m2dp_signature_size": (_i, []),
    "sodso_align_pca": (_i, [_vp, _vp, _vp, _i, _vp, _vp]),
    "sodso_stage_points": (_i, [_vp, _vp, _vp, _i, _vp, _vp, _vp, _i64, _d, _i, C.POINTER(_vp)]),
    "sodso_staged_destroy": (None, [_vp]),
    "sodso_staged_num_scans": (_i, [_vp]),
    "sodso_staged_num_points": (_i64, [_vp]),
    "sodso_staged_xyz": (_vp, [_vp]),
    "sodso_staged_inten": (_vp, [_vp]),
    "sodso_staged_scan_off": (_vp, [_vp]),
    "sodso_staged_copy": (_i, [_vp, _vp, _vp, _vp, _vp]),
    "sodso_sc_generate": (_i, [_vp, _vp, _vp, _vp, _i, _d, _vp]),
    "sodso_m2dp_generate": (_i, [_vp, _vp, _vp, _vp, _i, _d, _vp]),
    "sodso_m2dp_signature": (_i, [_vp, _vp, _vp, _vp, _i, _d, _vp]),
    "sodso_sc_match": (_i, [_vp, _vp, _i, _vp, _i, _vp, _vp]),
    "sodso_sc_match_f32": (_i, [_vp, _vp, _i, _vp, _i, _vp, _vp]),
    "sodso_m2dp_match": (_i, [_vp, _vp, _i, _vp, _i, _vp, _vp]),
    "sodso_m2dp_match_f32": (_i, [_vp, _vp, _i, _vp, _i, _vp, _vp]),
    "sodso_fuse_top1": (_i, [_vp, _vp, _vp, _i, _i, _i, _d, _vp, _vp]),
    "sodso_loop_top1": (_i, [_vp, _i, _vp, _i, _vp, _i, _i, _d, _vp, _vp, _vp, _vp]),
    "sodso_sc_scans_to_loops": (_i, [_vp, _vp, _vp, _vp, _i, _d, _i, _d, _vp, _vp, _vp, _vp, _vp]),
    "sodso_delight_signature_size": (_i, []),
    "sodso_delight_generate": (_i, [_vp, _vp, _vp, _vp, _i, _vp]),
    "sodso_delight_match": (_i, [_vp, _vp, _i, _vp, _i, _vp]),
    "sodso_top1_single": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp]),
    "sodso_gt_loops": (_i, [_vp, _vp, _i, _vp, _i, _d, _i, _vp, _vp, C.POINTER(_i)]),
    "sodso_pr_curve": (_i, [_vp, _vp, _vp, _i, _vp, _i, _d, _i, C.POINTER(_d), C.POINTER(_d), C.POINTER(_i), _vp, _vp, _vp]),
    "sodso_db_create": (_i, [_vp, _i, _vp, _i, _i64, C.POINTER(_vp)]),
    "sodso_db_destroy": (None, [_vp]),
    "sodso_db_append": (_i, [_vp, _vp, _i]),
    "sodso_db_reserve": (_i, [_vp, _i]),
    "sodso_db_reload": (_i, [_vp, _vp]),
    "sodso_db_stream_match": (_i, [_vp, _vp, _vp, _vp, _d, _vp, _i]),
    "sodso_db_size": (_i, [_vp]),
    "sodso_db_match": (_i, [_vp, _vp, _i]),
    "sodso_db_partial_stats": (_i, [_vp, _vp]),
    "sodso_db_topk": (_i, [_vp, _vp, _i64, _i64, _i, _d, _i, _vp, _vp, _vp, _vp]),
    "sodso_topk_merge": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "sodso_topk_merge_device": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "sodso_db_get_distances": (_i, [_vp, _vp, _vp]),
    "sodso_comm_unique_id": (_i, [_vp]),
    "sodso_comm_init": (_i, [_vp, _vp, _i, _i]),
    "sodso_comm_finalize": (_i, [_vp]),
    "sodso_comm_nranks": (_i, [_vp]),
    "sodso_comm_rank": (_i, [_vp]),
    "sodso_comm_nccl_version": (_i, []),
    "sodso_comm_exchange": (_i, [_vp]),
    "sodso_db_query_sharded": (_i, [_vp, _vp, _i, _i64, _i, _d, _i, _vp, _vp, _vp, _vp]),
    "sodso_db_finish_sharded": (_i, [_vp, _i64, _i, _d, _i, _vp, _vp, _vp, _vp]),
    "sodso_db_scans_query_sharded": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _d, _i64, _i, _d, _i, _vp, _vp,
                                          _vp, _vp, _vp]),
}

# include/sodso_pr_debug.h: test hooks (cross-check kernels, kernel variant switches)
_DEBUG_PROTOS = {
    "sodso_debug_set_match_algo": (_i, [_vp, _i]),
    "sodso_debug_set_sc_symmetry": (_i, [_vp, _i]),
    "sodso_debug_set_kernel_flags": (_i, [_i, _i, _i]),
    "sodso_debug_phase_profile": (_i, [_i, C.c_void_p]),
    "sodso_debug_set_peer_exchange": (_i, [_i]),
    "sodso_debug_fast_turns": (_i, [_vp, _vp, _vp, _i64, _vp]),
    "sodso_debug_sc_self_items": (_i64, [_i64, _i64, _i64]),
}


def exported_names():
    return sorted(_PROTOS)


def debug_names():
    return sorted(_DEBUG_PROTOS)


def lib():
    """Load libsodso_pr.so; raise if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `make -C so_dso_place_recognition_b200/csrc` "
                "(or __graft_entry__.build()); there is no CPU fallback")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in {**_PROTOS, **_DEBUG_PROTOS}.items():
            f = getattr(L, name)
            f.restype = res
            f.argtypes = args
        _lib = L
    return _lib


class SodsoError(RuntimeError):
    pass


def check(rc: int):
    if rc != 0:
        msg = lib().sodso_last_error()
        raise SodsoError(f"libsodso_pr error {rc}: {msg.decode(errors='replace') if msg else ''}")
