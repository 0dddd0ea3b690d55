"""ORACLE — TEST INFRASTRUCTURE ONLY.

ctypes front-end to oracle/_ref/libsodso_refsrc.so: the reference's OWN descriptor sources (SC.cpp, M2DP.cpp,
DELIGHT.cpp, utils/pts_align.h, utils/pts_preprocess.h, PosesPts.h) compiled unchanged from /root/reference against
oracle/eigen_shim (recipe: `make -C oracle refsrc`).  It exists to pin the restatement in sodso_oracle.cpp against the
reference's source lines; Eigen's own arithmetic (the two decompositions, product summation order) is NOT pinned by it
-- see oracle/eigen_shim/Eigen/Core.  Only tests/ and tests/golden/make_*.py may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libsodso_refsrc.so")
REFERENCE_ROOT = "/root/reference"
_lib = None


def available() -> bool:
    """True when the library is built or can be built (the reference sources are present)."""
    return os.path.exists(_SO) or os.path.isdir(os.path.join(REFERENCE_ROOT, "place_recognition"))


def build() -> str:
    if os.path.isdir(os.path.join(REFERENCE_ROOT, "place_recognition")):
        subprocess.check_call(["make", "-C", _HERE, "-s", "refsrc"], env={**os.environ, "CXX": "g++"},
                              stdout=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        _lib.ref_stage_run.restype = C.c_void_p
        _lib.ref_stage_num_points.restype = C.c_int64
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _scans(xyz, inten, off):
    xyz = np.ascontiguousarray(xyz, dtype=np.float64).reshape(-1, 3)
    inten = np.ascontiguousarray(inten, dtype=np.float32).reshape(-1)
    off = np.ascontiguousarray(off, dtype=np.int64)
    return xyz, inten, off, len(off) - 1


def sc_generate(xyz, inten, off, max_rho):
    """SC::getSignature per scan, rows [structure | intensity] as test_sc.cpp:52-55 -> nscan x 2400."""
    xyz, inten, off, n = _scans(xyz, inten, off)
    hist = np.zeros((n, 2400))
    lib().ref_sc_generate(_p(xyz, C.c_double), _p(inten, C.c_float), _p(off, C.c_int64), C.c_int(n),
                          C.c_double(max_rho), _p(hist, C.c_double))
    return hist


def m2dp_generate(xyz, inten, off, max_rho):
    """align_points_PCA + the four direction variants of test_m2dp.cpp:44-66 -> 4 nscan x 384."""
    xyz, inten, off, n = _scans(xyz, inten, off)
    hist = np.zeros((4 * n, 384))
    lib().ref_m2dp_generate(_p(xyz, C.c_double), _p(inten, C.c_float), _p(off, C.c_int64), C.c_int(n),
                            C.c_double(max_rho), _p(hist, C.c_double))
    return hist


def delight_generate(xyz, inten, off):
    """DELIGHT::getSignature per scan (test_delight.cpp:42-56) -> 16 nscan x 256."""
    xyz, inten, off, n = _scans(xyz, inten, off)
    hist = np.zeros((16 * n, 256))
    lib().ref_delight_generate(_p(xyz, C.c_double), _p(inten, C.c_float), _p(off, C.c_int64), C.c_int(n),
                               _p(hist, C.c_double))
    return hist


def stage_run(poses_file, pts_file, incoming_id_file, lidar_range, polar_filter):
    """pts_preprocess (pts_preprocess.h:169-233): returns (off, xyz, inten); writes incoming_id_file."""
    L = lib()
    # the reference prints a progress line per frame on stdout (pts_preprocess.h:160-163): silence it for the call
    import sys
    sys.stdout.flush()
    keep = os.dup(1)
    null = os.open(os.devnull, os.O_WRONLY)
    os.dup2(null, 1)
    try:
        h = C.c_void_p(L.ref_stage_run(str(poses_file).encode(), str(pts_file).encode(),
                                       str(incoming_id_file).encode(), C.c_double(lidar_range),
                                       C.c_int(int(polar_filter))))
        C.CDLL(None).fflush(None)
    finally:
        os.dup2(keep, 1)
        os.close(keep)
        os.close(null)
    try:
        n = L.ref_stage_num_scans(h)
        tot = L.ref_stage_num_points(h)
        off = np.zeros(n + 1, np.int64)
        xyz = np.zeros((max(tot, 1), 3))
        inten = np.zeros(max(tot, 1), np.float32)
        L.ref_stage_copy(h, _p(off, C.c_int64), _p(xyz, C.c_double), _p(inten, C.c_float))
    finally:
        L.ref_stage_free(h)
    return off, xyz[:tot], inten[:tot]
