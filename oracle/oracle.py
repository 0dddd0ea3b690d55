"""ORACLE — TEST INFRASTRUCTURE ONLY (see sodso_oracle.cpp header).

ctypes front-end to oracle/_build/libsodso_oracle.so plus a numpy/BLAS restatement of
the matcher (MATLAB's `*` is a multithreaded BLAS dgemm, so numpy+OpenBLAS is the fair
stand-in for timing).  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module.

parity unpinned for Eigen's arithmetic (decomposition signs, summation order): no golden
signatures exist in the reference.  Pinned: incoming_id_file.txt
(tests/test_oracle_golden.py) and, bit for bit, every statement of the reference's own
descriptor sources compiled unchanged against oracle/eigen_shim (oracle/refsrc.py,
tests/test_refsrc_pin.py).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import warnings

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libsodso_oracle.so")
_lib = None

SC_SIZE = 1200
M2DP_SIZE = 192


def build(force: bool = False) -> str:
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(
        os.path.join(_HERE, "sodso_oracle.cpp")
    ):
        subprocess.check_call(["make", "-C", _HERE, "-s"], env={**os.environ, "CXX": "g++"})
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        try:
            _lib = C.CDLL(_SO)
        except OSError:
            build(force=True)
            _lib = C.CDLL(_SO)
        _lib.orc_stage_run.restype = C.c_void_p
        _lib.orc_stage_arrays.restype = C.c_void_p
        _lib.orc_stage_num_points.restype = C.c_int64
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def sym_eig3(a):
    a = _f64(a).reshape(9)
    w = np.zeros(3)
    v = np.zeros(9)
    lib().orc_sym_eig3(_p(a, C.c_double), _p(w, C.c_double), _p(v, C.c_double))
    return w, v.reshape(3, 3)


def align_pca(xyz):
    """pts_align.h:7-46 -> (aligned n x 3, evec 3x3 (columns), mean 3)."""
    xyz = _f64(xyz).reshape(-1, 3)
    n = xyz.shape[0]
    out = np.zeros_like(xyz)
    ev = np.zeros(9)
    mean = np.zeros(3)
    lib().orc_align_pca(_p(xyz, C.c_double), C.c_int(n), _p(out, C.c_double), _p(ev, C.c_double), _p(mean, C.c_double))
    return out, ev.reshape(3, 3), mean


def sc_signature(xyz, inten, max_rho=45.0):
    """SC::getSignature (SC.cpp:12-76) -> (structure[1200], intensity[1200])."""
    xyz = _f64(xyz).reshape(-1, 3)
    inten = _f32(inten)
    s = np.zeros(SC_SIZE)
    i = np.zeros(SC_SIZE)
    lib().orc_sc_signature(_p(xyz, C.c_double), _p(inten, C.c_float), C.c_int(xyz.shape[0]), C.c_double(max_rho),
                           _p(s, C.c_double), _p(i, C.c_double))
    return s, i


def sc_generate(xyz, inten, off, max_rho=45.0, nthreads=1):
    """test_sc.cpp:36-57 -> history_sc (nscan x 2400)."""
    xyz = _f64(xyz).reshape(-1, 3)
    inten = _f32(inten)
    off = np.ascontiguousarray(off, dtype=np.int64)
    ns = off.shape[0] - 1
    hist = np.zeros((ns, 2 * SC_SIZE))
    lib().orc_sc_generate(_p(xyz, C.c_double), _p(inten, C.c_float), _p(off, C.c_int64), C.c_int(ns),
                          C.c_double(max_rho), _p(hist, C.c_double), C.c_int(nthreads))
    return hist


def sc_generate_flip(xyz, inten, off, pca_flip, max_rho=45.0, nthreads=1):
    """sign experiment: sc_generate with eigenvector k of scan s negated when bit k of pca_flip[s] is set"""
    xyz = _f64(xyz).reshape(-1, 3)
    inten = _f32(inten)
    off = np.ascontiguousarray(off, dtype=np.int64)
    ns = off.shape[0] - 1
    flip = np.ascontiguousarray(pca_flip, dtype=np.int32)
    assert flip.shape == (ns,)
    hist = np.zeros((ns, 2 * SC_SIZE))
    lib().orc_sc_generate_flip(_p(xyz, C.c_double), _p(inten, C.c_float), _p(off, C.c_int64), C.c_int(ns),
                               C.c_double(max_rho), _p(flip, C.c_int32), _p(hist, C.c_double), C.c_int(nthreads))
    return hist


def m2dp_generate_flip(xyz, inten, off, pca_flip, svd_flip, max_rho=45.0, nthreads=1):
    """sign experiment: m2dp_generate with per-scan PCA eigenvector flips (bits 0..2) and per-scan flips of the
    dominant singular pair (bit 0: count matrix, bit 1: intensity matrix)"""
    xyz = _f64(xyz).reshape(-1, 3)
    inten = _f32(inten)
    off = np.ascontiguousarray(off, dtype=np.int64)
    ns = off.shape[0] - 1
    pf = np.ascontiguousarray(pca_flip, dtype=np.int32)
    sf = np.ascontiguousarray(svd_flip, dtype=np.int32)
    assert pf.shape == (ns,) and sf.shape == (ns,)
    hist = np.zeros((4 * ns, 2 * M2DP_SIZE))
    lib().orc_m2dp_generate_flip(_p(xyz, C.c_double), _p(inten, C.c_float), _p(off, C.c_int64), C.c_int(ns),
                                 C.c_double(max_rho), _p(pf, C.c_int32), _p(sf, C.c_int32), _p(hist, C.c_double),
                                 C.c_int(nthreads))
    return hist


def m2dp_tables():
    x = np.zeros(3 * 64)
    y = np.zeros(3 * 64)
    lib().orc_m2dp_tables(_p(x, C.c_double), _p(y, C.c_double))
    return x.reshape(64, 3), y.reshape(64, 3)


def svd_dominant(A):
    A = _f64(A)
    r, c = A.shape
    u = np.zeros(r)
    v = np.zeros(c)
    s = C.c_double(0)
    lib().orc_svd_dominant(_p(A, C.c_double), C.c_int(r), C.c_int(c), _p(u, C.c_double), _p(v, C.c_double), C.byref(s))
    return u, v, s.value


def m2dp_signature(pts_aligned, inten, max_rho=45.0, want_hist=False):
    """M2DP::getSignature (M2DP.cpp:38-109) on already-aligned points."""
    pts = _f64(pts_aligned).reshape(-1, 3)
    inten = _f32(inten)
    c = np.zeros(M2DP_SIZE)
    i = np.zeros(M2DP_SIZE)
    Ac = np.zeros((64, 128)) if want_hist else None
    Ai = np.zeros((64, 128)) if want_hist else None
    lib().orc_m2dp_signature(_p(pts, C.c_double), _p(inten, C.c_float), C.c_int(pts.shape[0]), C.c_double(max_rho),
                             _p(c, C.c_double), _p(i, C.c_double), _p(Ac, C.c_double), _p(Ai, C.c_double))
    return (c, i, Ac, Ai) if want_hist else (c, i)


def m2dp_generate(xyz, inten, off, max_rho=45.0, nthreads=1):
    """test_m2dp.cpp:37-67 -> history_m2dp (4*nscan x 384)."""
    xyz = _f64(xyz).reshape(-1, 3)
    inten = _f32(inten)
    off = np.ascontiguousarray(off, dtype=np.int64)
    ns = off.shape[0] - 1
    hist = np.zeros((4 * ns, 2 * M2DP_SIZE))
    lib().orc_m2dp_generate(_p(xyz, C.c_double), _p(inten, C.c_float), _p(off, C.c_int64), C.c_int(ns),
                            C.c_double(max_rho), _p(hist, C.c_double), C.c_int(nthreads))
    return hist


def sc_match(hist1, hist2, nthreads=1):
    """processSC.m:1-45 -> (d_p, d_i), each m x n."""
    h1 = _f64(hist1)
    h2 = _f64(hist2)
    m, n = h1.shape[0], h2.shape[0]
    dp = np.zeros((m, n))
    di = np.zeros((m, n))
    lib().orc_sc_match(_p(h1, C.c_double), C.c_int(m), _p(h2, C.c_double), C.c_int(n), _p(dp, C.c_double),
                       _p(di, C.c_double), C.c_int(nthreads))
    return dp, di


def m2dp_match(hist1, hist2, nthreads=1):
    """processM2DP.m:1-22 -> (d_p, d_i), each (rows1/4) x (rows2/4)."""
    h1 = _f64(hist1)
    h2 = _f64(hist2)
    m, n = h1.shape[0] // 4, h2.shape[0] // 4
    dp = np.zeros((m, n))
    di = np.zeros((m, n))
    lib().orc_m2dp_match(_p(h1, C.c_double), C.c_int(m), _p(h2, C.c_double), C.c_int(n), _p(dp, C.c_double),
                         _p(di, C.c_double), C.c_int(nthreads))
    return dp, di


def fuse_top1(d_p, d_i, mask_width, p_weight=2.0, want_fused=False):
    """run_test.m:38-57 -> (idx 0-based int32[m], score[m][, fused m x n])."""
    d_p = _f64(d_p)
    d_i = _f64(d_i)
    m, n = d_p.shape
    idx = np.zeros(m, dtype=np.int32)
    score = np.zeros(m)
    fused = np.zeros((m, n)) if want_fused else None
    lib().orc_fuse_top1(_p(d_p, C.c_double), _p(d_i, C.c_double), C.c_int(m), C.c_int(n), C.c_int(mask_width),
                        C.c_double(p_weight), _p(idx, C.c_int32), _p(score, C.c_double), _p(fused, C.c_double))
    return (idx, score, fused) if want_fused else (idx, score)


def stage(poses_file, pts_file, lidar_range=45.0, polar_filter=False):
    """pts_preprocess (pts_preprocess.h:169-232) -> dict(ids, off, xyz, inten, n_poses)."""
    L = lib()
    h = C.c_void_p(L.orc_stage_run(poses_file.encode(), pts_file.encode(), C.c_double(lidar_range),
                                   C.c_int(1 if polar_filter else 0)))
    return _stage_result(L, h)


def _stage_result(L, h):
    try:
        ns = L.orc_stage_num_scans(h)
        npts = L.orc_stage_num_points(h)
        ids = np.zeros(ns, dtype=np.int32)
        off = np.zeros(ns + 1, dtype=np.int64)
        xyz = np.zeros((npts, 3))
        inten = np.zeros(npts, dtype=np.float32)
        L.orc_stage_copy(h, _p(ids, C.c_int), _p(off, C.c_int64), _p(xyz, C.c_double), _p(inten, C.c_float))
        return dict(ids=ids, off=off, xyz=xyz, inten=inten, n_poses=L.orc_stage_num_poses(h))
    finally:
        L.orc_stage_free(h)


def stage_arrays(pose_id, w2c, pt_id, pt_xyz, pt_inten, lidar_range=45.0, polar_filter=False):
    """pts_preprocess on in-memory records (what read_poses_pts parses, pts_preprocess.h:17-49)."""
    L = lib()
    pose_id = np.ascontiguousarray(pose_id, dtype=np.int32)
    w2c = np.ascontiguousarray(w2c, dtype=np.float64).reshape(-1, 12)
    pt_id = np.ascontiguousarray(pt_id, dtype=np.int32)
    pt_xyz = np.ascontiguousarray(pt_xyz, dtype=np.float64).reshape(-1, 3)
    pt_inten = np.ascontiguousarray(pt_inten, dtype=np.float32)
    h = C.c_void_p(L.orc_stage_arrays(_p(pose_id, C.c_int), _p(w2c, C.c_double), C.c_int(len(pose_id)),
                                      _p(pt_id, C.c_int), _p(pt_xyz, C.c_double), _p(pt_inten, C.c_float),
                                      C.c_int64(len(pt_id)), C.c_double(lidar_range), C.c_int(1 if polar_filter else 0)))
    return _stage_result(L, h)


# ---------------------------------------------------------------------------------------
# numpy / BLAS restatement of the matcher (used for timing the CPU baseline the way
# MATLAB would run it: one dgemm per query, processSC.m:22-33).
# ---------------------------------------------------------------------------------------
def _perm_index():
    """120 x 1200 gather index: row 2(k-1) forward shift k, row 2(k-1)+1 reversed (processSC.m:24-28,37-45)."""
    idx = np.zeros((120, SC_SIZE), dtype=np.int64)
    r = np.arange(20)
    for k in range(1, 61):
        c = np.arange(60)
        sf = (k - 1 + c) % 60
        sr = (k - 1 - c) % 60
        idx[2 * k - 2] = (sf[:, None] * 20 + r[None, :]).reshape(-1)
        idx[2 * k - 1] = (sr[:, None] * 20 + r[None, :]).reshape(-1)
    return idx


_PERM = None


def sc_match_numpy(hist1, hist2):
    global _PERM
    if _PERM is None:
        _PERM = _perm_index()
    h1 = _f64(hist1)
    h2 = _f64(hist2)
    out = []
    for ch in range(2):
        a = h1[:, ch * SC_SIZE:(ch + 1) * SC_SIZE]
        b = h2[:, ch * SC_SIZE:(ch + 1) * SC_SIZE]
        with np.errstate(invalid="ignore", divide="ignore"):
            a = a / np.linalg.norm(a, axis=1, keepdims=True)
            b = b / np.linalg.norm(b, axis=1, keepdims=True)
        bt = np.ascontiguousarray(b.T)
        res = np.empty((a.shape[0], b.shape[0]))
        for i in range(a.shape[0]):
            sig = a[i][_PERM]                       # 120 x 1200
            diff = (1.0 - sig @ bt) / 2.0           # processSC.m:30
            res[i] = np.fmin.reduce(diff, axis=0)   # min ignoring NaN (:31)
        out.append(res)
    return out[0], out[1]


def fuse_top1_numpy(d_p, d_i, mask_width, p_weight=2.0):
    def z(x):   # MATLAB normalize(x, 2): z-score per row, mean / std (N-1) with 'omitnan'
        with np.errstate(invalid="ignore", divide="ignore"), warnings.catch_warnings():
            warnings.simplefilter("ignore", RuntimeWarning)
            return (x - np.nanmean(x, axis=1, keepdims=True)) / np.nanstd(x, axis=1, ddof=1, keepdims=True)

    f = p_weight * z(d_p) + z(d_i)
    f = np.where(np.isnan(f), np.inf, f)          # min skips NaN (run_test.m:57); an all-NaN row -> index 0
    m, n = f.shape
    ii = np.arange(m)[:, None]
    jj = np.arange(n)[None, :]
    f = np.where(np.abs(ii - jj) < mask_width, np.inf, f)
    idx = np.argmin(f, axis=1).astype(np.int32)
    return idx, f[np.arange(m), idx]


# ---------------------------------------------------------------------------------------
# evaluation: run_test.m:2-22 (ground-truth loop set) and :56-85 (precision / recall, AUC, top recall)
# ---------------------------------------------------------------------------------------
def gt_loops(gt1, gt2, loop_diff, mask_width):
    """run_test.m:3-22 -> (lp_gt as 0-based (i, j) rows, total_lp = MATLAB length(lp_gt))."""
    gt1 = np.asarray(gt1, dtype=np.float64)
    gt2 = np.asarray(gt2, dtype=np.float64)
    m, n = gt1.shape[0], gt2.shape[0]
    lp = []
    jj = np.arange(n)
    for i in range(m):
        ok = np.abs(i - jj) >= mask_width                                  # :8-10
        if not ok.any():
            continue
        d = gt1[i] - gt2                                                   # :11
        d2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]   # :12
        d2 = np.where(ok & ~np.isnan(d2), d2, np.inf)
        j = int(np.argmin(d2))                                             # strict '>' at :13 -> first minimum
        if d2[j] < loop_diff * loop_diff:                                  # :19
            lp.append((i, j))
    lp = np.array(lp, dtype=np.int64).reshape(-1, 2)
    L = lp.shape[0]
    total_lp = 0 if L == 0 else max(L, 2)                                  # length() of an L x 2 matrix, :22
    return lp, total_lp


def pr_eval(diff_v, diff_idx, gt1, gt2, total_lp, loop_diff):
    """run_test.m:56-85 from the per-query decision (0-based diff_idx) -> dict(AUC, top_recall, top_count, rank,
    precision, recall)."""
    diff_v = np.asarray(diff_v, dtype=np.float64)
    m = diff_v.shape[0]
    key = np.where(np.isnan(diff_v), np.inf, diff_v)
    rank = np.lexsort((np.arange(m), np.isnan(diff_v), key))               # sort: ascending, stable, NaN last (:58)
    tp = fp = 0
    precision, recall = np.zeros(m), np.zeros(m)
    top_recall, top_count = 0.0, 0
    with np.errstate(all="ignore"):
        for i in range(m):
            a, b = rank[i], diff_idx[rank[i]]
            d = gt1[a] - gt2[b]
            if (d[0] * d[0] + d[1] * d[1]) + d[2] * d[2] < loop_diff * loop_diff:   # :68-70
                tp += 1
            else:
                fp += 1
            precision[i] = tp / (tp + fp)
            recall[i] = np.float64(tp) / np.float64(total_lp)
            if precision[i] == 1:                                          # :78-81
                top_count, top_recall = i + 1, recall[i]
        auc = float(np.sum((recall[1:] - recall[:-1]) * (precision[1:] + precision[:-1]) * 0.5)) if m > 1 else 0.0  # :84
    return dict(AUC=auc, top_recall=float(top_recall), top_count=top_count, rank=rank, precision=precision, recall=recall)


# ---------------------------------------------------------------------------------------
# DELIGHT
# ---------------------------------------------------------------------------------------
def delight_generate(xyz, inten, off, nthreads=1):
    """test_delight.cpp:38-56 -> history_delight (16*nscan x 256)."""
    xyz = _f64(xyz).reshape(-1, 3)
    inten = _f32(inten)
    off = np.ascontiguousarray(off, dtype=np.int64)
    ns = off.shape[0] - 1
    hist = np.zeros((16 * ns, 256))
    lib().orc_delight_generate(_p(xyz, C.c_double), _p(inten, C.c_float), _p(off, C.c_int64), C.c_int(ns),
                               _p(hist, C.c_double), C.c_int(nthreads))
    return hist


def delight_match(hist1, hist2, nthreads=1):
    """dist = processDELIGHT(hist1, hist2) (processDELIGHT.m:1-38)."""
    h1, h2 = _f64(hist1), _f64(hist2)
    m, n = h1.shape[0] // 16, h2.shape[0] // 16
    d = np.zeros((m, n))
    lib().orc_delight_match(_p(h1, C.c_double), C.c_int(m), _p(h2, C.c_double), C.c_int(n), _p(d, C.c_double),
                            C.c_int(nthreads))
    return d


def top1_single(d, mask_width):
    """run_test.m:47-57 on one distance matrix -> (idx 0-based, score)."""
    d = _f64(d)
    m, n = d.shape
    idx = np.zeros(m, dtype=np.int32)
    score = np.zeros(m)
    lib().orc_top1_single(_p(d, C.c_double), C.c_int(m), C.c_int(n), C.c_int(mask_width), _p(idx, C.c_int32),
                          _p(score, C.c_double))
    return idx, score
