"""Host-side mirror of the reference interface for the descriptor hot path, over the C ABI.

Names, argument meaning and results follow the reference:
  SC / M2DP classes      <- place_recognition/generate_signatures/src/SC/SC.h:10-23, M2DP/M2DP.h:12-30
  align_points_PCA       <- .../src/utils/pts_align.h:7-9
  sc_generate / m2dp_generate (batch drivers) <- SC/test_sc.cpp:36-57, M2DP/test_m2dp.cpp:37-67
  processSC / processM2DP <- place_recognition/match_signatures/processSC.m:1, processM2DP.m:1
  run_test (lines 25-57 only: match, fuse, mask, argmin) <- match_signatures/run_test.m

Arrays may be numpy arrays (host) or torch CUDA tensors (device; outputs are then torch CUDA
tensors too).  Every call goes to libsodso_pr.so; nothing here computes on the CPU.
"""
from __future__ import annotations

import ctypes as C
import threading

import numpy as np

from . import _native as N
from ._native import SC_SIZE, M2DP_SIZE, SODSO_TYPE_SC, SODSO_TYPE_M2DP, SODSO_ALGO_TC, SODSO_ALGO_SIMT

_ctx_lock = threading.Lock()
_ctxs: dict[int, "Context"] = {}


def _is_torch(x):
    return type(x).__module__.startswith("torch")


def _ptr(a):
    if a is None:
        return None
    if _is_torch(a):
        return C.c_void_p(a.data_ptr())
    return C.c_void_p(a.ctypes.data)


def _prep(a, dtype_np, dtype_name):
    """contiguous array of the right dtype, numpy or torch-cuda"""
    if a is None:
        return None
    if _is_torch(a):
        import torch

        dt = getattr(torch, dtype_name)
        if a.dtype != dt:
            a = a.to(dt)
        return a.contiguous()
    return np.ascontiguousarray(a, dtype=dtype_np)


def _empty_like_kind(ref, shape, dtype_np, dtype_name):
    if _is_torch(ref):
        import torch

        return torch.empty(shape, dtype=getattr(torch, dtype_name), device=ref.device)
    return np.empty(shape, dtype=dtype_np)


class Context:
    """One per GPU (sodso_ctx)."""

    def __init__(self, device: int = 0):
        self._h = C.c_void_p()
        N.check(N.lib().sodso_ctx_create(int(device), C.byref(self._h)))
        self.device = int(device)

    def close(self):
        if self._h:
            N.lib().sodso_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    def set_match_algo(self, algo: int):
        """test hook (include/sodso_pr_debug.h): SODSO_ALGO_SIMT selects the fp32 cross-check kernels"""
        N.check(N.lib().sodso_debug_set_match_algo(self._h, int(algo)))

    def set_sc_symmetry(self, on: bool):
        """test hook (include/sodso_pr_debug.h): off = a self-match computes every pair like distinct operands"""
        N.check(N.lib().sodso_debug_set_sc_symmetry(self._h, 1 if on else 0))

    def set_stream_threshold(self, min_scans: int):
        N.check(N.lib().sodso_ctx_set_stream_threshold(self._h, int(min_scans)))

    def sync(self):
        N.check(N.lib().sodso_ctx_sync(self._h))

    # ---- multi-GPU: NCCL communicator inside the library (include/sodso_pr.h, sodso_comm_*) ----
    @staticmethod
    def comm_unique_id() -> bytes:
        buf = C.create_string_buffer(N.COMM_ID_BYTES)
        N.check(N.lib().sodso_comm_unique_id(buf))
        return buf.raw

    def comm_init(self, unique_id: bytes | None, nranks: int, rank: int):
        """collective: every rank calls it with rank 0's id (sodso_comm_init -> ncclCommInitRank)"""
        buf = C.create_string_buffer(unique_id, N.COMM_ID_BYTES) if unique_id else None
        N.check(N.lib().sodso_comm_init(self._h, buf, int(nranks), int(rank)))

    def comm_finalize(self):
        N.check(N.lib().sodso_comm_finalize(self._h))

    @property
    def comm_exchange(self) -> str:
        """transport of the per-batch exchange of sharded queries: 'peer-memory' (NVLink writes from the kernels) or 'nccl'"""
        return "peer-memory" if N.lib().sodso_comm_exchange(self._h) else "nccl"

    @property
    def comm_nranks(self) -> int:
        return int(N.lib().sodso_comm_nranks(self._h))

    @property
    def comm_rank(self) -> int:
        return int(N.lib().sodso_comm_rank(self._h))

    def set_stream(self, cuda_stream_ptr):
        N.check(N.lib().sodso_ctx_set_stream(self._h, C.c_void_p(cuda_stream_ptr or 0)))

    @property
    def stream(self):
        return N.lib().sodso_ctx_stream(self._h)

    @property
    def launch_count(self) -> int:
        return int(N.lib().sodso_ctx_launch_count(self._h))

    @property
    def last_kernel_ms(self) -> float:
        return float(N.lib().sodso_ctx_last_kernel_ms(self._h))

    @property
    def last_kernel_name(self) -> str:
        return N.lib().sodso_ctx_last_kernel_name(self._h).decode()


def default_context(device: int | None = None) -> Context:
    if device is None:
        device = 0
    with _ctx_lock:
        c = _ctxs.get(device)
        if c is None:
            c = Context(device)
            _ctxs[device] = c
        return c


def _ctx_for(*arrays, ctx=None):
    """The context of a call.  If any argument is a torch CUDA tensor, the context's stream is made to wait for
    torch's current stream first, so that tensors produced by earlier torch work (copies, NCCL collectives)
    are complete before the library reads them; every library call that writes a device output synchronises
    its stream before returning, which orders the other direction."""
    c = ctx
    for a in arrays:
        if a is not None and _is_torch(a) and a.is_cuda:
            if c is None:
                c = default_context(a.device.index or 0)
            import torch

            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(a.device))
            torch.cuda.ExternalStream(c.stream or 0, device=a.device).wait_event(ev)
            break
    return c if c is not None else default_context(0)


# ----------------------------------------------------------------------------------------------
# signature files: the reference's text matrices (test_sc.cpp:63-66 / test_kitti.m:26) and the binary container of
# so_dso_place_recognition_b200/host/sodso_host.hpp
# ----------------------------------------------------------------------------------------------
_HIST_MAGIC = b"SODSOHST"


def save_history(path, hist):
    """.bin -> binary container (64-byte header + row-major little-endian f64), anything else -> text, 6 significant
    digits like Eigen's operator<< (column alignment is cosmetic and not reproduced here)."""
    hist = np.ascontiguousarray(hist.cpu().numpy() if _is_torch(hist) else hist, dtype=np.float64)
    if str(path).endswith(".bin"):
        with open(path, "wb") as f:
            f.write(_HIST_MAGIC + np.array([1, 0], dtype="<u4").tobytes() + np.array(hist.shape, dtype="<u8").tobytes() + bytes(32))
            f.write(hist.astype("<f8").tobytes())
    else:
        np.savetxt(path, hist, fmt="%.6g")


def load_history(path, mmap=True):
    """text matrix (numpy.loadtxt, like MATLAB load) or binary container (memory-mapped by default)."""
    with open(path, "rb") as f:
        head = f.read(64)
    if head[:8] != _HIST_MAGIC:
        return np.loadtxt(path, ndmin=2)
    version, dtype = np.frombuffer(head[8:16], dtype="<u4")
    rows, cols = (int(v) for v in np.frombuffer(head[16:32], dtype="<u8"))
    if version != 1 or dtype != 0:
        raise ValueError(f"{path}: unsupported signature container version {version} / dtype {dtype}")
    if mmap:
        return np.memmap(path, dtype="<f8", mode="r", offset=64, shape=(rows, cols))
    return np.fromfile(path, dtype="<f8", offset=64, count=rows * cols).reshape(rows, cols)


# ----------------------------------------------------------------------------------------------
# point staging
# ----------------------------------------------------------------------------------------------
def pts_preprocess(pose_id, w2c, pt_id, pt_xyz, pt_inten, lidar_range=45.0, polar_filter=False, device=False, ctx=None):
    """pts_preprocess(poses_file, pts_file, id_file, lidarRange, out_vec, polar_filter) (pts_preprocess.h:169-232)
    on the parsed records -> dict(ids, off, xyz, inten).  Points of a scan are ordered by voxel index.
    device=True returns xyz / inten / off as torch CUDA tensors (ready for sc_generate / m2dp_generate)."""
    pose_id = np.ascontiguousarray(pose_id, dtype=np.int32)
    w2c = np.ascontiguousarray(w2c, dtype=np.float64).reshape(-1, 12)
    pt_id = np.ascontiguousarray(pt_id, dtype=np.int32)
    pt_xyz = _prep(pt_xyz, np.float64, "float64")
    pt_inten = _prep(pt_inten, np.float32, "float32")
    c = _ctx_for(pt_xyz, ctx=ctx)
    h = C.c_void_p()
    L = N.lib()
    N.check(L.sodso_stage_points(c.handle, _ptr(pose_id), _ptr(w2c), len(pose_id), _ptr(pt_id), _ptr(pt_xyz),
                                 _ptr(pt_inten), len(pt_id), float(lidar_range), 1 if polar_filter else 0, C.byref(h)))
    try:
        ns, npts = L.sodso_staged_num_scans(h), L.sodso_staged_num_points(h)
        ids = np.zeros(ns, dtype=np.int32)
        if device:
            import torch

            dev = torch.device("cuda", c.device)
            off = torch.empty(ns + 1, dtype=torch.int64, device=dev)
            xyz = torch.empty((npts, 3), dtype=torch.float64, device=dev)
            inten = torch.empty(npts, dtype=torch.float32, device=dev)
        else:
            off = np.zeros(ns + 1, dtype=np.int64)
            xyz = np.zeros((npts, 3))
            inten = np.zeros(npts, dtype=np.float32)
        N.check(L.sodso_staged_copy(h, _ptr(ids), _ptr(off), _ptr(xyz), _ptr(inten)))
    finally:
        L.sodso_staged_destroy(h)
    return dict(ids=ids, off=off, xyz=xyz, inten=inten)


# ----------------------------------------------------------------------------------------------
# generation
# ----------------------------------------------------------------------------------------------
def align_points_PCA(xyz, off=None, want_evec=False, ctx=None):
    """pts_align.h:7-46.  xyz: (n,3) [+ off for a batch].  -> aligned (n,3) [, evec (nscan,3,3)]"""
    xyz = _prep(xyz, np.float64, "float64")
    n = xyz.shape[0]
    if off is None:
        off = np.array([0, n], dtype=np.int64)
    off = _prep(off, np.int64, "int64")
    nscan = off.shape[0] - 1
    c = _ctx_for(xyz, ctx=ctx)
    out = _empty_like_kind(xyz, (n, 3), np.float64, "float64")
    ev = _empty_like_kind(xyz, (nscan, 3, 3), np.float64, "float64") if want_evec else None
    N.check(N.lib().sodso_align_pca(c.handle, _ptr(xyz), _ptr(off), nscan, _ptr(out), _ptr(ev)))
    return (out, ev) if want_evec else out


def sc_generate(xyz, inten, off, max_rho=45.0, ctx=None):
    """test_sc.cpp:36-57: history_sc (nscan x 2400) = rows [structure(1200), intensity(1200)]."""
    xyz = _prep(xyz, np.float64, "float64")
    inten = _prep(inten, np.float32, "float32")
    off = _prep(off, np.int64, "int64")
    nscan = off.shape[0] - 1
    c = _ctx_for(xyz, ctx=ctx)
    hist = _empty_like_kind(xyz, (nscan, 2 * SC_SIZE), np.float64, "float64")
    N.check(N.lib().sodso_sc_generate(c.handle, _ptr(xyz), _ptr(inten), _ptr(off), nscan, float(max_rho), _ptr(hist)))
    return hist


def m2dp_generate(xyz, inten, off, max_rho=45.0, ctx=None):
    """test_m2dp.cpp:37-67: history_m2dp (4*nscan x 384)."""
    xyz = _prep(xyz, np.float64, "float64")
    inten = _prep(inten, np.float32, "float32")
    off = _prep(off, np.int64, "int64")
    nscan = off.shape[0] - 1
    c = _ctx_for(xyz, ctx=ctx)
    hist = _empty_like_kind(xyz, (4 * nscan, 2 * M2DP_SIZE), np.float64, "float64")
    N.check(N.lib().sodso_m2dp_generate(c.handle, _ptr(xyz), _ptr(inten), _ptr(off), nscan, float(max_rho), _ptr(hist)))
    return hist


class SC:
    """SC.h:10-23.  getSignature takes the points of ONE scan (n x 3 doubles, n floats)."""

    def __init__(self, max_rho: float, ctx=None):
        self.max_rho = float(max_rho)
        self._ctx = ctx

    def getSignatureSize(self) -> int:
        return int(N.lib().sodso_sc_signature_size())

    def getSignature(self, pts, inten):
        pts = _prep(pts, np.float64, "float64")
        n = pts.shape[0]
        h = sc_generate(pts, inten, np.array([0, n], dtype=np.int64), self.max_rho, ctx=self._ctx)
        return h[0, :SC_SIZE], h[0, SC_SIZE:]


class M2DP:
    """M2DP.h:12-30.  getSignature expects ALREADY aligned (+ sign-flipped) points, like the class."""

    def __init__(self, max_rho: float, ctx=None):
        self.max_rho = float(max_rho)
        self._ctx = ctx

    def getSignatureSize(self) -> int:
        return int(N.lib().sodso_m2dp_signature_size())

    def getSignature(self, pts_aligned, inten):
        pts = _prep(pts_aligned, np.float64, "float64")
        inten = _prep(inten, np.float32, "float32")
        n = pts.shape[0]
        off = np.array([0, n], dtype=np.int64)
        c = _ctx_for(pts, ctx=self._ctx)
        sig = _empty_like_kind(pts, (1, 2 * M2DP_SIZE), np.float64, "float64")
        N.check(N.lib().sodso_m2dp_signature(c.handle, _ptr(pts), _ptr(inten), _ptr(off), 1, self.max_rho, _ptr(sig)))
        return sig[0, :M2DP_SIZE], sig[0, M2DP_SIZE:]


# ----------------------------------------------------------------------------------------------
# matching
# ----------------------------------------------------------------------------------------------
def _match(fn64, fn32, hist1, hist2, rows_per_sig, f32, ctx):
    hist1 = _prep(hist1, np.float64, "float64")
    hist2 = _prep(hist2, np.float64, "float64")
    m, n = hist1.shape[0] // rows_per_sig, hist2.shape[0] // rows_per_sig
    c = _ctx_for(hist1, hist2, ctx=ctx)
    ref = hist1 if _is_torch(hist1) else hist2
    if f32:
        dp = _empty_like_kind(ref, (m, n), np.float32, "float32")
        di = _empty_like_kind(ref, (m, n), np.float32, "float32")
        N.check(fn32(c.handle, _ptr(hist1), m, _ptr(hist2), n, _ptr(dp), _ptr(di)))
    else:
        dp = _empty_like_kind(ref, (m, n), np.float64, "float64")
        di = _empty_like_kind(ref, (m, n), np.float64, "float64")
        N.check(fn64(c.handle, _ptr(hist1), m, _ptr(hist2), n, _ptr(dp), _ptr(di)))
    return dp, di


def processSC(hist1, hist2, f32=False, ctx=None):
    """[diff_m_p, diff_m_i] = processSC(hist1, hist2)  (processSC.m:1-45)."""
    L = N.lib()
    return _match(L.sodso_sc_match, L.sodso_sc_match_f32, hist1, hist2, 1, f32, ctx)


def processM2DP(hist1, hist2, f32=False, ctx=None):
    """[diff_m_p, diff_m_i] = processM2DP(hist1, hist2)  (processM2DP.m:1-22)."""
    L = N.lib()
    return _match(L.sodso_m2dp_match, L.sodso_m2dp_match_f32, hist1, hist2, 4, f32, ctx)


def fuse_top1(d_p, d_i, mask_width, p_weight=2.0, ctx=None):
    """run_test.m:38-57 on given matrices -> (idx 0-based int32[m], score[m])."""
    d_p = _prep(d_p, np.float64, "float64")
    d_i = _prep(d_i, np.float64, "float64")
    m, n = d_p.shape
    c = _ctx_for(d_p, d_i, ctx=ctx)
    idx = _empty_like_kind(d_p, (m,), np.int32, "int32")
    score = _empty_like_kind(d_p, (m,), np.float64, "float64")
    N.check(N.lib().sodso_fuse_top1(c.handle, _ptr(d_p), _ptr(d_i), m, n, int(mask_width), float(p_weight),
                                    _ptr(idx), _ptr(score)))
    return idx, score


def run_test(type, hist1, hist2, mask_width, p_weight=2.0, want_channels=False, ctx=None):
    """The timed + decision part of run_test(type, hist1, hist2, gt1, gt2, loop_diff, mask_width)
    (run_test.m:25-57): -> (diff_idx 0-based int32[m], diff_v[m] [, d_p_at, d_i_at])."""
    if type == "delight":     # run_test.m:31-32: one distance matrix, no fusion
        idx, score = top1_single(processDELIGHT(hist1, hist2, ctx=ctx), mask_width, ctx=ctx)
        return (idx, score, None, None) if want_channels else (idx, score)
    t = {"sc": SODSO_TYPE_SC, "m2dp": SODSO_TYPE_M2DP}[type]
    rows = 1 if t == SODSO_TYPE_SC else 4
    hist1 = _prep(hist1, np.float64, "float64")
    hist2 = _prep(hist2, np.float64, "float64")
    m, n = hist1.shape[0] // rows, hist2.shape[0] // rows
    c = _ctx_for(hist1, hist2, ctx=ctx)
    ref = hist1 if _is_torch(hist1) else hist2
    idx = _empty_like_kind(ref, (m,), np.int32, "int32")
    score = _empty_like_kind(ref, (m,), np.float64, "float64")
    dpa = _empty_like_kind(ref, (m,), np.float64, "float64") if want_channels else None
    dia = _empty_like_kind(ref, (m,), np.float64, "float64") if want_channels else None
    N.check(N.lib().sodso_loop_top1(c.handle, t, _ptr(hist1), m, _ptr(hist2), n, int(mask_width), float(p_weight),
                                    _ptr(idx), _ptr(score), _ptr(dpa), _ptr(dia)))
    return (idx, score, dpa, dia) if want_channels else (idx, score)


# ----------------------------------------------------------------------------------------------
# DELIGHT
# ----------------------------------------------------------------------------------------------
def delight_generate(xyz, inten, off, ctx=None):
    """test_delight.cpp:38-56: history_delight (16*nscan x 256)."""
    xyz = _prep(xyz, np.float64, "float64")
    inten = _prep(inten, np.float32, "float32")
    off = _prep(off, np.int64, "int64")
    nscan = off.shape[0] - 1
    c = _ctx_for(xyz, ctx=ctx)
    hist = _empty_like_kind(xyz, (16 * nscan, 256), np.float64, "float64")
    N.check(N.lib().sodso_delight_generate(c.handle, _ptr(xyz), _ptr(inten), _ptr(off), nscan, _ptr(hist)))
    return hist


def processDELIGHT(hist1, hist2, ctx=None):
    """dist = processDELIGHT(hist1, hist2)  (processDELIGHT.m:1-38)."""
    hist1 = _prep(hist1, np.float64, "float64")
    hist2 = _prep(hist2, np.float64, "float64")
    m, n = hist1.shape[0] // 16, hist2.shape[0] // 16
    c = _ctx_for(hist1, hist2, ctx=ctx)
    ref = hist1 if _is_torch(hist1) else hist2
    d = _empty_like_kind(ref, (m, n), np.float64, "float64")
    N.check(N.lib().sodso_delight_match(c.handle, _ptr(hist1), m, _ptr(hist2), n, _ptr(d)))
    return d


def top1_single(dist, mask_width, ctx=None):
    """run_test.m:47-57 on one distance matrix -> (diff_idx 0-based int32[m], diff_v[m])."""
    dist = _prep(dist, np.float64, "float64")
    m, n = dist.shape
    c = _ctx_for(dist, ctx=ctx)
    idx = _empty_like_kind(dist, (m,), np.int32, "int32")
    score = _empty_like_kind(dist, (m,), np.float64, "float64")
    N.check(N.lib().sodso_top1_single(c.handle, _ptr(dist), m, n, int(mask_width), _ptr(idx), _ptr(score)))
    return idx, score


def gt_loops(gt1, gt2, loop_diff, mask_width, ctx=None):
    """run_test.m:3-22 -> (lp_gt: 0-based (i, j) rows, n_loops)."""
    gt1 = np.ascontiguousarray(gt1, dtype=np.float64).reshape(-1, 3)
    gt2 = np.ascontiguousarray(gt2, dtype=np.float64).reshape(-1, 3)
    m, n = gt1.shape[0], gt2.shape[0]
    c = _ctx_for(ctx=ctx)
    near = np.zeros(m, dtype=np.int32)
    flag = np.zeros(m, dtype=np.int32)
    cnt = C.c_int(0)
    N.check(N.lib().sodso_gt_loops(c.handle, _ptr(gt1), m, _ptr(gt2), n, float(loop_diff), int(mask_width), _ptr(near),
                                   _ptr(flag), C.byref(cnt)))
    ii = np.nonzero(flag)[0]
    return np.stack([ii, near[ii]], axis=1).astype(np.int64).reshape(-1, 2), int(cnt.value)


def run_test_full(type, hist1, hist2, gt1, gt2, loop_diff, mask_width, p_weight=2.0, ctx=None):
    """[AUC, top_recall, lp_detected] = run_test(type, hist1, hist2, gt1, gt2, loop_diff, mask_width)
    (run_test.m:1-85; lp_detected 0-based).  Match / fusion / argmin and the ground-truth loop search run on the
    GPU, the cumulative precision-recall walk on the host."""
    gt1 = np.ascontiguousarray(gt1, dtype=np.float64).reshape(-1, 3)
    gt2 = np.ascontiguousarray(gt2, dtype=np.float64).reshape(-1, 3)
    lp_gt, n_loops = gt_loops(gt1, gt2, loop_diff, mask_width, ctx=ctx)
    idx, score = run_test(type, hist1, hist2, mask_width, p_weight, ctx=ctx)
    if _is_torch(idx):
        idx, score = idx.cpu().numpy(), score.cpu().numpy()
    idx = np.ascontiguousarray(idx, dtype=np.int32)
    score = np.ascontiguousarray(score, dtype=np.float64)
    m = idx.shape[0]
    auc, tr, tc = C.c_double(0), C.c_double(0), C.c_int(0)
    rank = np.zeros(m, dtype=np.int32)
    N.check(N.lib().sodso_pr_curve(_ptr(score), _ptr(idx), _ptr(gt1), m, _ptr(gt2), gt2.shape[0], float(loop_diff),
                                   n_loops, C.byref(auc), C.byref(tr), C.byref(tc), _ptr(rank), None, None))
    k = int(tc.value)
    lp_detected = np.stack([rank[:k], idx[rank[:k]]], axis=1).astype(np.int64).reshape(-1, 2)
    return float(auc.value), float(tr.value), lp_detected


def sc_scans_to_loops(xyz, inten, scan_off, mask_width, p_weight=2.0, max_rho=45.0, want_hist=False, ctx=None,
                      host_out=False):
    """test_sc.cpp:36-57 + run_test('sc', hist, hist, ...) (run_test.m:25-57, self-match as in test_kitti.m:28)
    in one call: scans in -> (diff_idx 0-based int32[n], diff_v[n] [, history_sc]).  Host (numpy / pinned torch)
    point buffers are streamed to the GPU chunk by chunk, overlapped with binning and matching.
    host_out: return diff_idx / diff_v as numpy arrays even for CUDA inputs (the library copies them out inside the
    call, one synchronisation instead of one per `.cpu()`)."""
    if _is_torch(xyz) and not xyz.is_cuda:      # pinned host tensors: pass as host pointers
        dev_like = None
    else:
        dev_like = xyz if _is_torch(xyz) else None
    xyz = _prep(xyz, np.float64, "float64")
    inten = _prep(inten, np.float32, "float32")
    scan_off = _prep(scan_off, np.int64, "int64")
    n = int(scan_off.shape[0]) - 1
    c = _ctx_for(*( [dev_like] if dev_like is not None else []), ctx=ctx)
    ref = dev_like if dev_like is not None else np.empty(0)
    out_ref = np.empty(0) if host_out else ref
    idx = _empty_like_kind(out_ref, (n,), np.int32, "int32")
    score = _empty_like_kind(out_ref, (n,), np.float64, "float64")
    hist = _empty_like_kind(ref, (n, 2 * SC_SIZE), np.float64, "float64") if want_hist else None
    N.check(N.lib().sodso_sc_scans_to_loops(c.handle, _ptr(xyz), _ptr(inten), _ptr(scan_off), n, float(max_rho),
                                            int(mask_width), float(p_weight), _ptr(hist), _ptr(idx), _ptr(score),
                                            None, None))
    return (idx, score, hist) if want_hist else (idx, score)


# ----------------------------------------------------------------------------------------------
# resident, row-sharded database
# ----------------------------------------------------------------------------------------------
class SignatureDB:
    """A shard of the signature database resident in HBM (sodso_db)."""

    def __init__(self, type, hist2, global_row0=0, ctx=None):
        self.type = {"sc": SODSO_TYPE_SC, "m2dp": SODSO_TYPE_M2DP}[type]
        rows = 1 if self.type == SODSO_TYPE_SC else 4
        if hist2 is None:     # an empty shard, to be filled by append()
            hist2 = np.empty((0, 2 * (SC_SIZE if self.type == SODSO_TYPE_SC else M2DP_SIZE)))
        hist2 = _prep(hist2, np.float64, "float64")
        self.n = hist2.shape[0] // rows
        self.row0 = int(global_row0)
        self.ctx = _ctx_for(hist2, ctx=ctx)
        self._h = C.c_void_p()
        self._rows = rows
        self.m = 0
        N.check(N.lib().sodso_db_create(self.ctx.handle, self.type, _ptr(hist2), self.n, self.row0, C.byref(self._h)))

    def close(self):
        if self._h:
            N.lib().sodso_db_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def append(self, hist_new):
        """sodso_db_append: further signatures behind the shard's last row (only their operand rows are written)"""
        hist_new = _prep(hist_new, np.float64, "float64")
        k = hist_new.shape[0] // self._rows
        _ctx_for(hist_new, ctx=self.ctx)
        N.check(N.lib().sodso_db_append(self._h, _ptr(hist_new), k))
        self.n += k

    def reserve(self, capacity: int):
        N.check(N.lib().sodso_db_reserve(self._h, int(capacity)))

    def query_sharded(self, hist1, q_global_row0=0, mask_width=100, p_weight=2.0, k=8, out=None):
        """sodso_db_query_sharded (collective over the context's communicator): -> (idx int64 [m,k] global 0-based,
        score, d_p, d_i).  numpy queries -> numpy results (one synchronisation); torch CUDA queries -> torch CUDA
        results, only enqueued (ctx.sync() or stream order before reading).  out: preallocated result tuple."""
        hist1 = _prep(hist1, np.float64, "float64")
        _ctx_for(hist1, ctx=self.ctx)
        self.m = hist1.shape[0] // self._rows
        self._ref = hist1
        if out is None:
            out = (_empty_like_kind(hist1, (self.m, k), np.int64, "int64"),) + tuple(
                _empty_like_kind(hist1, (self.m, k), np.float64, "float64") for _ in range(3))
        idx, score, dp, di = out
        N.check(N.lib().sodso_db_query_sharded(self._h, _ptr(hist1), self.m, int(q_global_row0), int(mask_width),
                                               float(p_weight), int(k), _ptr(idx), _ptr(score), _ptr(dp), _ptr(di)))
        return out

    def finish_sharded(self, q_global_row0=0, mask_width=100, p_weight=2.0, k=8, like=None):
        """sodso_db_finish_sharded for the last match / stream_match"""
        ref = like if like is not None else np.empty(0)
        idx = _empty_like_kind(ref, (self.m, k), np.int64, "int64")
        score, dp, di = (_empty_like_kind(ref, (self.m, k), np.float64, "float64") for _ in range(3))
        N.check(N.lib().sodso_db_finish_sharded(self._h, int(q_global_row0), int(mask_width), float(p_weight), int(k),
                                                _ptr(idx), _ptr(score), _ptr(dp), _ptr(di)))
        return idx, score, dp, di

    def scans_query_sharded(self, q_xyz, q_inten, q_off, m_total, q_first=0, db_scans=None, q_global_row0=0,
                            mask_width=100, p_weight=2.0, k=8, max_rho=45.0, want_hist=False, host_out=True):
        """sodso_db_scans_query_sharded: one step of the sharded pipeline from points (collective).
        q_*: this rank's slice [q_first, q_first + m_slice) of the m_total query scans (all of them: no exchange);
        db_scans: None (resident operand), (xyz, inten, off) of the shard's scans (rebuilt in place, streamed from
        host buffers), or "same": the shard's scans ARE the query scans (same buffers: copied and binned once).
        -> (idx, score, d_p, d_i [, q_hist])"""
        q_xyz = _prep(q_xyz, np.float64, "float64")
        q_inten = _prep(q_inten, np.float32, "float32")
        q_off = _prep(q_off, np.int64, "int64")
        m_slice = int(q_off.shape[0]) - 1
        dev_like = q_xyz if (_is_torch(q_xyz) and q_xyz.is_cuda) else None
        dx = di_ = do = None
        if isinstance(db_scans, str):
            assert db_scans == "same" and m_slice == int(m_total) == self.n
            dx, di_, do = q_xyz, q_inten, q_off
        elif db_scans is not None:
            dx, di_, do = (_prep(db_scans[0], np.float64, "float64"), _prep(db_scans[1], np.float32, "float32"),
                           _prep(db_scans[2], np.int64, "int64"))
            assert do.shape[0] - 1 == self.n
            if dev_like is None and _is_torch(dx) and dx.is_cuda:
                dev_like = dx
        _ctx_for(*([dev_like] if dev_like is not None else []), ctx=self.ctx)
        self.m = int(m_total)
        ref = np.empty(0) if (host_out or dev_like is None) else dev_like
        idx = _empty_like_kind(ref, (self.m, k), np.int64, "int64")
        score, dp, di = (_empty_like_kind(ref, (self.m, k), np.float64, "float64") for _ in range(3))
        hist = None
        if want_hist:
            hist = _empty_like_kind(dev_like if dev_like is not None else np.empty(0), (self.m, 2 * SC_SIZE), np.float64,
                                    "float64")
        N.check(N.lib().sodso_db_scans_query_sharded(
            self._h, _ptr(dx), _ptr(di_), _ptr(do), _ptr(q_xyz), _ptr(q_inten), _ptr(q_off), self.m, int(q_first),
            m_slice, float(max_rho), int(q_global_row0), int(mask_width), float(p_weight), int(k), _ptr(hist), _ptr(idx),
            _ptr(score), _ptr(dp), _ptr(di)))
        return (idx, score, dp, di, hist) if want_hist else (idx, score, dp, di)

    def reload(self, hist2):
        """new signatures for the same shard (same number of rows), operand buffers rewritten in place"""
        hist2 = _prep(hist2, np.float64, "float64")
        assert hist2.shape[0] // self._rows == self.n
        _ctx_for(hist2, ctx=self.ctx)
        N.check(N.lib().sodso_db_reload(self._h, _ptr(hist2)))

    def stream_match(self, xyz, inten, scan_off, hist1, max_rho=45.0):
        """reload + match in one streamed pass: the shard's scans as points (host buffers are streamed in chunks,
        binned and matched against hist1 as they land)"""
        xyz = _prep(xyz, np.float64, "float64")
        inten = _prep(inten, np.float32, "float32")
        scan_off = _prep(scan_off, np.int64, "int64")
        hist1 = _prep(hist1, np.float64, "float64")
        assert scan_off.shape[0] - 1 == self.n
        self.m = hist1.shape[0] // self._rows
        self._ref = hist1
        _ctx_for(hist1, xyz, ctx=self.ctx)
        N.check(N.lib().sodso_db_stream_match(self._h, _ptr(xyz), _ptr(inten), _ptr(scan_off), float(max_rho),
                                              _ptr(hist1), self.m))

    def match(self, hist1):
        hist1 = _prep(hist1, np.float64, "float64")
        _ctx_for(hist1, ctx=self.ctx)
        self.m = hist1.shape[0] // self._rows
        self._ref = hist1
        N.check(N.lib().sodso_db_match(self._h, _ptr(hist1), self.m))

    def partial_stats(self, like=None):
        ref = like if like is not None else self._ref
        st = _empty_like_kind(ref, (self.m, N.STATS_W), np.float64, "float64")
        N.check(N.lib().sodso_db_partial_stats(self._h, _ptr(st)))
        return st

    def topk(self, global_stats, n_global, q_global_row0, mask_width, p_weight=2.0, k=8):
        gs = _prep(global_stats, np.float64, "float64")
        _ctx_for(gs, ctx=self.ctx)
        idx = _empty_like_kind(gs, (self.m, k), np.int64, "int64")
        score = _empty_like_kind(gs, (self.m, k), np.float64, "float64")
        dp = _empty_like_kind(gs, (self.m, k), np.float64, "float64")
        di = _empty_like_kind(gs, (self.m, k), np.float64, "float64")
        N.check(N.lib().sodso_db_topk(self._h, _ptr(gs), int(n_global), int(q_global_row0), int(mask_width),
                                      float(p_weight), int(k), _ptr(idx), _ptr(score), _ptr(dp), _ptr(di)))
        return idx, score, dp, di

    def distances(self, like=None):
        """fp32 (m x n_local) copies of the last match result (numpy, or torch if `like` is a CUDA tensor)."""
        ref = like if like is not None else np.empty(0)
        dp = _empty_like_kind(ref, (self.m, self.n), np.float32, "float32")
        di = _empty_like_kind(ref, (self.m, self.n), np.float32, "float32")
        N.check(N.lib().sodso_db_get_distances(self._h, _ptr(dp), _ptr(di)))
        return dp, di


def topk_merge_device(idx, score, d_p, d_i, ctx=None):
    """topk_merge for gathered lists in HBM (R x m x k torch CUDA tensors) -> (m x k) torch CUDA tensors."""
    import torch

    idx, score, d_p, d_i = (t.contiguous() for t in (idx, score, d_p, d_i))
    R, m, k = idx.shape
    c = _ctx_for(idx, score, ctx=ctx)
    oi = torch.empty((m, k), dtype=torch.int64, device=idx.device)
    os_, op, od = (torch.empty((m, k), dtype=torch.float64, device=idx.device) for _ in range(3))
    N.check(N.lib().sodso_topk_merge_device(c.handle, _ptr(idx), _ptr(score), _ptr(d_p), _ptr(d_i), R, m, k, _ptr(oi),
                                            _ptr(os_), _ptr(op), _ptr(od)))
    return oi, os_, op, od


def topk_merge(idx, score, d_p, d_i):
    """Merge gathered per-shard top-k lists (R x m x k numpy arrays) -> (m x k) each."""
    idx = np.ascontiguousarray(idx, dtype=np.int64)
    score = np.ascontiguousarray(score, dtype=np.float64)
    d_p = np.ascontiguousarray(d_p, dtype=np.float64)
    d_i = np.ascontiguousarray(d_i, dtype=np.float64)
    R, m, k = idx.shape
    oi = np.empty((m, k), dtype=np.int64)
    os_ = np.empty((m, k))
    op = np.empty((m, k))
    od = np.empty((m, k))
    N.check(N.lib().sodso_topk_merge(_ptr(idx), _ptr(score), _ptr(d_p), _ptr(d_i), R, m, k, _ptr(oi), _ptr(os_),
                                     _ptr(op), _ptr(od)))
    return oi, os_, op, od
