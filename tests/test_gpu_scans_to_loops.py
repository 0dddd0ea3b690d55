"""GPU parity of the one-call path scans -> loop candidates (sodso_sc_scans_to_loops = test_sc.cpp:36-57 +
run_test.m:25-57 self-match): against the oracle at a size it finishes in seconds, and the streamed
(host buffers, chunked copies overlapped with block-wise matching) variant against the unstreamed one
at a size where streaming is active."""
import numpy as np
import pytest
import torch

from so_dso_place_recognition_b200 import api, synth

pytestmark = pytest.mark.gpu


def test_small_vs_oracle(gpu_ctx, oracle):
    xyz, inten, off = synth.make_scan_set(96, 1024, planted_loops=True)
    idx, score, hist = api.sc_scans_to_loops(xyz, inten, off, 5, want_hist=True)
    ref = oracle.sc_generate(xyz, inten, off, nthreads=8)
    assert np.array_equal(hist[:, 1200:], ref[:, 1200:])
    assert np.abs(hist[:, :1200] - ref[:, :1200]).max() < 1e-9
    rp, ri = oracle.sc_match_numpy(ref, ref)
    ridx, rscore = oracle.fuse_top1(rp, ri, 5)
    assert np.array_equal(idx, ridx)
    sp, si_ = np.nanstd(rp, axis=1, ddof=1), np.nanstd(ri, axis=1, ddof=1)
    assert np.all(np.abs(score - rscore) <= 1e-6 + 4e-5 * (2.0 / sp + 1.0 / si_))
    assert (idx == (np.arange(96) + 48) % 96).all()          # the planted loops


@pytest.mark.parametrize("nscan", [2048, 2100, 2561])
def test_streamed_equals_unstreamed(gpu_ctx, nscan):
    """>= 2048 scans in host memory -> 512-scan chunks; last chunk ragged (52 / 1 scans)."""
    xyz, inten, off = synth.make_scan_set(nscan, 192, planted_loops=True)
    # ragged scan sizes: drop a few points from some scans
    keep = np.ones(len(inten), bool)
    for s in range(0, nscan, 7):
        keep[off[s]:off[s] + (s % 5)] = False
    sizes = np.array([keep[off[s]:off[s + 1]].sum() for s in range(nscan)])
    xyz, inten = xyz[keep], inten[keep]
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)

    l0 = gpu_ctx.launch_count
    idx_s, score_s, hist_s = api.sc_scans_to_loops(xyz, inten, off, 100, want_hist=True)   # host buffers: streamed
    launches_streamed = gpu_ctx.launch_count - l0
    d = lambda a: torch.from_numpy(a).cuda()
    l0 = gpu_ctx.launch_count
    idx_u, score_u, hist_u = api.sc_scans_to_loops(d(xyz), d(inten), d(off), 100, want_hist=True)  # device: one block
    launches_unstreamed = gpu_ctx.launch_count - l0
    assert launches_streamed > launches_unstreamed
    assert np.array_equal(hist_s, hist_u.cpu().numpy())
    assert np.array_equal(idx_s, idx_u.cpu().numpy())
    assert np.array_equal(score_s, score_u.cpu().numpy())
    # and the two-call path (generate, then run_test) gives the same answer
    idx_t, score_t = api.run_test("sc", hist_u, hist_u, 100)
    assert np.array_equal(idx_s, idx_t.cpu().numpy())
    assert np.array_equal(score_s, score_t.cpu().numpy())
    # pinned host tensors take the streamed path too
    pin = lambda a: torch.from_numpy(a).pin_memory()
    idx_p, score_p = api.sc_scans_to_loops(pin(xyz), pin(inten), pin(off), 100)
    assert np.array_equal(idx_p, idx_s) and np.array_equal(score_p, score_s)
    # device inputs, results copied out inside the call
    idx_h, score_h = api.sc_scans_to_loops(d(xyz), d(inten), d(off), 100, host_out=True)
    assert isinstance(idx_h, np.ndarray) and np.array_equal(idx_h, idx_s) and np.array_equal(score_h, score_s)


def test_db_stream_match_equals_reload_match(gpu_ctx):
    """resident shard: streamed reload + match from HOST point buffers == sodso_db_reload + sodso_db_match"""
    n, m = 2300, 96
    xyz, inten, off = synth.make_scan_set(n, 160, planted_loops=True)
    hist = api.sc_generate(xyz, inten, off)
    db = api.SignatureDB("sc", hist, global_row0=0)
    q = hist[n // 2:n // 2 + m]
    db.match(q)
    dp0, di0 = db.distances()
    db.reload(np.zeros_like(hist))                      # wipe the operand
    db.stream_match(xyz, inten, off, q)                 # host buffers: 5 chunks, the last one ragged
    dp1, di1 = db.distances()
    assert np.array_equal(dp0, dp1) and np.array_equal(di0, di1)
    st = db.partial_stats()
    idx, score, _, _ = db.topk(st, n, n // 2, 100, 2.0, 4)
    assert (idx[:, 0] == np.arange(m)).mean() > 0.5     # planted loops (sparse 160-point scans: not all of them)
    db.close()
