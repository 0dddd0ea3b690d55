"""BASELINE.json full-size runs (configs[2] and configs[4]) checked through size-independent properties, plus
oracle spot checks on a few rows -- the CPU oracle cannot do 25 x 10^6 pairs in test time.

Scan Context (processSC.m): d(i, j) = min over the 120 shift / reversal variants is symmetric in (i, j) (shifting the
query by s is shifting the DB entry by -s, and the reversal is an involution), d(i, i) = 0, planted revisits are the
top-1.  M2DP (processM2DP.m): every signature half is a unit vector with non-negative entries (Perron pair), so the
un-normalised distance of a scan to itself is (1 - 2)/2 = -0.5."""
import numpy as np
import pytest
import torch

from so_dso_place_recognition_b200 import api, synth

pytestmark = pytest.mark.gpu
N, NPTS = 5000, 4096


@pytest.fixture(scope="module")
def scans():
    xyz, inten, off = synth.make_scan_set(N, NPTS, planted_loops=True)
    d = lambda a: torch.from_numpy(a).cuda()
    return xyz, inten, off, d(xyz), d(inten), d(off)


def test_sc_5k_all_pairs_properties(gpu_ctx, oracle, scans):
    xyz, inten, off, dx, di_, do = scans
    sig = api.sc_generate(dx, di_, do)
    dp, di = api.processSC(sig, sig, f32=True)                       # 5000 x 5000, tcgen05 path
    assert gpu_ctx.last_kernel_name == "sc_match_tc_kernel"
    assert dp.shape == (N, N) and not torch.isnan(dp).any() and not torch.isnan(di).any()
    assert float((dp - dp.T).abs().max()) < 4e-6 and float((di - di.T).abs().max()) < 4e-6
    assert float(dp.diagonal().abs().max()) < 2e-6 and float(di.diagonal().abs().max()) < 2e-6
    assert float(dp.min()) > -2e-6 and float(dp.max()) <= 0.5 + 2e-6    # cosine distance of non-negative vectors
    # oracle spot check: a few query rows against the whole database
    rows = [0, 1234, 2500, 4999]
    sig_h = sig.cpu().numpy()
    np.testing.assert_array_equal(sig_h[rows, 1200:], oracle.sc_generate(
        np.concatenate([xyz[off[r]:off[r + 1]] for r in rows]), np.concatenate([inten[off[r]:off[r + 1]] for r in rows]),
        np.arange(len(rows) + 1, dtype=np.int64) * NPTS)[:, 1200:])
    rp, ri = oracle.sc_match_numpy(sig_h[rows], sig_h)
    assert np.abs(dp[rows].cpu().numpy() - rp).max() < 1e-5 and np.abs(di[rows].cpu().numpy() - ri).max() < 1e-5
    # decision: planted revisits are the top-1 for every query, and the one-call path agrees
    idx, score = api.run_test("sc", sig, sig, 100)
    expect = (np.arange(N) + N // 2) % N
    assert (idx.cpu().numpy() == expect).all()
    idx2, score2 = api.sc_scans_to_loops(xyz, inten, off, 100)           # host buffers: streamed path
    assert np.array_equal(idx2, expect) and np.array_equal(score2, score.cpu().numpy())


def test_m2dp_5k_properties(gpu_ctx, oracle, scans):
    xyz, inten, off, dx, di_, do = scans
    sig = api.m2dp_generate(dx, di_, do)                             # 20000 x 384
    assert sig.shape == (4 * N, 384)
    parts = sig.view(4 * N, 2, 192)
    u, v = parts[..., :64], parts[..., 64:]
    assert float((u.norm(dim=-1) - 1).abs().max()) < 1e-12 and float((v.norm(dim=-1) - 1).abs().max()) < 1e-12
    assert float(sig.min()) > -1e-12                                 # Perron vectors
    # oracle spot check on two scans (4 variants each)
    for r in (7, 4321):
        ref = oracle.m2dp_generate(xyz[off[r]:off[r + 1]], inten[off[r]:off[r + 1]], np.array([0, NPTS], dtype=np.int64))
        np.testing.assert_allclose(sig[4 * r:4 * r + 4].cpu().numpy(), ref, rtol=0, atol=1e-9)
    dp, di = api.processM2DP(sig, sig, f32=True)                     # 5000 x 5000
    assert float((dp.diagonal() + 0.5).abs().max()) < 2e-6 and float((di.diagonal() + 0.5).abs().max()) < 2e-6
    assert float((dp - dp.T).abs().max()) < 2e-6
    idx, score = api.run_test("m2dp", sig, sig, 100)
    assert (idx.cpu().numpy() == (np.arange(N) + N // 2) % N).mean() > 0.9
