"""Hand-constructed known-answer tests of the oracle: each pins one
behaviour of the reference code that the GPU path must reproduce."""
import numpy as np
import pytest

from so_dso_place_recognition_b200 import synth


def _frame_points(y, z, x):
    """points given directly in the PCA frame (x up = least variance) -> a raw cloud whose PCA frame is
    that frame: add a symmetric scaffold so that mean = 0 and axes are x<y<z in variance."""
    pts = np.stack([x, y, z], axis=1)
    return pts


def test_sc_single_bin_and_height_range(oracle):
    # 4 symmetric far points fix the PCA frame (variance x < y < z), 2 probe points share a bin
    base = np.array([[0.0, 10.0, 0.0], [0.0, -10.0, 0.0], [0.0, 0.0, 30.0], [0.0, 0.0, -30.0],
                     [0.5, 0.0, 0.0], [-0.5, 0.0, 0.0]])
    probe = np.array([[1.0, 3.0, 4.0], [-1.0, 3.0, 4.0], [1.0, -3.0, -4.0], [-1.0, -3.0, -4.0]])
    # probes come in +/- pairs w.r.t. (y, z) and their x sums to zero => mean stays 0
    xyz = np.concatenate([base, probe])
    assert np.allclose(xyz.mean(0), 0)
    inten = np.arange(len(xyz), dtype=np.float32)
    s, i = oracle.sc_signature(xyz, inten)
    al, ev, mean = oracle.align_pca(xyz)
    # SC.cpp:33-39 on the aligned points
    yp, zp = al[:, 1], al[:, 2]
    si = np.floor((np.arctan2(zp, yp) + np.pi) * 60 / (2 * np.pi)).astype(int)
    ri = np.floor(np.sqrt(yp * yp + zp * zp) * 20 / 45.0).astype(int)
    idx = si * 20 + ri
    b = idx[6]
    assert idx[7] == b
    assert s[b] == pytest.approx(abs(al[6, 0] - al[7, 0]), abs=1e-12) and s[b] == pytest.approx(2.0, abs=1e-9)
    # single-point bins have structure 0 (SC.cpp:46-49,74)
    for k in range(6):
        if (idx == idx[k]).sum() == 1:
            assert s[idx[k]] == 0.0
    # the two probe pairs + the (+-0.5, 0, 0) pair at the origin (atan2(0,0) = 0 -> sector 30, ring 0)
    assert (s != 0).sum() == 3 and s[30 * 20 + 0] == pytest.approx(1.0, abs=1e-12)


def test_sc_ring_aliasing_quirk(oracle):
    """SC.cpp:42 only range-checks the flat index: a point with ri >= 20 in sector si < 59 lands in
    sector si+1."""
    rng = np.random.default_rng(5)
    xyz, inten = synth.make_scan(3, 2048)
    # push a few points beyond 45 m in the PCA (y,z) plane but keep the cloud's frame
    far = xyz.copy()
    sel = rng.choice(len(far), 16, replace=False)
    r = np.linalg.norm(far[sel][:, [0, 2]], axis=1)
    far[sel, 0] *= 46.5 / r
    far[sel, 2] *= 46.5 / r
    al, _, _ = oracle.align_pca(far)
    yp, zp = al[:, 1], al[:, 2]
    ri = np.floor(np.sqrt(yp * yp + zp * zp) * 20 / 45.0).astype(int)
    si = np.floor((np.arctan2(zp, yp) + np.pi) * 60 / (2 * np.pi)).astype(int)
    assert (ri >= 20).sum() >= 8
    idx = si * 20 + ri
    keep = (idx >= 0) & (idx < 1200)
    s, i = oracle.sc_signature(far, inten)
    # occupancy must equal the aliased (not the ring-clipped) index set
    occ_alias = np.zeros(1200, bool)
    occ_alias[idx[keep]] = True
    cnt = np.bincount(idx[keep], minlength=1200)
    multi = cnt >= 2
    assert ((s != 0) <= occ_alias).all()
    aliased = keep & (ri >= 20)
    assert aliased.sum() >= 4
    # an aliased point shares the bin (si+1, ri-20); with >= 2 points there the height range is > 0
    hit = [k for k in np.where(aliased)[0] if multi[idx[k]]]
    assert len(hit) > 0 and all(s[idx[k]] > 0 for k in hit)


def test_sc_match_shift_and_reverse(oracle):
    """processSC.m:24-31: a sector-rotated copy and a column-reversed copy both match at d ~ 0."""
    rng = np.random.default_rng(1)
    img = rng.random((60, 20)) * (rng.random((60, 20)) < 0.3)
    inten = (rng.random((60, 20)) < 0.2).astype(float)
    inten[0, 0] = 1

    def row(a, b):
        return np.concatenate([a.reshape(-1), b.reshape(-1)])[None, :]

    q = row(img, inten)
    rot = row(np.roll(img, 17, axis=0), np.roll(inten, 17, axis=0))
    rev = row(img[::-1], inten[::-1])
    other = row(rng.random((60, 20)), (rng.random((60, 20)) < 0.2).astype(float))
    db = np.concatenate([rot, rev, other, q])
    dp, di = oracle.sc_match(q, db)
    assert dp[0, [0, 1, 3]].max() < 1e-15 and di[0, [0, 1, 3]].max() < 1e-15
    assert dp[0, 2] > 0.05
    # numpy/BLAS restatement agrees with the C loops
    dp2, di2 = oracle.sc_match_numpy(q, db)
    np.testing.assert_allclose(dp, dp2, atol=1e-14)
    np.testing.assert_allclose(di, di2, atol=1e-14)


def test_sc_match_zero_norm_is_nan(oracle):
    """processSC.m:15-20: x/norm(x) of an all-zero row is NaN; MATLAB min skips NaN unless all are."""
    rng = np.random.default_rng(2)
    h = rng.random((3, 2400))
    h[1, :1200] = 0
    dp, di = oracle.sc_match(h, h)
    assert np.isnan(dp[1]).all() and np.isnan(dp[:, 1]).all()
    assert not np.isnan(di).any()


def test_m2dp_degenerate_plane_row(oracle):
    """M2DP.cpp:10-25: plane (p=2,q=0) has normal (1,0,0): xProj = yProj = 0, every point falls into
    bin si=8, ri=0 of row 32."""
    xyz, inten = synth.make_scan(11, 1024)
    al, _, _ = oracle.align_pca(xyz)
    c, i, Ac, Ai = oracle.m2dp_signature(al, inten, want_hist=True)
    row = Ac[32]
    # xp = yp = +-0: atan2(+0,+0) = 0 -> si 8; atan2(+0,-0) = pi -> si 16 (aliases into ring 1, M2DP.cpp:63-68);
    # atan2(-0,-0) = -pi -> si 0.  The signs of the zeros depend on the signs of the coordinates.
    assert row.sum() == 1024 and set(np.nonzero(row)[0]) <= {0, 8, 16} and row[8] > 512
    xp, yp = oracle.m2dp_tables()
    assert np.all(xp[32] == 0) and np.all(yp[32] == 0)
    assert abs(np.linalg.norm(c[:64]) - 1) < 1e-12 and abs(np.linalg.norm(c[64:]) - 1) < 1e-12
    assert c[:64].sum() >= 0


def test_svd_dominant_against_numpy(oracle):
    rng = np.random.default_rng(3)
    A = rng.integers(0, 40, (64, 128)).astype(float)
    u, v, s = oracle.svd_dominant(A)
    U, S, Vt = np.linalg.svd(A, full_matrices=False)
    sg = 1.0 if U[:, 0].sum() >= 0 else -1.0
    assert s == pytest.approx(S[0], rel=1e-12)
    np.testing.assert_allclose(u, sg * U[:, 0], atol=1e-10)
    np.testing.assert_allclose(v, sg * Vt[0], atol=1e-10)


def test_m2dp_match_block_min(oracle):
    rng = np.random.default_rng(4)
    h1 = rng.normal(size=(8, 384))
    h2 = rng.normal(size=(12, 384))
    dp, di = oracle.m2dp_match(h1, h2)
    full = (1 - h1[:, :192] @ h2[:, :192].T) / 2
    ref = full.reshape(2, 4, 3, 4).min(axis=(1, 3))
    np.testing.assert_allclose(dp, ref, atol=1e-13)


def test_fuse_top1_against_numpy(oracle):
    rng = np.random.default_rng(6)
    dp = rng.random((40, 300)) * 0.5
    di = rng.random((40, 300)) * 0.5
    idx, sc = oracle.fuse_top1(dp, di, 10)
    idx2, sc2 = oracle.fuse_top1_numpy(dp, di, 10)
    np.testing.assert_array_equal(idx, idx2)
    np.testing.assert_allclose(sc, sc2, rtol=1e-12)
    assert (np.abs(idx - np.arange(40)) >= 10).all()
    # ties: first index wins (run_test.m:57)
    dp[:, 7] = -1.0
    dp[:, 200] = -1.0
    di[:, 7] = -1.0
    di[:, 200] = -1.0
    idx, _ = oracle.fuse_top1(dp, di, 0)
    assert (idx == 7).all()
    # a constant channel has std 0 -> 0/0 = NaN everywhere -> MATLAB min returns the first index
    di[:] = 0.25
    idx, sc = oracle.fuse_top1(dp, di, 0)
    assert (idx == 0).all() and np.isnan(sc).all()


def test_planted_loops_are_recovered(oracle):
    """The synthetic set of SURVEY §8d: scan j+n/2 revisits scan j under a random yaw."""
    n = 24
    xyz, inten, off = synth.make_scan_set(n, 2048, planted_loops=True)
    hist = oracle.sc_generate(xyz, inten, off, nthreads=8)
    dp, di = oracle.sc_match_numpy(hist, hist)
    idx, score = oracle.fuse_top1(dp, di, 3)
    expect = (np.arange(n) + n // 2) % n
    assert (idx == expect).mean() >= 0.9


def test_fuse_omits_nan_entries_from_row_statistics(oracle):
    """run_test.m:40: MATLAB's normalize computes mean / std with 'omitnan'.  One zero-norm DB signature (e.g. an empty
    scan, processSC.m:15-20) gives a NaN column; every other candidate must still be ranked exactly as if that column
    were not there, and the NaN column is never chosen."""
    rng = np.random.default_rng(21)
    m, n, bad = 40, 90, 33
    dp, di = rng.random((m, n)) * 0.5, rng.random((m, n)) * 0.5
    dpn, din = dp.copy(), di.copy()
    dpn[:, bad] = np.nan
    din[:, bad] = np.nan
    idx, sc = oracle.fuse_top1(dpn, din, 5)
    keep = np.arange(n) != bad
    # without the column: same statistics; the mask must use the ORIGINAL column numbers -> emulate with +inf rows
    ridx, rsc, fused = oracle.fuse_top1(dp[:, keep], di[:, keep], 0, want_fused=True)
    cols = np.arange(n)[keep]
    fused = np.where(np.abs(np.arange(m)[:, None] - cols[None, :]) < 5, np.inf, fused)
    want = cols[np.argmin(fused, axis=1)]
    assert np.array_equal(idx, want) and not (idx == bad).any()
    np.testing.assert_allclose(sc, fused.min(axis=1), rtol=1e-12)
    # numpy twin used by the bench's CPU arm agrees
    nidx, _ = oracle.fuse_top1_numpy(dpn, din, 5)
    assert np.array_equal(nidx, want)
    # one channel NaN only: that channel's entry is NaN -> the fused entry is NaN -> skipped
    dpo = dp.copy()
    dpo[3, 7] = np.nan
    idx2, _ = oracle.fuse_top1(dpo, di, 0)
    assert idx2[3] != 7 or np.argmin(np.delete(dp[3], 7)) != 7
