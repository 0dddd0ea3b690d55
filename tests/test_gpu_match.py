"""GPU parity: Scan Context matching (processSC.m:1-45), fusion + decision (run_test.m:38-57)
through the C ABI vs the CPU oracle.

Bar (BASELINE.json north_star): SC distances within 1e-5 of the fp64 reference path; identical
integer top-1 loop indices."""
import numpy as np
import pytest

from so_dso_place_recognition_b200 import api, synth

pytestmark = pytest.mark.gpu
TOL_D = 1e-5
ALGOS = [api.SODSO_ALGO_TC, api.SODSO_ALGO_SIMT]


@pytest.fixture(scope="module")
def sigs(oracle):
    """oracle signatures of a planted-loop synthetic set (match parity decoupled from generation)"""
    xyz, inten, off = synth.make_scan_set(384, 2048, planted_loops=True)
    return oracle.sc_generate(xyz, inten, off, nthreads=16)


@pytest.fixture(autouse=True)
def _restore_algo(gpu_ctx):
    yield
    gpu_ctx.set_match_algo(api.SODSO_ALGO_TC)


@pytest.mark.parametrize("algo", ALGOS)
def test_sc_match_vs_oracle(gpu_ctx, oracle, sigs, algo):
    gpu_ctx.set_match_algo(algo)
    q, db = sigs[:40], sigs[:300]
    dp, di = api.processSC(q, db)
    rp, ri = oracle.sc_match_numpy(q, db)
    assert np.abs(dp - rp).max() < TOL_D and np.abs(di - ri).max() < TOL_D
    dp32, di32 = api.processSC(q, db, f32=True)
    assert dp32.dtype == np.float32 and np.abs(dp32 - rp).max() < TOL_D


@pytest.mark.parametrize("algo", ALGOS)
@pytest.mark.parametrize("m,n", [(1, 1), (3, 129), (130, 7), (257, 255)])
def test_sc_match_ragged_shapes(gpu_ctx, oracle, sigs, algo, m, n):
    gpu_ctx.set_match_algo(algo)
    q, db = sigs[100:100 + m], sigs[5:5 + n]
    dp, di = api.processSC(q, db)
    rp, ri = oracle.sc_match_numpy(q, db)
    assert dp.shape == (m, n)
    assert np.abs(dp - rp).max() < TOL_D and np.abs(di - ri).max() < TOL_D


@pytest.mark.parametrize("algo", ALGOS)
def test_sc_match_shift_reverse_and_nan(gpu_ctx, algo):
    gpu_ctx.set_match_algo(algo)
    rng = np.random.default_rng(1)
    img = rng.random((60, 20)) * (rng.random((60, 20)) < 0.3)
    it = (rng.random((60, 20)) < 0.2).astype(float)

    def row(a, b):
        return np.concatenate([a.reshape(-1), b.reshape(-1)])[None, :]

    q = row(img, it)
    db = np.concatenate([row(np.roll(img, k, axis=0), np.roll(it, k, axis=0)) for k in range(60)] +
                        [row(np.roll(img[::-1], k, axis=0), np.roll(it[::-1], k, axis=0)) for k in range(60)] +
                        [row(np.zeros((60, 20)), it)])
    dp, di = api.processSC(q, db)
    assert np.abs(dp[0, :120]).max() < 2e-6 and np.abs(di[0, :120]).max() < 2e-6
    assert np.isnan(dp[0, 120]) and abs(di[0, 120]) < 2e-6      # zero-norm row -> NaN (processSC.m:15-20)


@pytest.mark.parametrize("algo", ALGOS)
def test_loop_top1_identical_to_oracle(gpu_ctx, oracle, sigs, algo):
    gpu_ctx.set_match_algo(algo)
    rp, ri = oracle.sc_match_numpy(sigs, sigs)
    ref_idx, ref_score = oracle.fuse_top1(rp, ri, 20)
    idx, score, dpa, dia = api.run_test("sc", sigs, sigs, 20, want_channels=True)
    np.testing.assert_array_equal(idx, ref_idx)
    np.testing.assert_allclose(score, ref_score, rtol=0, atol=2e-3)
    n = sigs.shape[0]
    assert (idx == (np.arange(n) + n // 2) % n).mean() > 0.95
    np.testing.assert_allclose(dpa, rp[np.arange(n), ref_idx], atol=TOL_D)
    np.testing.assert_allclose(dia, ri[np.arange(n), ref_idx], atol=TOL_D)


def test_fuse_top1_given_matrices(gpu_ctx, oracle):
    rng = np.random.default_rng(6)
    dp = rng.random((70, 1000)) * 0.5
    di = rng.random((70, 1000)) * 0.5
    dp[:, 7] = -1.0
    dp[:, 200] = -1.0
    di[:, 7] = di[:, 200] = -1.0
    dp[3, 50] = np.nan                                  # one NaN entry: omitted from the row statistics, row still ranked
    dp[4, :] = np.nan                                   # an all-NaN channel row
    for mw in (0, 10, 100):
        idx, sc = api.fuse_top1(dp, di, mw)
        with np.errstate(all="ignore"):
            ridx, rsc = oracle.fuse_top1(dp, di, mw)
        np.testing.assert_array_equal(idx, ridx)
        np.testing.assert_allclose(sc, rsc, rtol=1e-10, equal_nan=True)
        assert idx[3] == (7 if mw == 0 else 200) and np.isfinite(sc[3])     # MATLAB normalize omits NaN (run_test.m:40)
        if mw == 0:                                     # all-NaN row: MATLAB min returns the first index
            assert idx[4] == 0 and np.isnan(sc[4])
        else:                                           # ... but the mask overwrites NaN with Inf (run_test.m:47-53)
            assert idx[4] == 0 and np.isinf(sc[4])


def test_tc_equals_simt_at_scale(gpu_ctx, oracle):
    """Cross-check of the tensor-core path against the fp32 CUDA-core kernel on the GPU at a size the CPU
    oracle cannot do in test time, plus an oracle spot check on a few rows."""
    xyz, inten, off = synth.make_scan_set(1500, 1024, planted_loops=True)
    sig = api.sc_generate(xyz, inten, off)
    gpu_ctx.set_match_algo(api.SODSO_ALGO_TC)
    dp, di = api.processSC(sig, sig, f32=True)
    gpu_ctx.set_match_algo(api.SODSO_ALGO_SIMT)
    sp, si = api.processSC(sig, sig, f32=True)
    assert np.abs(dp - sp).max() < TOL_D and np.abs(di - si).max() < TOL_D
    rows = [0, 749, 1499]
    rp, ri = oracle.sc_match_numpy(sig[rows], sig)
    assert np.abs(dp[rows] - rp).max() < TOL_D and np.abs(di[rows] - ri).max() < TOL_D
    # symmetry-like property: d(q, q) == 0 on the diagonal (shift 0 variant)
    assert np.abs(np.diag(dp)).max() < 2e-6


def test_channel_modes_binary_and_generic(gpu_ctx, oracle, sigs):
    """The matcher picks its arithmetic per channel on the device: e2m1 exact counts when every value of a channel is 0 / 1
    on both sides, the 3-term fp16 split otherwise.  All four combinations, including a real-valued 'intensity' channel,
    a binary 'structure' channel and a binary query against a non-binary database, against the oracle."""
    rng = np.random.default_rng(8)
    base = sigs[:150].copy()
    real_i = base.copy()
    real_i[:, 1200:] *= rng.uniform(0.5, 3.0, (150, 1200))            # intensity channel no longer binary
    bin_s = base.copy()
    bin_s[:, :1200] = (base[:, :1200] > 0.5).astype(float)             # structure channel binary too
    mixed_db = base.copy()
    mixed_db[7, 1200 + 5] = 0.25                                        # one non-binary value in the whole database
    for q, db in ((real_i[:33], real_i), (bin_s[:33], bin_s), (base[:33], mixed_db), (mixed_db[:33], base)):
        dp, di = api.processSC(q, db)
        rp, ri = oracle.sc_match_numpy(q, db)
        assert np.nanmax(np.abs(dp - rp)) < TOL_D and np.nanmax(np.abs(di - ri)) < TOL_D
    # a single non-zero bin per signature: the worst case for any reduced-precision scheme (no averaging of rounding errors)
    one = np.zeros((60, 2400))
    k = rng.integers(0, 1200, 60)
    one[np.arange(60), k] = rng.uniform(0.1, 9.0, 60)
    one[np.arange(60), 1200 + k] = 1.0
    dp, di = api.processSC(one, one)
    rp, ri = oracle.sc_match_numpy(one, one)
    assert np.abs(dp - rp).max() < 1e-6 and np.abs(di - ri).max() < 1e-6


def _topk_numpy(dp, di, q0, row0, mask, k):
    n = dp.shape[1]
    out = []
    for i in range(dp.shape[0]):
        with np.errstate(invalid="ignore", divide="ignore"):
            mu_p, mu_i = np.nanmean(dp[i]), np.nanmean(di[i])       # MATLAB normalize omits NaN (run_test.m:40)
            sd_p, sd_i = np.nanstd(dp[i], ddof=1), np.nanstd(di[i], ddof=1)
            f = 2.0 * ((dp[i] - mu_p) / sd_p) + (di[i] - mu_i) / sd_i
        jg = row0 + np.arange(n)
        f[np.abs((q0 + i) - jg) < mask] = np.inf
        order = [j for j in np.lexsort((jg, f)) if not np.isnan(f[j])][:k]
        out.append((order, f))
    return out


@pytest.mark.parametrize("k", [1, 3, 8])
@pytest.mark.parametrize("n_db", [384, 12])
def test_topk_vs_numpy_on_device_distances(gpu_ctx, sigs, k, n_db):
    """sodso_db_topk (run_test.m:38-57 generalised to the k best) against a numpy restatement on the SAME fp32
    distances: indices identical incl. order, scores to 1e-9; with the mask, a NaN query row, global row offsets as a
    shard would have them, and (12-row DB, mask 5) rows whose unmasked entries are fewer than k, where the masked
    (+inf) entries are handed out in index order."""
    h = sigs[:n_db]
    mask = 20 if n_db > 100 else 5
    row0 = 1000
    q0 = row0 + (17 if n_db > 100 else 2)        # query i is global row q0 + i
    q = sigs[:40].copy() if n_db > 100 else sigs[:8].copy()
    q[5] = 0.0                                   # zero-norm query: NaN distances
    db = api.SignatureDB("sc", h, global_row0=row0)
    db.match(q)
    st = db.partial_stats()
    idx, score, dpa, dia = db.topk(st, n_db, q0, mask, 2.0, k)
    dp, di = db.distances()
    db.close()
    dp, di = dp.astype(np.float64), di.astype(np.float64)
    for i, (order, f) in enumerate(_topk_numpy(dp, di, q0, row0, mask, k)):
        want = [row0 + j for j in order] + [-1] * (k - len(order))
        assert list(idx[i]) == want, (i, idx[i], want)
        for r, j in enumerate(order):
            if np.isfinite(f[j]):
                assert abs(score[i, r] - f[j]) < 1e-9 * (1 + abs(f[j]))
            np.testing.assert_array_equal([dpa[i, r], dia[i, r]], [dp[i, j], di[i, j]])   # NaN == NaN here
    # the zero-norm query has NaN everywhere; only masked entries (+inf by run_test.m:47-53) can be handed out
    sel = idx[5][idx[5] >= 0]
    assert (np.abs((q0 + 5) - sel) < mask).all()


def test_zero_norm_db_row_does_not_poison_fusion(gpu_ctx, oracle, sigs):
    """One degenerate (all-zero, e.g. empty scan) DB signature gives a NaN column (processSC.m:15-20).  MATLAB's
    normalize (run_test.m:40) omits NaN from the row statistics, so every other candidate is still ranked: same top-1
    as the oracle, never the NaN column, and -- on the fused kernels and on sodso_fuse_top1 -- finite scores."""
    h = sigs[:300].copy()
    h[123] = 0.0
    q = sigs[300:340]
    idx, score = api.run_test("sc", q, h, 0)
    rp, ri = oracle.sc_match_numpy(q, h)
    assert np.isnan(rp[:, 123]).all()
    ridx, rscore = oracle.fuse_top1(rp, ri, 0)
    np.testing.assert_array_equal(idx, ridx)
    assert np.isfinite(score).all() and not (idx == 123).any()
    np.testing.assert_allclose(score, rscore, rtol=0, atol=2e-3)     # fp32 distances vs fp64: scores, not decisions
    dp, di = api.processSC(q, h)
    fidx, fscore = api.fuse_top1(dp, di, 0)
    np.testing.assert_array_equal(fidx, oracle.fuse_top1(dp, di, 0)[0])
    assert np.isfinite(fscore).all()
    # the same through the resident-database entry points (partial statistics carry the non-NaN counts)
    db = api.SignatureDB("sc", h)
    kidx, kscore, _, _ = db.query_sharded(q, 0, 0, 2.0, 2)
    st_m = db.partial_stats()
    db.close()
    np.testing.assert_array_equal(kidx[:, 0], ridx)
    assert (st_m[:, 2] == 299).all() and (st_m[:, 5] == 299).all()    # the NaN column is not counted
