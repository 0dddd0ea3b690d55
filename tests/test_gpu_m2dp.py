"""GPU parity: M2DP generation (test_m2dp.cpp:41-67, M2DP.cpp:38-109) and matching
(processM2DP.m:1-22) through the C ABI vs the CPU oracle.

Bar: signatures within 1e-9 of the oracle (dominant singular pair, oracle sign convention
sum(U1) >= 0); distances within 1e-5; identical top-1 (BASELINE.json north_star)."""
import numpy as np
import pytest

from so_dso_place_recognition_b200 import api, synth

pytestmark = pytest.mark.gpu
TOL_SIG = 1e-9
TOL_D = 1e-5


def test_m2dp_generate_vs_oracle(gpu_ctx, oracle):
    xyz, inten, off = synth.make_scan_set(12, 2048)
    h = api.m2dp_generate(xyz, inten, off)
    ref = oracle.m2dp_generate(xyz, inten, off, nthreads=8)
    assert h.shape == (48, 384)
    np.testing.assert_allclose(h, ref, rtol=0, atol=TOL_SIG)
    # unit-norm singular vectors, non-negative orientation
    assert np.allclose(np.linalg.norm(h[:, :64], axis=1), 1) and np.allclose(np.linalg.norm(h[:, 64:192], axis=1), 1)
    assert (h[:, :64].sum(1) >= 0).all()


def test_m2dp_class_contract_prealigned(gpu_ctx, oracle):
    """M2DP::getSignature takes already aligned points (M2DP.h:18-20)."""
    xyz, inten = synth.make_scan(21, 3000)
    al, _, _ = oracle.align_pca(xyz)
    c_ref, i_ref = oracle.m2dp_signature(al, inten)
    m = api.M2DP(45.0)
    assert m.getSignatureSize() == 192
    c, i = m.getSignature(al, inten)
    np.testing.assert_allclose(c, c_ref, rtol=0, atol=TOL_SIG)
    np.testing.assert_allclose(i, i_ref, rtol=0, atol=TOL_SIG)


def test_m2dp_real_scans_and_ragged(gpu_ctx, oracle, real_scans):
    h = api.m2dp_generate(real_scans["m2dp_xyz"], real_scans["m2dp_inten"], real_scans["m2dp_off"])
    np.testing.assert_allclose(h, real_scans["m2dp_hist"], rtol=0, atol=TOL_SIG)
    sizes = [5, 300, 4096, 4500]
    parts = [synth.make_scan(300 + k, n) for k, n in enumerate(sizes)]
    xyz = np.concatenate([p[0] for p in parts])
    inten = np.concatenate([p[1] for p in parts])
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    np.testing.assert_allclose(api.m2dp_generate(xyz, inten, off), oracle.m2dp_generate(xyz, inten, off, nthreads=4),
                               rtol=0, atol=TOL_SIG)


@pytest.mark.parametrize("algo", [api.SODSO_ALGO_TC, api.SODSO_ALGO_SIMT])
def test_m2dp_match_and_top1(gpu_ctx, oracle, algo):
    """tcgen05 matcher (product path) and the fp32 CUDA-core cross-check against the oracle"""
    gpu_ctx.set_match_algo(algo)
    xyz, inten, off = synth.make_scan_set(96, 1024, planted_loops=True)
    sig = oracle.m2dp_generate(xyz, inten, off, nthreads=16)
    dp, di = api.processM2DP(sig[:4 * 37], sig)
    assert gpu_ctx.last_kernel_name == ("m2dp_match_tc_kernel" if algo == api.SODSO_ALGO_TC else "m2dp_match_kernel")
    rp, ri = oracle.m2dp_match(sig[:4 * 37], sig, nthreads=8)
    assert dp.shape == (37, 96)
    assert np.abs(dp - rp).max() < TOL_D and np.abs(di - ri).max() < TOL_D
    rp, ri = oracle.m2dp_match(sig, sig, nthreads=8)
    ridx, rsc = oracle.fuse_top1(rp, ri, 5)
    idx, sc = api.run_test("m2dp", sig, sig, 5)
    np.testing.assert_array_equal(idx, ridx)
    np.testing.assert_allclose(sc, rsc, atol=2e-3)
    gpu_ctx.set_match_algo(api.SODSO_ALGO_TC)


@pytest.mark.parametrize("m,n", [(1, 1), (3, 70), (65, 64), (130, 257)])
def test_m2dp_match_ragged_shapes_tc_vs_simt(gpu_ctx, m, n):
    """tile edges of the tensor-core matcher (256 variant rows = 64 scans per tile) against the fp32 kernel"""
    rng = np.random.default_rng(m * 1000 + n)

    def sigs(k):     # rows like real signatures: non-negative unit vectors [u(64) v(128)] per channel
        a = rng.random((4 * k, 2, 192))
        a[..., :64] /= np.linalg.norm(a[..., :64], axis=-1, keepdims=True)
        a[..., 64:] /= np.linalg.norm(a[..., 64:], axis=-1, keepdims=True)
        return a.reshape(4 * k, 384)

    h1, h2 = sigs(m), sigs(n)
    gpu_ctx.set_match_algo(api.SODSO_ALGO_TC)
    dp, di = api.processM2DP(h1, h2)
    gpu_ctx.set_match_algo(api.SODSO_ALGO_SIMT)
    sp, si = api.processM2DP(h1, h2)
    gpu_ctx.set_match_algo(api.SODSO_ALGO_TC)
    assert dp.shape == (m, n)
    ref = (1.0 - np.einsum("ivk,jwk->ijvw", h1[:, :192].reshape(m, 4, 192), h2[:, :192].reshape(n, 4, 192))) / 2
    assert np.abs(dp - ref.min(axis=(2, 3))).max() < TOL_D                  # processM2DP.m:15-21 in numpy
    assert np.abs(dp - sp).max() < TOL_D and np.abs(di - si).max() < TOL_D


def test_m2dp_end_to_end_planted_loops(gpu_ctx):
    """BASELINE configs[4] in small: GPU generation -> GPU match -> loops recovered."""
    n = 64
    xyz, inten, off = synth.make_scan_set(n, 2048, planted_loops=True)
    sig = api.m2dp_generate(xyz, inten, off)
    idx, sc = api.run_test("m2dp", sig, sig, 3)
    assert (idx == (np.arange(n) + n // 2) % n).mean() > 0.8


def test_m2dp_points_in_the_guard_band(gpu_ctx, oracle):
    """The fp32 bin proposal is only trusted away from bin edges; everything closer takes M2DP.cpp:56-63 in fp64.
    Pre-aligned points (class contract, identity transform) placed within 1e-12 .. 1e-4 of sector / ring edges of the
    x-z projection plane (p=0, q=0), plus points on the axes, at the origin and just inside / outside max_rho: the
    histograms must come out exactly as the oracle's, which shows in the singular vectors at 1e-9."""
    rng = np.random.default_rng(17)
    n = 4000
    k = rng.integers(0, 16, n)
    j = rng.integers(1, 8, n)
    eps_t = rng.choice([-1, 1], n) * 10.0 ** rng.uniform(-12, -4, n)
    eps_r = rng.choice([-1, 1], n) * 10.0 ** rng.uniform(-12, -4, n)
    theta = -np.pi + k * (2 * np.pi / 16) + eps_t * (rng.random(n) < 0.7)
    r = np.where(rng.random(n) < 0.5, j * 45.0 / 8 + eps_r, rng.uniform(0.2, 44, n))
    pts = np.stack([r * np.cos(theta), rng.normal(0, 1.0, n), r * np.sin(theta)], axis=1)
    special = np.array([[0, 0, 0], [1, 0, 0], [-1, 0, 0], [0, 0, 1], [0, 0, -1], [0, 3, 0], [0, -3, 0], [-2, -2, -2],
                        [44.9999999, 0, 0], [45.0000001, 0, 0], [1e-300, 0, 1e-300], [-0.0, 0.0, -0.0]], dtype=float)
    pts = np.concatenate([pts, special])
    inten = (rng.integers(0, 2041, len(pts)) / 8.0).astype(np.float32)
    c_ref, i_ref = oracle.m2dp_signature(pts, inten)
    c, i = api.M2DP(45.0).getSignature(pts, inten)
    np.testing.assert_allclose(c, c_ref, rtol=0, atol=TOL_SIG)
    np.testing.assert_allclose(i, i_ref, rtol=0, atol=TOL_SIG)


def _symmetric_cloud(rng, n_half, scale):
    """points in +- pairs: the centroid is the origin to rounding, the principal axes are the coordinate axes"""
    p = rng.normal(size=(n_half, 3)) * np.asarray(scale, dtype=np.float64)
    return np.concatenate([p, -p])


@pytest.mark.parametrize("n_centre", [40, 60, 600])
def test_m2dp_variant_sharing_with_a_full_guard_band_queue(gpu_ctx, oracle, n_centre):
    """The four variants share plane evaluations (pairs (0, 2), (1, 3) exactly; the p = 0 planes across the pairs to
    within the guard band), and evaluations inside the guard band are queued and replayed in fp64 per variant.  Points
    within 1e-5 m of the centroid are inside the guard band of EVERY plane: 40 of them nearly fill the queue (3 200 of
    4 096 entries, all with twin and cross-pair duties), 60 overflow it in the pass over the second variant's p = 0
    planes only (those entries are served on the spot and the second pair may not be seeded from the first), 600
    overflow it in the main pass (the pair is redone variant by variant).  Either way: the oracle's signatures."""
    rng = np.random.default_rng(100 + n_centre)
    cloud = _symmetric_cloud(rng, 1800, (2.0, 9.0, 21.0))
    centre = _symmetric_cloud(rng, n_centre // 2, (1e-5, 1e-5, 1e-5))
    xyz = np.concatenate([cloud, centre])
    xyz = xyz[rng.permutation(len(xyz))] + np.array([3.0, -7.0, 11.0])
    inten = rng.integers(0, 256, len(xyz)).astype(np.float32)
    off = np.array([0, len(xyz)], dtype=np.int64)
    ref = oracle.m2dp_generate(xyz, inten, off)
    np.testing.assert_allclose(api.m2dp_generate(xyz, inten, off), ref, rtol=0, atol=TOL_SIG)
    # the same scan between ordinary ones (the per-CTA stash and the queue are reused from scan to scan)
    a, ai, _ = synth.make_scan_set(1, 2048)
    xyz3 = np.concatenate([a, xyz, a])
    inten3 = np.concatenate([ai, inten, ai])
    off3 = np.array([0, len(a), len(a) + len(xyz), 2 * len(a) + len(xyz)], dtype=np.int64)
    h3 = api.m2dp_generate(xyz3, inten3, off3)
    np.testing.assert_allclose(h3[4:8], ref, rtol=0, atol=TOL_SIG)
    np.testing.assert_array_equal(h3[0:4], h3[8:12])


def test_m2dp_far_points_and_large_range(gpu_ctx, oracle):
    """lidarRange 200 m: points beyond 64 m (no cross-pair sharing for them) and beyond 128 m (no fp32 proposal at all)"""
    rng = np.random.default_rng(7)
    xyz = _symmetric_cloud(rng, 1500, (6.0, 35.0, 70.0)) + np.array([1.0, 2.0, 3.0])
    assert (np.abs(xyz).max(axis=1) > 64).sum() > 300 and (np.abs(xyz).max(axis=1) > 128).sum() > 20
    inten = rng.integers(0, 256, len(xyz)).astype(np.float32)
    off = np.array([0, len(xyz)], dtype=np.int64)
    np.testing.assert_allclose(api.m2dp_generate(xyz, inten, off, 200.0), oracle.m2dp_generate(xyz, inten, off, 200.0),
                               rtol=0, atol=TOL_SIG)


def test_m2dp_large_scan_64bit_gram(gpu_ctx, oracle):
    """70 000 points: Gram entries can exceed 2^32 (64-bit integer path), bin counts exceed 2^12"""
    rng = np.random.default_rng(8)
    xyz = rng.normal(size=(70000, 3)) * np.array([1.5, 6.0, 14.0])
    inten = rng.integers(0, 256, len(xyz)).astype(np.float32)
    off = np.array([0, len(xyz)], dtype=np.int64)
    np.testing.assert_allclose(api.m2dp_generate(xyz, inten, off), oracle.m2dp_generate(xyz, inten, off, nthreads=4),
                               rtol=0, atol=TOL_SIG)


def test_m2dp_empty_and_tiny_scans(gpu_ctx, oracle):
    """0 points (zero matrices: u = e_0, v = 0), 1 point (it sits at the centroid: signed zeros through the sign flips
    of the four variants, M2DP.cpp:59 at (+-0, +-0)) and a 40-point scan.  Scans of 2 or 3 points are left out on
    purpose: their scatter matrix is rank deficient, the eigenvectors of the repeated zero eigenvalue are an
    arbitrary basis (pts_align.h:31-34 takes whatever Eigen returns), and the 1e-16 coordinates along them decide
    sectors -- the oracle and the kernels disagree there, and so would two Eigen versions."""
    rng = np.random.default_rng(5)
    sizes = [0, 1, 0, 40]
    xyz = rng.normal(size=(sum(sizes), 3)) * np.array([3.0, 8.0, 15.0])
    inten = rng.integers(0, 256, sum(sizes)).astype(np.float32)
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    h = api.m2dp_generate(xyz, inten, off)
    ref = oracle.m2dp_generate(xyz, inten, off)
    assert h.shape == (16, 384)
    np.testing.assert_allclose(h, ref, rtol=0, atol=TOL_SIG)
