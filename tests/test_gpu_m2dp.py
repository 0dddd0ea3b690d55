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
