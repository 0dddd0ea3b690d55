"""GPU parity of DELIGHT: generation (DELIGHT.cpp:6-24, test_delight.cpp:38-56), the chi-square matcher
(processDELIGHT.m:1-38) and the single-matrix decision (run_test.m:47-57) against the CPU oracle."""
import numpy as np
import pytest

from so_dso_place_recognition_b200 import api, synth

pytestmark = pytest.mark.gpu


def test_delight_generate_vs_oracle(gpu_ctx, oracle, real_scans):
    xyz, inten, off = synth.make_scan_set(64, 2048, planted_loops=True)
    h = api.delight_generate(xyz, inten, off)
    ref = oracle.delight_generate(xyz, inten, off, nthreads=8)
    assert h.shape == (16 * 64, 256) and gpu_ctx.last_kernel_name == "delight_generate_kernel"
    np.testing.assert_array_equal(h, ref)                     # integer counts
    assert (h.reshape(64, -1).sum(axis=1) == 2048).all()
    # real SO-DSO scans (polar-filtered, non-dyadic intensities), ragged, an empty scan, out-of-range intensities
    h = api.delight_generate(real_scans["m2dp_xyz"], real_scans["m2dp_inten"], real_scans["m2dp_off"])
    np.testing.assert_array_equal(h, oracle.delight_generate(real_scans["m2dp_xyz"], real_scans["m2dp_inten"], real_scans["m2dp_off"]))
    sizes = [0, 1, 5, 700]
    parts = [synth.make_scan(900 + k, max(n, 1)) for k, n in enumerate(sizes)]
    x = np.concatenate([p[0][:n] for p, n in zip(parts, sizes)])
    it = np.concatenate([p[1][:n] for p, n in zip(parts, sizes)])
    it[3] = 255.99; it[4] = 256.0; it[5] = -0.5; it[6] = -1.0; it[7] = np.nan; it[8] = 1e9
    o = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    with np.errstate(all="ignore"):
        ref = oracle.delight_generate(x, it, o)
    np.testing.assert_array_equal(api.delight_generate(x, it, o), ref)
    assert not ref[:16].any()


def test_delight_match_and_top1(gpu_ctx, oracle):
    xyz, inten, off = synth.make_scan_set(80, 2048, planted_loops=True)
    h = oracle.delight_generate(xyz, inten, off, nthreads=8)
    d = api.processDELIGHT(h[:16 * 23], h)
    ref = oracle.delight_match(h[:16 * 23], h, nthreads=8)
    assert d.shape == (23, 80) and gpu_ctx.last_kernel_name == "delight_match_kernel"
    np.testing.assert_allclose(d, ref, rtol=1e-5, atol=1e-9)
    assert np.abs(np.diag(d)).max() == 0.0
    # empty signatures: 0/0 -> NaN for every permutation -> Inf (processDELIGHT.m:16,31-34)
    hz = h.copy()
    hz[16 * 5:16 * 6] = 0
    dz = api.processDELIGHT(hz[16 * 5:16 * 6], hz[16 * 5:16 * 6])
    assert np.isinf(dz[0, 0]) and np.isinf(oracle.delight_match(hz[16 * 5:16 * 6], hz[16 * 5:16 * 6])[0, 0])
    # decision
    dfull = oracle.delight_match(h, h, nthreads=8)
    ridx, rscore = oracle.top1_single(dfull, 5)
    idx, score = api.run_test("delight", h, h, 5)
    np.testing.assert_array_equal(idx, ridx)
    np.testing.assert_allclose(score, rscore, rtol=1e-5)
    assert (idx == (np.arange(80) + 40) % 80).mean() > 0.6     # DELIGHT is the weakest of the three descriptors
    # masks and NaN rows of the single-matrix decision
    dd = dfull.copy()
    dd[3] = np.nan
    dd[4, 10] = -1.0
    for mw in (0, 3, 200):
        i1, s1 = api.top1_single(dd, mw)
        i2, s2 = oracle.top1_single(dd, mw)
        np.testing.assert_array_equal(i1, i2)
        np.testing.assert_array_equal(s1, s2)
