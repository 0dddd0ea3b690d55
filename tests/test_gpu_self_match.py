"""Self-match symmetry of the tcgen05 Scan Context matcher (so_dso_place_recognition_b200/csrc/sc_match_tc.cu,
`launch_sc_match_tc_self`): when hist1 and hist2 are the same n rows the library computes the lower block triangle of
(query group, DB tile) items and stores each value at its transposed position too.  Checked against the oracle
(processSC.m:1-45), against the full computation (test hook sodso_debug_set_sc_symmetry) and for exact symmetry of the mirrored part."""
import numpy as np
import pytest

from so_dso_place_recognition_b200 import api, synth

pytestmark = pytest.mark.gpu


def _sigs(n, seed=0):
    xyz, inten, off = synth.make_scan_set(n, 1024, planted_loops=True, first=seed)
    return api.sc_generate(xyz, inten, off)


@pytest.mark.parametrize("n", [3, 255, 256, 257, 515, 777, 1024, 1301])
def test_self_match_vs_oracle(gpu_ctx, oracle, n):
    h = _sigs(n, seed=n)
    dp, di = api.processSC(h, h)
    rp, ri = oracle.sc_match(h, h, nthreads=16)
    np.testing.assert_allclose(dp, rp, rtol=0, atol=1e-5)
    np.testing.assert_allclose(di, ri, rtol=0, atol=1e-5)


def test_self_match_equals_full_computation(gpu_ctx):
    h = _sigs(1301, seed=7)
    dp, di = api.processSC(h, h)
    gpu_ctx.set_sc_symmetry(False)
    try:
        fp, fi = api.processSC(h, h)
    finally:
        gpu_ctx.set_sc_symmetry(True)
    # the directly computed part is bit-identical, the mirrored part differs by fp32 summation order only
    np.testing.assert_allclose(dp, fp, rtol=0, atol=2e-6)
    np.testing.assert_allclose(di, fi, rtol=0, atol=2e-6)
    low = np.tril_indices(1301)
    blk_same = (low[0] // 4 * 4 + 3) >= (low[1] // 256 * 256)      # items of the lower block triangle
    assert np.array_equal(dp[low][blk_same], fp[low][blk_same])
    assert np.array_equal(di[low][blk_same], fi[low][blk_same])
    # mirrored entries are exact copies
    i, j = np.meshgrid(np.arange(1301), np.arange(1301), indexing="ij")
    mirrored = (j // 256 * 256) > (i // 4 * 4 + 3)
    assert np.array_equal(dp[mirrored], dp.T[mirrored]) and np.array_equal(di[mirrored], di.T[mirrored])


def test_same_values_different_buffers_take_the_general_path(gpu_ctx, oracle):
    """Symmetry is keyed on hist1 and hist2 being the SAME buffer; equal contents in two buffers are matched pair by pair."""
    h = _sigs(300, seed=3)
    dp, di = api.processSC(h, h.copy())
    rp, ri = oracle.sc_match(h, h, nthreads=8)
    np.testing.assert_allclose(dp, rp, rtol=0, atol=1e-5)
    np.testing.assert_allclose(di, ri, rtol=0, atol=1e-5)


def test_top1_identical_with_and_without_symmetry(gpu_ctx):
    h = _sigs(2000, seed=11)
    idx, score = api.run_test("sc", h, h, 100)
    gpu_ctx.set_sc_symmetry(False)
    try:
        idx0, score0 = api.run_test("sc", h, h, 100)
    finally:
        gpu_ctx.set_sc_symmetry(True)
    assert np.array_equal(idx, idx0)
    np.testing.assert_allclose(score, score0, rtol=0, atol=1e-3)


def test_self_match_into_misaligned_output(gpu_ctx, oracle):
    """The transposed values are written with 16-byte stores only where the address allows it: an fp32 output matrix
    that starts 4 bytes into an allocation (a caller's view) takes the scalar path."""
    import ctypes as C

    import torch

    from so_dso_place_recognition_b200 import _native as N

    n = 516
    h = torch.from_numpy(_sigs(n, seed=21)).cuda()
    buf_p = torch.zeros(n * n + 1, dtype=torch.float32, device="cuda")
    buf_i = torch.zeros(n * n + 1, dtype=torch.float32, device="cuda")
    dp, di = buf_p[1:], buf_i[1:]
    assert dp.data_ptr() % 16 == 4
    N.check(N.lib().sodso_sc_match_f32(gpu_ctx.handle, C.c_void_p(h.data_ptr()), n, C.c_void_p(h.data_ptr()), n,
                                       C.c_void_p(dp.data_ptr()), C.c_void_p(di.data_ptr())))
    torch.cuda.synchronize()
    rp, ri = oracle.sc_match(h.cpu().numpy(), h.cpu().numpy(), nthreads=16)
    np.testing.assert_allclose(dp.view(n, n).cpu().numpy(), rp, rtol=0, atol=1e-5)
    np.testing.assert_allclose(di.view(n, n).cpu().numpy(), ri, rtol=0, atol=1e-5)
