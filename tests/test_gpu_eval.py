"""GPU: ground-truth loop search (run_test.m:3-22) vs the oracle, and the whole run_test (match, fuse, mask, argmin,
precision-recall) on the committed KITTI seq06 signatures: identical top-1 indices, AUC and top recall."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from so_dso_place_recognition_b200 import api

pytestmark = pytest.mark.gpu


def test_gt_loops_vs_oracle(gpu_ctx, oracle):
    rng = np.random.default_rng(4)
    t = np.linspace(0, 4 * np.pi, 700)
    gt1 = np.stack([40 * np.cos(t), rng.normal(0, 0.2, 700), 40 * np.sin(t)], axis=1)      # two laps of a circle
    gt2 = gt1 + rng.normal(0, 0.5, gt1.shape)
    for mask, ld in ((100, 5.0), (0, 1.0), (699, 5.0), (5000, 5.0)):
        lp, n = api.gt_loops(gt1, gt2, ld, mask)
        ref, total = oracle.gt_loops(gt1, gt2, ld, mask)
        np.testing.assert_array_equal(lp, ref)
        assert n == ref.shape[0]
    lp, n = api.gt_loops(gt1[:0], gt2, 5.0, 10)
    assert n == 0 and lp.shape == (0, 2)


def test_run_test_seq06(gpu_ctx):
    g = np.load(os.path.join(GOLDEN, "seq06_sc_eval.npz"))
    hist = np.concatenate([g["structure6"].astype(np.float64),
                           np.unpackbits(g["intensity_bits"], axis=1)[:, :1200].astype(np.float64)], axis=1)
    idx, score = api.run_test("sc", hist, hist, 100)
    np.testing.assert_array_equal(idx, g["idx"])                       # identical loop candidates on real data
    assert np.abs(score - g["score"]).max() < 1e-3
    auc, top_recall, lp_detected = api.run_test_full("sc", hist, hist, g["gt"], g["gt"], 10.0, 100)
    assert abs(auc - float(g["AUC"])) < 1e-9 and abs(top_recall - float(g["top_recall"])) < 1e-12
    assert lp_detected.shape == (int(g["top_count"]), 2)
    lp, n = api.gt_loops(g["gt"], g["gt"], 10.0, 100)
    np.testing.assert_array_equal(lp, g["lp_gt"])
