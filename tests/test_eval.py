"""Evaluation (run_test.m:2-22, 56-85): the oracle's restatement on hand cases, the host-side sodso_pr_curve against it
(host code: runs without a GPU), and the committed KITTI seq06 fixture's self-consistency."""
import ctypes as C
import os

import numpy as np

from conftest import GOLDEN
from so_dso_place_recognition_b200 import _native as N


def _pr_curve(score, idx, gt1, gt2, loop_diff, n_loops):
    score = np.ascontiguousarray(score, dtype=np.float64)
    idx = np.ascontiguousarray(idx, dtype=np.int32)
    gt1 = np.ascontiguousarray(gt1, dtype=np.float64)
    gt2 = np.ascontiguousarray(gt2, dtype=np.float64)
    m = len(score)
    auc, tr, tc = C.c_double(0), C.c_double(0), C.c_int(0)
    rank, pr, rc = np.zeros(m, np.int32), np.zeros(m), np.zeros(m)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    N.check(N.lib().sodso_pr_curve(p(score), p(idx), p(gt1), m, p(gt2), len(gt2), float(loop_diff), int(n_loops),
                                   C.byref(auc), C.byref(tr), C.byref(tc), p(rank), p(pr), p(rc)))
    return dict(AUC=auc.value, top_recall=tr.value, top_count=tc.value, rank=rank, precision=pr, recall=rc)


def test_gt_loops_hand_case(oracle):
    # a straight line walked forth and back: position i revisits position 19 - i
    x = np.concatenate([np.arange(10.0), np.arange(9.0, -1.0, -1.0)])
    gt = np.stack([x, np.zeros(20), np.zeros(20)], axis=1)
    lp, total = oracle.gt_loops(gt, gt, 0.5, 3)
    assert total == len(lp) == 18            # positions 9 and 10 only have each other, inside the mask
    assert all(j == 19 - i for i, j in lp) and {i for i, _ in lp} == set(range(0, 9)) | set(range(11, 20))
    # exactly one loop: MATLAB's length() of a 1 x 2 matrix is 2 (run_test.m:22)
    gt1 = np.array([[0.0, 0, 0], [100, 0, 0], [200, 0, 0]])
    gt2 = np.array([[50.0, 0, 0], [300, 0, 0], [200.2, 0, 0]])
    lp, total = oracle.gt_loops(gt1, gt2, 1.0, 0)
    assert lp.tolist() == [[2, 2]] and total == 2
    lp, total = oracle.gt_loops(gt1, gt2, 0.1, 0)
    assert lp.shape == (0, 2) and total == 0


def test_pr_eval_hand_case(oracle):
    gt = np.stack([np.arange(6.0) * 10, np.zeros(6), np.zeros(6)], axis=1)
    gt2 = gt + np.array([0.5, 0, 0])
    score = np.array([0.3, np.nan, 0.1, 0.1, np.inf, 0.2])
    idx = np.array([0, 1, 2, 0, 4, 5], dtype=np.int32)      # query 3 points to the wrong place
    ev = oracle.pr_eval(score, idx, gt, gt2, 5, 1.0)
    assert ev["rank"].tolist() == [2, 3, 5, 0, 4, 1]          # stable ties, Inf before NaN (MATLAB sort)
    np.testing.assert_allclose(ev["precision"], [1, 1 / 2, 2 / 3, 3 / 4, 4 / 5, 5 / 6])
    np.testing.assert_allclose(ev["recall"], [1 / 5, 1 / 5, 2 / 5, 3 / 5, 4 / 5, 1])
    assert ev["top_count"] == 1 and ev["top_recall"] == 0.2
    np.testing.assert_allclose(ev["AUC"], np.trapezoid(ev["precision"], ev["recall"]))


def test_host_pr_curve_matches_oracle(oracle):
    rng = np.random.default_rng(2)
    for m, n, nl in ((1, 1, 0), (7, 9, 1), (500, 400, 37), (300, 300, 300)):
        gt1, gt2 = rng.uniform(0, 50, (m, 3)), rng.uniform(0, 50, (n, 3))
        score = rng.normal(size=m).round(1)                   # many ties
        score[rng.random(m) < 0.05] = np.nan
        score[rng.random(m) < 0.05] = np.inf
        idx = rng.integers(0, n, m).astype(np.int32)
        total = 0 if nl == 0 else max(nl, 2)
        with np.errstate(all="ignore"):
            ref = oracle.pr_eval(score, idx, gt1, gt2, total, 8.0)
        got = _pr_curve(score, idx, gt1, gt2, 8.0, nl)
        np.testing.assert_array_equal(got["rank"], ref["rank"])
        np.testing.assert_array_equal(got["precision"], ref["precision"])
        np.testing.assert_array_equal(got["recall"], ref["recall"])
        assert got["top_count"] == ref["top_count"]
        np.testing.assert_allclose(got["top_recall"], ref["top_recall"], rtol=0, atol=0, equal_nan=True)
        np.testing.assert_allclose(got["AUC"], ref["AUC"], rtol=1e-12, equal_nan=True)


def test_seq06_fixture_self_consistent(oracle):
    g = np.load(os.path.join(GOLDEN, "seq06_sc_eval.npz"))
    lp, total = oracle.gt_loops(g["gt"], g["gt"], 10.0, 100)
    assert lp.shape[0] == int(g["n_gt_loops"]) == 377
    ev = oracle.pr_eval(g["score"], g["idx"], g["gt"], g["gt"], total, 10.0)
    assert ev["AUC"] == float(g["AUC"]) and ev["top_recall"] == float(g["top_recall"])
    assert 0.85 < ev["AUC"] < 0.95 and 0.6 < ev["top_recall"] < 0.75      # the ballpark of SURVEY §4
    got = _pr_curve(g["score"], g["idx"], g["gt"], g["gt"], 10.0, lp.shape[0])
    assert got["top_count"] == int(g["top_count"]) and abs(got["AUC"] - float(g["AUC"])) < 1e-12
