"""Eigen's implementation-defined signs: the oracle and the GPU fix a convention for the PCA
eigenvectors (pts_align.h:31-39: largest-magnitude component positive) and for the dominant singular pair
(M2DP.cpp:94-103: sum(U1) >= 0).  The real reference inherits whatever its Eigen build returns, which cannot be
observed here.  tests/golden/make_golden_real.py flips OUR convention per scan on KITTI seq06 and records what changes
(tests/golden/sign_flip_experiment.json) -- the only available bound on divergence from a real Eigen build.  This
test checks the committed record and, where the reference's data files are present (this container), recomputes the
key cases with fresh random flips."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN

RES = "/root/reference/place_recognition/results/KITTI/seq06/"


@pytest.fixture(scope="module")
def record():
    return json.load(open(os.path.join(GOLDEN, "sign_flip_experiment.json")))


def test_committed_record(record):
    sc, m2 = record["scan_context"], record["m2dp"]
    base = sc["baseline"]
    # the up axis (x = least variance) only enters through max - min of the height: its sign changes NOTHING
    assert sc["v0_up_axis"]["signature_rows_changed_frac"] == 0.0 and sc["v0_up_axis"]["top1_changed_frac"] == 0.0
    # the in-plane axes mirror the polar image: a sector reversal + shift, covered by the 120 variants of
    # processSC.m:24-28 EXCEPT for the ring-aliased points (SC.cpp:37-44 lands a point with ri >= 20 in sector si + 1,
    # which is a different physical neighbour in the mirrored frame).  Decisions of near-tie queries move, quality
    # does not: AUC to 1e-3, top recall to 0.01.
    for name in ("v1_middle_axis", "v2_major_axis", "all_three"):
        r = sc[name]
        assert 0.0 < r["top1_changed_frac"] < 0.2 and r["top1_changed_frac_gt_loop_queries"] < 0.08
        assert abs(r["AUC"] - base["AUC"]) < 1e-3 and abs(r["top_recall"] - base["top_recall"]) < 0.01
    # M2DP: the 4 variants of test_m2dp.cpp:47-57 are the PROPER sign patterns only (dx, dy, dx*dy); a single flipped
    # eigenvector is a reflection and is not covered: distances move by up to ~0.06 and many near-tie decisions with
    # them, quality by < 0.01 AUC
    r = m2["pca_all_three_random"]
    assert r["max_abs_distance_change"] < 0.1 and abs(r["AUC"] - m2["baseline"]["AUC"]) < 0.01
    # the SVD pair sign: flipping it for ALL scans changes nothing at all (both factors of every dot product flip) ...
    r = m2["svd_all_scans_flipped"]
    assert r["top1_changed_frac"] == 0.0 and r["max_abs_distance_change"] == 0.0
    # ... while scan-dependent signs destroy the descriptor (d jumps between ~0 and ~1, processM2DP.m:15 has no
    # sign compensation): a reference that recognises places at all has scan-independent signs, i.e. our Perron
    # convention up to the global flip above
    r = m2["svd_random_per_scan"]
    assert r["AUC"] < 0.65 and r["top_recall"] < 0.05 and r["max_abs_distance_change"] > 1.5


@pytest.mark.skipif(not os.path.exists(RES + "pts_history_file.txt"), reason="needs the reference's data files")
def test_recompute_on_seq06(oracle, record):
    st = oracle.stage(RES + "poses_history_file.txt", RES + "pts_history_file.txt", 45.0, False)
    gt = np.loadtxt(RES + "gt.txt")[st["ids"], :][:, [3, 7, 11]]
    ns = len(st["ids"])
    assert ns == record["scan_context"]["n_scans"]
    lp, total_lp = oracle.gt_loops(gt, gt, 10.0, 100)
    rng = np.random.default_rng(5)

    def run(flip):
        h = oracle.sc_generate_flip(st["xyz"], st["inten"], st["off"], flip, nthreads=8)
        dp, di = oracle.sc_match_numpy(h, h)
        idx, score = oracle.fuse_top1(dp, di, 100)
        return h, dp, idx, oracle.pr_eval(score, idx, gt, gt, total_lp, 10.0)

    h0, dp0, idx0, ev0 = run(np.zeros(ns, dtype=np.int32))
    assert abs(ev0["AUC"] - record["scan_context"]["baseline"]["AUC"]) < 1e-12
    np.testing.assert_array_equal(h0, oracle.sc_generate(st["xyz"], st["inten"], st["off"], nthreads=8))
    hu, _, idxu, _ = run(rng.integers(0, 2, ns).astype(np.int32))            # up axis
    assert np.array_equal(hu, h0) and np.array_equal(idxu, idx0)
    hf, dpf, idxf, evf = run(rng.integers(0, 8, ns).astype(np.int32))        # all three, at random per scan
    changed = idxf != idx0
    assert changed.mean() < 0.2 and abs(evf["AUC"] - ev0["AUC"]) < 1e-3 and abs(evf["top_recall"] - ev0["top_recall"]) < 0.01
    # the distances themselves move only through the aliased points
    assert np.abs(dpf - dp0).max() < 0.05 and np.median(np.abs(dpf - dp0)) < 1e-3
    # M2DP on a slice: global SVD flip = identical signatures up to sign and identical distances
    stp = oracle.stage(RES + "poses_history_file.txt", RES + "pts_history_file.txt", 45.0, True)
    k = 40
    off = stp["off"][:k + 1]
    xyz, inten = stp["xyz"][:off[-1]], stp["inten"][:off[-1]]
    z = np.zeros(k, dtype=np.int32)
    a = oracle.m2dp_generate_flip(xyz, inten, off, z, z, nthreads=8)
    b = oracle.m2dp_generate_flip(xyz, inten, off, z, np.full(k, 3, dtype=np.int32), nthreads=8)
    np.testing.assert_array_equal(a, oracle.m2dp_generate(xyz, inten, off, nthreads=8))
    np.testing.assert_array_equal(a, -b)
    da, _ = oracle.m2dp_match(a, a, nthreads=8)
    db, _ = oracle.m2dp_match(b, b, nthreads=8)
    np.testing.assert_array_equal(da, db)
