"""End-to-end through the C++ host drivers on a synthetic SO-DSO sequence written in the reference's
file formats (PosesPts.h:12-24,35-39): files -> gen_signatures -> history_sc.txt / history_m2dp.txt ->
match_signatures -> loop candidates, checked against the oracle run on the same files."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu
HOST = os.path.join(ROOT, "so_dso_place_recognition_b200", "host")
BIN = os.path.join(ROOT, "so_dso_place_recognition_b200", "bin")


def _write_sequence(d, n_pose=46, pts_per_frame=260, seed=3):
    rng = np.random.default_rng(seed)
    poses, pts = [], []
    for i in range(n_pose):
        z = 2.0 + 0.8 * i                      # camera moves forward; |t| >= 1 so no reset is triggered
        yaw = 0.02 * i
        c, s = np.cos(yaw), np.sin(yaw)
        R = np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])
        cam = np.array([0.3 * i, 0.0, z])
        t = -R @ cam
        poses.append((2 * i, np.hstack([R, t[:, None]])))
        p = cam + np.stack([rng.uniform(-30, 30, pts_per_frame), np.clip(rng.normal(0, 1.5, pts_per_frame), -5, 5),
                            rng.uniform(-10, 45, pts_per_frame)], axis=1)
        it = rng.uniform(5, 250, pts_per_frame)
        pts += [(2 * i, q, v) for q, v in zip(p, it)]
    with open(d / "poses.txt", "w") as f:
        for i, w in poses:
            f.write(f"{i} " + " ".join("%g" % v for v in w.reshape(-1)) + " \n")
    with open(d / "pts.txt", "w") as f:
        for i, q, v in pts:
            f.write(f"{i} {q[0]:g} {q[1]:g} {q[2]:g} {v:g}\n")
    return str(d / "poses.txt"), str(d / "pts.txt")


def test_files_to_loops(gpu_ctx, oracle, tmp_path):
    subprocess.check_call(["make", "-C", HOST, "-s"])
    gen, match = os.path.join(BIN, "gen_signatures"), os.path.join(BIN, "match_signatures")
    poses, pts = _write_sequence(tmp_path)
    for kind, polar, width in (("sc", False, 2400), ("m2dp", True, 384)):
        hist_file, ids_file = str(tmp_path / f"history_{kind}.txt"), str(tmp_path / "incoming_id_file.txt")
        subprocess.check_call([gen, kind, poses, pts, hist_file, ids_file, "45"], stdout=subprocess.DEVNULL)
        st = oracle.stage(poses, pts, 45.0, polar)
        np.testing.assert_array_equal(np.loadtxt(ids_file, dtype=np.int64), st["ids"])
        n = len(st["ids"])
        assert n == 16
        ref = (oracle.sc_generate if kind == "sc" else oracle.m2dp_generate)(st["xyz"], st["inten"], st["off"])
        hist = np.loadtxt(hist_file)
        assert hist.shape == ((n if kind == "sc" else 4 * n), width)
        np.testing.assert_allclose(hist, ref, rtol=1e-5, atol=1e-9)      # 6 significant digits in the text file
        # match on the text round-tripped signatures (what MATLAB would load)
        loops = str(tmp_path / "loops.txt")
        subprocess.check_call([match, kind, hist_file, hist_file, "3", loops], stdout=subprocess.DEVNULL)
        got = np.loadtxt(loops)
        dp, di = (oracle.sc_match_numpy if kind == "sc" else oracle.m2dp_match)(hist, hist)
        ridx, rscore = oracle.fuse_top1(dp, di, 3)
        np.testing.assert_array_equal(got[:, 0].astype(int) - 1, ridx)     # file holds MATLAB's 1-based index
        # fused score = 2 (d_p - mu_p)/sigma_p + (d_i - mu_i)/sigma_i (run_test.m:38-41): the 1e-5 distance bar
        # is amplified by 1/sigma of the row (16 near-identical frames => small sigma)
        sp, si_ = np.nanstd(dp, axis=1, ddof=1), np.nanstd(di, axis=1, ddof=1)
        tol = 1e-6 + 4.0 * 1e-5 * (2.0 / sp + 1.0 / si_)
        err = np.abs(got[:, 1] - rscore)
        assert np.all(err <= tol), (kind, err.max(), tol.min(), sp.min(), si_.min())
