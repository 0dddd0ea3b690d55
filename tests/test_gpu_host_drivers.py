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
    for kind, polar, width in (("sc", False, 2400), ("m2dp", True, 384), ("delight", True, 256)):
        hist_file, ids_file = str(tmp_path / f"history_{kind}.txt"), str(tmp_path / "incoming_id_file.txt")
        subprocess.check_call([gen, kind, poses, pts, hist_file, ids_file, "45"], stdout=subprocess.DEVNULL)
        st = oracle.stage(poses, pts, 45.0, polar)
        np.testing.assert_array_equal(np.loadtxt(ids_file, dtype=np.int64), st["ids"])
        n = len(st["ids"])
        assert n == 16
        ogen = {"sc": oracle.sc_generate, "m2dp": oracle.m2dp_generate, "delight": oracle.delight_generate}[kind]
        ref = ogen(st["xyz"], st["inten"], st["off"])
        hist = np.loadtxt(hist_file)
        assert hist.shape == ({"sc": n, "m2dp": 4 * n, "delight": 16 * n}[kind], width)
        np.testing.assert_allclose(hist, ref, rtol=1e-5, atol=1e-9)      # 6 significant digits in the text file
        # match on the text round-tripped signatures (what MATLAB would load)
        loops = str(tmp_path / "loops.txt")
        subprocess.check_call([match, kind, hist_file, hist_file, "3", loops], stdout=subprocess.DEVNULL)
        got = np.loadtxt(loops)
        if kind == "delight":
            ridx, rscore = oracle.top1_single(oracle.delight_match(hist, hist), 3)
            np.testing.assert_array_equal(got[:, 0].astype(int) - 1, ridx)
            np.testing.assert_allclose(got[:, 1], rscore, rtol=1e-5)
            continue
        dp, di = (oracle.sc_match_numpy if kind == "sc" else oracle.m2dp_match)(hist, hist)
        ridx, rscore = oracle.fuse_top1(dp, di, 3)
        np.testing.assert_array_equal(got[:, 0].astype(int) - 1, ridx)     # file holds MATLAB's 1-based index
        # fused score = 2 (d_p - mu_p)/sigma_p + (d_i - mu_i)/sigma_i (run_test.m:38-41): the 1e-5 distance bar
        # is amplified by 1/sigma of the row (16 near-identical frames => small sigma)
        sp, si_ = np.nanstd(dp, axis=1, ddof=1), np.nanstd(di, axis=1, ddof=1)
        tol = 1e-6 + 4.0 * 1e-5 * (2.0 / sp + 1.0 / si_)
        err = np.abs(got[:, 1] - rscore)
        assert np.all(err <= tol), (kind, err.max(), tol.min(), sp.min(), si_.min())


def _write_two_laps(d, n_pose=170, seed=9):
    """two laps around a circle through a static landmark field, written in the reference's file formats
    (PosesPts.h:12-24,35-39) + a ground-truth position file with one 'x y z' row per incoming id"""
    rng = np.random.default_rng(seed)
    land = np.stack([rng.uniform(-70, 70, 30000), np.clip(rng.normal(0, 1.5, 30000), -5, 5), rng.uniform(-70, 70, 30000)], 1)
    land_i = rng.integers(40, 2000, 30000) / 8.0
    poses, pts, gt = [], [], []
    for i in range(n_pose):
        a = 4 * np.pi * i / n_pose
        cam = np.array([40 * np.cos(a), 0.0, 40 * np.sin(a)])
        yaw = -a                                               # looking along the track
        c, s = np.cos(yaw), np.sin(yaw)
        R = np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])
        t = -R @ cam
        poses.append((i, np.hstack([R, t[:, None]])))
        near = np.nonzero(np.linalg.norm(land - cam, axis=1) < 35)[0]
        sel = rng.choice(near, min(300, len(near)), replace=False)
        pts += [(i, land[k], land_i[k]) for k in sel]
        gt.append(cam)
    with open(d / "poses.txt", "w") as f:
        for i, w in poses:
            f.write(f"{i} " + " ".join("%.9g" % v for v in w.reshape(-1)) + " \n")
    with open(d / "pts.txt", "w") as f:
        for i, q, v in pts:
            f.write(f"{i} {q[0]:.9g} {q[1]:.9g} {q[2]:.9g} {v:g}\n")
    np.savetxt(d / "gt.txt", np.array(gt), fmt="%.9g")
    return str(d / "poses.txt"), str(d / "pts.txt"), str(d / "gt.txt"), np.array(gt)


def test_place_recognition_end_to_end(gpu_ctx, oracle, tmp_path):
    """files -> GPU staging -> signatures -> loop candidates -> AUC in ONE C++ process, against the same pipeline
    assembled from the Python mirror and against the oracle's decision on the same staged scans"""
    import re

    from so_dso_place_recognition_b200 import api

    subprocess.check_call(["make", "-C", HOST, "-s"])
    poses_f, pts_f, gt_f, gt_all = _write_two_laps(tmp_path)
    po, pt = np.loadtxt(poses_f), np.loadtxt(pts_f)
    for kind, polar in (("sc", False), ("m2dp", True)):
        loops = str(tmp_path / f"loops_{kind}.txt")
        out = subprocess.check_output([os.path.join(BIN, "place_recognition"), kind, poses_f, pts_f, "20", loops,
                                       "--gt", gt_f, "--loop-diff", "6"], text=True)
        got = np.loadtxt(loops)
        st = api.pts_preprocess(po[:, 0], po[:, 1:13], pt[:, 0], pt[:, 1:4], pt[:, 4], 45.0, polar)
        np.testing.assert_array_equal(got[:, 0].astype(int), st["ids"])
        gen = api.sc_generate if kind == "sc" else api.m2dp_generate
        hist = gen(st["xyz"], st["inten"], st["off"])
        gt = gt_all[st["ids"]]
        auc, tr, lp = api.run_test_full(kind, hist, hist, gt, gt, 6.0, 20)
        idx, score = api.run_test(kind, hist, hist, 20)
        np.testing.assert_array_equal(got[:, 1].astype(int) - 1, idx)
        np.testing.assert_allclose(got[:, 2], score, rtol=1e-12, atol=1e-12)
        assert abs(float(re.search(r"AUC = ([0-9.eE+-]+|nan)", out).group(1)) - auc) < 1e-5
        assert abs(float(re.search(r"top_recall = ([0-9.eE+-]+)", out).group(1)) - tr) < 1e-5
        n_loops = int(re.search(r"total_lp = (\d+)", out).group(1))
        assert n_loops > 40 and auc > 0.5                       # the second lap is recognised
        # the oracle's decision on the same staged scans
        ref = (oracle.sc_generate if kind == "sc" else oracle.m2dp_generate)(st["xyz"], st["inten"], st["off"])
        dp, di = (oracle.sc_match_numpy if kind == "sc" else oracle.m2dp_match)(ref, ref)
        ridx, rscore = oracle.fuse_top1(dp, di, 20)
        np.testing.assert_array_equal(idx, ridx)


def test_sharded_cpp_host(gpu_ctx, tmp_path):
    """match_signatures_sharded: the C++ host of the row-sharded path (one thread per GPU, sodso_comm_init +
    sodso_db_query_sharded), on 1 GPU and -- when the box has them -- 2 GPUs, against the Python binding on one GPU."""
    import torch

    from so_dso_place_recognition_b200 import api, synth

    subprocess.check_call(["make", "-C", HOST, "-s"])
    exe = os.path.join(BIN, "match_signatures_sharded")
    xyz, inten, off = synth.make_scan_set(600, 512, planted_loops=True, first=77)
    hist = api.sc_generate(xyz, inten, off)
    hfile = str(tmp_path / "history_sc.bin")
    api.save_history(hfile, hist)
    qfile = str(tmp_path / "queries.bin")
    api.save_history(qfile, hist[300:420])
    db = api.SignatureDB("sc", hist)
    idx, score, dp, di = db.query_sharded(hist[300:420], 300, 20, 2.0, 4)
    db.close()
    for gpus in ([1, 2] if torch.cuda.device_count() >= 2 else [1]):
        out = str(tmp_path / f"topk_{gpus}.txt")
        subprocess.check_call([exe, "sc", hfile, qfile, "20", "4", out, "--gpus", str(gpus), "--batch", "50", "--q-row0",
                               "300"], stdout=subprocess.DEVNULL)
        got = np.loadtxt(out).reshape(120, 4, 4)
        np.testing.assert_array_equal(got[:, :, 0].astype(np.int64) - 1, idx)
        np.testing.assert_allclose(got[:, :, 1], score, rtol=1e-9, atol=1e-12)
        np.testing.assert_array_equal(got[:, :, 2], dp)
        np.testing.assert_array_equal(got[:, :, 3], di)
