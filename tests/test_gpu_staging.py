"""GPU parity of the point staging (pts_preprocess.h:169-232: sliding accumulation, w2c transform, 45 m crop, voxel-grid
/ polar de-duplication) against the CPU oracle.  The staged point SET of every scan must be identical, bit for bit (the
reference orders a scan by libstdc++ hash iteration, the GPU by voxel index: compared after sorting)."""
import hashlib
import os

import numpy as np
import pytest

from conftest import GOLDEN
from so_dso_place_recognition_b200 import api

pytestmark = pytest.mark.gpu


def _sorted_rows(st, s):
    blk = np.concatenate([st["xyz"][st["off"][s]:st["off"][s + 1]],
                          st["inten"][st["off"][s]:st["off"][s + 1], None].astype(np.float64)], axis=1)
    return np.ascontiguousarray(blk[np.lexsort(blk.T[::-1])])


def _synthetic_sequence(n_pose=150, pts_per_frame=400, seed=5, resets=(0, 70)):
    """camera moving forward and turning; VO re-initialisations (|t| < 1, pts_preprocess.h:189) at `resets`;
    unsorted-looking but non-decreasing ids with gaps, several points per id, some ids without points"""
    rng = np.random.default_rng(seed)
    pose_id, w2c, pt_id, pt_xyz, pt_inten = [], [], [], [], []
    cam = np.zeros(3)
    for i in range(n_pose):
        yaw = 0.015 * i
        c, s = np.cos(yaw), np.sin(yaw)
        R = np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])
        if i in resets:
            cam = np.array([0.05, 0.0, 0.1])          # |t| < 1 -> reset
        else:
            cam = cam + np.array([0.25 * np.sin(yaw), 0.0, 0.9])
            if np.linalg.norm(R @ cam) < 1.5:
                cam = cam + np.array([0, 0, 2.0])
        t = -R @ cam
        pose_id.append(3 * i + 1)
        w2c.append(np.hstack([R, t[:, None]]).reshape(-1))
        if i % 7 != 3:                                  # some frames contribute no points
            k = pts_per_frame
            p = cam + np.stack([rng.uniform(-40, 40, k), np.clip(rng.normal(0, 1.5, k), -5, 5), rng.uniform(-15, 60, k)], 1)
            p[:8] = p[0]                                # exact duplicates: ties on the selection key
            pt_xyz.append(p)
            pt_id += [3 * i + (j % 2) for j in range(k)]
            pt_inten.append(rng.uniform(5, 250, k).astype(np.float32))
    order = np.argsort(np.array(pt_id), kind="stable")  # the VO wrapper sorts by id (OutputWrapperSODSO.cpp:16-18)
    return (np.array(pose_id, np.int32), np.array(w2c), np.array(pt_id, np.int32)[order],
            np.concatenate(pt_xyz)[order], np.concatenate(pt_inten)[order])


@pytest.mark.parametrize("polar", [False, True])
def test_synthetic_sequence_vs_oracle(gpu_ctx, oracle, polar):
    seq = _synthetic_sequence()
    ref = oracle.stage_arrays(*seq, 45.0, polar)
    got = api.pts_preprocess(*seq, 45.0, polar)
    np.testing.assert_array_equal(got["ids"], ref["ids"])
    np.testing.assert_array_equal(got["off"], ref["off"])
    assert len(ref["ids"]) == 150 - 2 * 30 and ref["off"][-1] > 10000
    for s in range(len(ref["ids"])):
        np.testing.assert_array_equal(_sorted_rows(got, s), _sorted_rows(ref, s))
    if not polar:   # canonical order: ascending voxel index (pts_preprocess.h:71-75)
        x = got["xyz"]
        loc = (np.floor((x[:, 0] + 45) / 1.5) + 61 * np.floor((x[:, 1] + 45) / 0.75) + 61 * 121 * np.floor((x[:, 2] + 45) / 1.5))
        for s in range(len(ref["ids"])):
            assert np.all(np.diff(loc[got["off"][s]:got["off"][s + 1]]) > 0)


@pytest.mark.parametrize("polar,tag", [(False, "grid"), (True, "polar")])
def test_real_seq06_head_golden(gpu_ctx, polar, tag):
    """first 110 poses of KITTI seq06 (the reference's committed SO-DSO output) vs the oracle's staged scans"""
    g = np.load(os.path.join(GOLDEN, "staging_seq06_head.npz"))
    got = api.pts_preprocess(g["pose_id"], g["w2c"], g["pt_id"], g["pt_xyz"], g["pt_inten"], 45.0, polar)
    np.testing.assert_array_equal(got["ids"], g[tag + "_ids"])
    np.testing.assert_array_equal(got["off"], g[tag + "_off"])
    for s in range(len(got["ids"])):
        dig = np.frombuffer(hashlib.sha256(_sorted_rows(got, s).tobytes()).digest(), dtype=np.uint8)
        np.testing.assert_array_equal(dig, g[tag + "_sha256"][s])


def test_staged_scans_feed_generation(gpu_ctx, oracle):
    """device-resident staged scans go straight into sc_generate; the signatures equal the oracle's on the same
    (voxel-ordered) points"""
    seq = _synthetic_sequence(n_pose=60, resets=(0,))
    st = api.pts_preprocess(*seq, 45.0, False, device=True)
    h = api.sc_generate(st["xyz"], st["inten"], st["off"])
    xyz, inten, off = st["xyz"].cpu().numpy(), st["inten"].cpu().numpy(), st["off"].cpu().numpy()
    ref = oracle.sc_generate(xyz, inten, off)
    np.testing.assert_array_equal(h.cpu().numpy()[:, 1200:], ref[:, 1200:])
    np.testing.assert_allclose(h.cpu().numpy()[:, :1200], ref[:, :1200], rtol=0, atol=1e-9)


def test_degenerate_inputs(gpu_ctx, oracle):
    pose_id = np.arange(40, dtype=np.int32)
    w2c = np.tile(np.array([1, 0, 0, 5, 0, 1, 0, 0, 0, 0, 1, 0], float), (40, 1))
    # no points at all
    got = api.pts_preprocess(pose_id, w2c, np.zeros(0, np.int32), np.zeros((0, 3)), np.zeros(0, np.float32))
    assert len(got["ids"]) == 10 and got["off"][-1] == 0
    # fewer than INIT_FRAME poses: no scans
    got = api.pts_preprocess(pose_id[:20], w2c[:20], np.zeros(3, np.int32), np.zeros((3, 3)), np.ones(3, np.float32))
    assert len(got["ids"]) == 0
    # all points out of range
    far = np.full((5, 3), 1e3)
    got = api.pts_preprocess(pose_id, w2c, np.zeros(5, np.int32), far, np.ones(5, np.float32))
    ref = oracle.stage_arrays(pose_id, w2c, np.zeros(5, np.int32), far, np.ones(5, np.float32))
    np.testing.assert_array_equal(got["off"], ref["off"])
