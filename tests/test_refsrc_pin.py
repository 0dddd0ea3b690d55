"""Pins the oracle's restatement against the reference's OWN source lines.

`make -C oracle refsrc` compiles SC.cpp, M2DP.cpp, DELIGHT.cpp, utils/pts_align.h, utils/pts_preprocess.h and PosesPts.h
UNCHANGED from /root/reference against oracle/eigen_shim (Eigen itself is not in the image) into
oracle/_ref/libsodso_refsrc.so.  Everything the reference's sources state -- bin arithmetic, the float / double mix, drop
conditions, accumulation order, std::unordered_map iteration order, output layout -- runs as written there and must
equal the restatement (oracle/sodso_oracle.cpp) BIT FOR BIT.  Not pinned by this: Eigen's own arithmetic (eigenvector /
singular-vector signs, product summation order), which the shim takes from the restatement.

The tests that need the library are skipped where neither the library nor /root/reference exists;
test_oracle_equals_committed_refsrc_outputs runs everywhere (fixture made by tests/golden/make_golden_refsrc.py).
"""
import os

import numpy as np
import pytest

from conftest import GOLDEN

from oracle import refsrc as R

REF_RESULTS = "/root/reference/place_recognition/results"
needs_refsrc = pytest.mark.skipif(not R.available(), reason="oracle/_ref not built and /root/reference absent")


def _synthetic(seed, nscan, lo, hi, scale, inten_hi=256):
    rng = np.random.default_rng(seed)
    sizes = rng.integers(lo, hi, nscan)
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    xyz = rng.normal(size=(off[-1], 3)) * np.asarray(scale, dtype=np.float64)
    inten = rng.integers(0, inten_hi, off[-1]).astype(np.float32)
    return xyz, inten, off


@needs_refsrc
@pytest.mark.parametrize("max_rho", [45.0, 20.0, 3.0])
def test_sc_and_m2dp_source_equal_restatement_synthetic(oracle, max_rho):
    # ragged scans; max_rho 20 / 3 push many points past the last ring (the `idx >= size` drop, SC.cpp:42, M2DP.cpp:66)
    xyz, inten, off = _synthetic(11, 24, 150, 4096, (20, 8, 2))
    np.testing.assert_array_equal(R.sc_generate(xyz, inten, off, max_rho), oracle.sc_generate(xyz, inten, off, max_rho))
    np.testing.assert_array_equal(R.m2dp_generate(xyz, inten, off, max_rho),
                                  oracle.m2dp_generate(xyz, inten, off, max_rho))


@needs_refsrc
def test_source_equal_restatement_degenerate_scans(oracle):
    # tiny scans (3..12 points), a planar scan (one zero eigenvalue), constant intensity (nothing above the average),
    # duplicated points
    xyz, inten, off = _synthetic(5, 12, 3, 12, (10, 5, 1))
    rng = np.random.default_rng(6)
    flat = rng.normal(size=(300, 3)) * np.array([12.0, 7.0, 0.0])
    dup = np.repeat(rng.normal(size=(50, 3)) * 9.0, 4, axis=0)
    xyz = np.concatenate([xyz, flat, dup])
    inten = np.concatenate([inten, np.full(300, 17, np.float32), rng.integers(0, 256, 200).astype(np.float32)])
    off = np.concatenate([off, [off[-1] + 300, off[-1] + 500]]).astype(np.int64)
    np.testing.assert_array_equal(R.sc_generate(xyz, inten, off, 45.0), oracle.sc_generate(xyz, inten, off, 45.0))
    np.testing.assert_array_equal(R.m2dp_generate(xyz, inten, off, 45.0), oracle.m2dp_generate(xyz, inten, off, 45.0))
    np.testing.assert_array_equal(R.delight_generate(xyz, inten, off), oracle.delight_generate(xyz, inten, off))


@needs_refsrc
def test_delight_source_equals_restatement(oracle):
    xyz, inten, off = _synthetic(2, 16, 500, 3000, (15, 9, 3))
    np.testing.assert_array_equal(R.delight_generate(xyz, inten, off), oracle.delight_generate(xyz, inten, off))


@needs_refsrc
def test_source_equals_restatement_on_real_scans(oracle):
    """36 staged scans of 13 sequences (KITTI + RobotCar): the fixture the GPU parity tests use."""
    g = np.load(os.path.join(GOLDEN, "real_scans_multi.npz"))
    sc = R.sc_generate(g["sc_xyz"], g["sc_inten"], g["sc_off"], 45.0)
    np.testing.assert_array_equal(sc, g["sc_hist"])
    np.testing.assert_array_equal(sc, oracle.sc_generate(g["sc_xyz"], g["sc_inten"], g["sc_off"]))
    m2 = R.m2dp_generate(g["m2dp_xyz"], g["m2dp_inten"], g["m2dp_off"], 45.0)
    np.testing.assert_array_equal(m2, oracle.m2dp_generate(g["m2dp_xyz"], g["m2dp_inten"], g["m2dp_off"]))
    np.testing.assert_array_equal(R.delight_generate(g["m2dp_xyz"], g["m2dp_inten"], g["m2dp_off"]), g["delight_hist"])


def _write_sodso_files(tmp_path, g):
    """PosesPts.h:12-40 text records; repr() round-trips the doubles and floats exactly"""
    poses, pts = tmp_path / "poses_history_file.txt", tmp_path / "pts_history_file.txt"
    with open(poses, "w") as f:
        for i, w in zip(g["pose_id"], g["w2c"]):
            f.write(str(int(i)) + " " + " ".join(repr(float(v)) for v in w) + " \n")
    with open(pts, "w") as f:
        for i, p, t in zip(g["pt_id"], g["pt_xyz"], g["pt_inten"]):
            f.write("%d %r %r %r %r\n" % (int(i), float(p[0]), float(p[1]), float(p[2]), float(t)))
    return str(poses), str(pts)


@needs_refsrc
@pytest.mark.parametrize("polar", [False, True])
def test_pts_preprocess_source_equals_restatement(oracle, tmp_path, polar):
    """the first 110 poses of KITTI seq06 (committed records): frame selection, sphere crop, voxel / polar filter and
    the ORDER the filtered points leave the std::unordered_map in"""
    g = np.load(os.path.join(GOLDEN, "staging_seq06_head.npz"))
    poses, pts = _write_sodso_files(tmp_path, g)
    ids_file = tmp_path / "incoming_id_file.txt"
    off, xyz, inten = R.stage_run(poses, pts, ids_file, 45.0, polar)
    st = oracle.stage(poses, pts, 45.0, polar)
    np.testing.assert_array_equal(np.loadtxt(ids_file, dtype=np.int64), st["ids"])
    np.testing.assert_array_equal(st["ids"], g["polar_ids" if polar else "grid_ids"])
    np.testing.assert_array_equal(off, st["off"])
    np.testing.assert_array_equal(xyz, st["xyz"])
    np.testing.assert_array_equal(inten, st["inten"])
    assert len(off) == 81 and off[-1] > 20000


@needs_refsrc
@pytest.mark.skipif(not os.path.isdir(REF_RESULTS), reason="reference data only exists in the build container")
def test_whole_sequence_through_reference_sources(oracle, tmp_path):
    """KITTI seq06 end to end through the reference's sources: pts_preprocess on the committed SO-DSO files, then
    SC (grid filter, test_sc.cpp) and M2DP (polar filter, test_m2dp.cpp) on every 8th staged scan."""
    d = REF_RESULTS + "/KITTI/seq06/"
    for polar in (False, True):
        ids_file = tmp_path / ("ids_%d.txt" % polar)
        off, xyz, inten = R.stage_run(d + "poses_history_file.txt", d + "pts_history_file.txt", ids_file, 45.0, polar)
        np.testing.assert_array_equal(np.loadtxt(ids_file, dtype=np.int64),
                                      np.loadtxt(d + "incoming_id_file.txt", dtype=np.int64))
        st = oracle.stage(d + "poses_history_file.txt", d + "pts_history_file.txt", 45.0, polar)
        np.testing.assert_array_equal(off, st["off"])
        np.testing.assert_array_equal(xyz, st["xyz"])
        np.testing.assert_array_equal(inten, st["inten"])
        pick = np.arange(0, len(off) - 1, 8)
        sub_off = np.concatenate([[0], np.cumsum(off[pick + 1] - off[pick])]).astype(np.int64)
        sub_xyz = np.concatenate([xyz[off[s]:off[s + 1]] for s in pick])
        sub_int = np.concatenate([inten[off[s]:off[s + 1]] for s in pick])
        if polar:
            np.testing.assert_array_equal(R.m2dp_generate(sub_xyz, sub_int, sub_off, 45.0),
                                          oracle.m2dp_generate(sub_xyz, sub_int, sub_off, 45.0, nthreads=8))
        else:
            np.testing.assert_array_equal(R.sc_generate(sub_xyz, sub_int, sub_off, 45.0),
                                          oracle.sc_generate(sub_xyz, sub_int, sub_off, 45.0, nthreads=8))


def test_oracle_equals_committed_refsrc_outputs(oracle):
    """runs everywhere: signatures written by the reference-source build (make_golden_refsrc.py) for committed inputs"""
    g = np.load(os.path.join(GOLDEN, "refsrc_pin.npz"))
    np.testing.assert_array_equal(oracle.sc_generate(g["xyz"], g["inten"], g["off"], 45.0), g["sc_hist"])
    np.testing.assert_array_equal(oracle.m2dp_generate(g["xyz"], g["inten"], g["off"], 45.0), g["m2dp_hist"])
    np.testing.assert_array_equal(oracle.delight_generate(g["xyz"], g["inten"], g["off"]), g["delight_hist"])
    np.testing.assert_array_equal(oracle.sc_generate(g["xyz"], g["inten"], g["off"], 12.0), g["sc_hist_rho12"])
    np.testing.assert_array_equal(oracle.m2dp_generate(g["xyz"], g["inten"], g["off"], 12.0), g["m2dp_hist_rho12"])
