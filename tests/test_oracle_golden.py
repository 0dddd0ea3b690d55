"""Pins the oracle against the only golden vectors the reference ships: the 13 committed
incoming_id_file.txt (outputs of pts_preprocess, pts_preprocess.h:181-215).  Descriptor values
are unpinned by the reference (history_*.txt is git-ignored there), see SURVEY.md §8c."""
import os

import numpy as np
import pytest

from conftest import GOLDEN

REF = "/root/reference/place_recognition/results"


def _write_poses(path, pose_id, t):
    with open(path, "w") as f:
        for i, tr in zip(pose_id, t):
            # PosesPts.h:12-24 token order: id, then the 3x4 row-major; rotation is irrelevant to frame selection
            row = [1, 0, 0, tr[0], 0, 1, 0, tr[1], 0, 0, 1, tr[2]]
            f.write(str(int(i)) + " " + " ".join(repr(float(v)) for v in row) + " \n")


def test_incoming_id_kat_all_13_sequences(oracle, tmp_path):
    kat = np.load(os.path.join(GOLDEN, "incoming_id_kat.npz"))
    names = [str(n) for n in kat["names"]]
    assert len(names) == 13
    empty = tmp_path / "pts.txt"
    empty.write_text("")
    for name in names:
        key = name.replace("/", "__")
        poses = tmp_path / "poses.txt"
        _write_poses(poses, kat[key + "__pose_id"], kat[key + "__t"])
        for polar in (False, True):
            st = oracle.stage(str(poses), str(empty), 45.0, polar)
            assert st["n_poses"] == kat[key + "__pose_id"].shape[0]
            np.testing.assert_array_equal(st["ids"], kat[key + "__ids"], err_msg=name)


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference data only exists in the build container")
def test_incoming_id_from_reference_files(oracle):
    d = REF + "/KITTI/seq07/"
    st = oracle.stage(d + "poses_history_file.txt", d + "pts_history_file.txt", 45.0, False)
    ids = np.loadtxt(d + "incoming_id_file.txt", dtype=np.int64)
    np.testing.assert_array_equal(st["ids"], ids)
    n = np.diff(st["off"])
    assert n.min() > 100 and n.max() < 10000


def test_real_scan_fixture_matches_oracle(oracle, real_scans):
    """The committed oracle signatures of the real-scan fixture are reproducible (guards against
    silent oracle drift; the GPU tests compare against the same fixture)."""
    h = oracle.sc_generate(real_scans["sc_xyz"], real_scans["sc_inten"], real_scans["sc_off"])
    np.testing.assert_array_equal(h, real_scans["sc_hist"])
    h = oracle.m2dp_generate(real_scans["m2dp_xyz"], real_scans["m2dp_inten"], real_scans["m2dp_off"])
    np.testing.assert_allclose(h, real_scans["m2dp_hist"], rtol=0, atol=1e-12)
