"""C++ host side (so_dso_place_recognition_b200/host): file formats and staging, CPU only.
  * PosesPts.h-compatible readers + pts_preprocess frame selection: the driver's incoming_id_file.txt is
    byte-identical to the reference's committed one (golden KAT).
  * Eigen `operator<<` text layout of history_*.txt (test_sc.cpp:63-66)."""
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, ROOT

HOST = os.path.join(ROOT, "so_dso_place_recognition_b200", "host")
BIN = os.path.join(ROOT, "so_dso_place_recognition_b200", "bin")


@pytest.fixture(scope="module")
def drivers():
    subprocess.check_call(["make", "-C", HOST, "-s"])
    return os.path.join(BIN, "gen_signatures"), os.path.join(BIN, "match_signatures")


def test_stage_only_reproduces_incoming_ids(drivers, tmp_path):
    kat = np.load(os.path.join(GOLDEN, "incoming_id_kat.npz"))
    pts = tmp_path / "pts.txt"
    pts.write_text("")
    for name in ("KITTI/seq06", "RobotCar/2014-11-28-12-07-13", "RobotCar/2015-05-19-14-06-38"):
        key = name.replace("/", "__")
        poses = tmp_path / "poses.txt"
        with open(poses, "w") as f:
            for i, t in zip(kat[key + "__pose_id"], kat[key + "__t"]):
                row = [1, 0, 0, t[0], 0, 1, 0, t[1], 0, 0, 1, t[2]]
                f.write(str(int(i)) + " " + " ".join(repr(float(v)) for v in row) + " \n")   # PosesPts.h:12-24
        ids = tmp_path / "ids.txt"
        for kind in ("sc", "m2dp"):
            subprocess.check_call([drivers[0], kind, str(poses), str(pts), str(tmp_path / "h.txt"), str(ids), "--stage-only"],
                                  stdout=subprocess.DEVNULL)
            got = np.loadtxt(ids, dtype=np.int64)
            np.testing.assert_array_equal(got, kat[key + "__ids"], err_msg=name)


def test_missing_parameters_return_1(drivers):
    assert subprocess.call([drivers[0]], stderr=subprocess.DEVNULL) == 1       # test_sc.cpp:19-25
    assert subprocess.call([drivers[1], "sc"], stderr=subprocess.DEVNULL) == 1


def test_history_text_layout_is_eigen_style(drivers, tmp_path):
    rng = np.random.default_rng(0)
    m = rng.normal(size=(5, 7)) * 10.0 ** rng.integers(-8, 6, size=(5, 7))
    m[0, 0] = 0
    m[1, 1] = 1
    m[2, 2] = -123456789.0
    src = tmp_path / "in.txt"
    np.savetxt(src, m, fmt="%.17g")
    out = tmp_path / "out.txt"
    subprocess.check_call([drivers[0], "reformat", str(src), str(out)])
    cells = [["%.6g" % v for v in row] for row in m]        # ostream default: %g, precision 6
    width = max(len(c) for row in cells for c in row)
    expect = "\n".join(" ".join(c.rjust(width) for c in row) for row in cells)   # no trailing newline
    assert out.read_text() == expect
    back = np.loadtxt(out)                                     # MATLAB load / numpy.loadtxt accept it
    np.testing.assert_allclose(back, m, rtol=1e-5)


def test_binary_signature_container(drivers, tmp_path):
    """the binary container written by the C++ side is lossless, mmap-able from Python, recognised by
    the C++ reader in place of a text file, and Python-written containers read back in C++."""
    from so_dso_place_recognition_b200 import api

    rng = np.random.default_rng(1)
    m = rng.normal(size=(9, 2400)) * 10.0 ** rng.integers(-12, 3, size=(9, 2400))
    src = tmp_path / "in.txt"
    np.savetxt(src, m, fmt="%.17g")
    out = tmp_path / "hist.bin"
    subprocess.check_call([drivers[0], "reformat", str(src), str(out)])      # text -> container
    assert os.path.getsize(out) == 64 + m.size * 8
    back = api.load_history(str(out))
    assert isinstance(back, np.memmap) and back.shape == m.shape
    np.testing.assert_array_equal(np.asarray(back), m)                      # lossless (17 digits round-trip)
    txt = tmp_path / "again.txt"
    subprocess.check_call([drivers[0], "reformat", str(out), str(txt)])      # container -> text (6 digits)
    np.testing.assert_allclose(np.loadtxt(txt), m, rtol=1e-5)
    py = tmp_path / "py.bin"
    api.save_history(str(py), m)
    assert py.read_bytes() == out.read_bytes()                               # same format from both sides
    np.testing.assert_array_equal(api.load_history(str(py), mmap=False), m)
    api.save_history(str(tmp_path / "py.txt"), m)
    np.testing.assert_allclose(api.load_history(str(tmp_path / "py.txt")), m, rtol=1e-5)
