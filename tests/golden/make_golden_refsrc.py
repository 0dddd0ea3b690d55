"""Writes tests/golden/refsrc_pin.npz: signatures computed by the reference's OWN sources (SC.cpp, M2DP.cpp, DELIGHT.cpp,
pts_align.h compiled unchanged against oracle/eigen_shim: `make -C oracle refsrc`) for a small committed input, so that
the oracle restatement can be held against them on machines without /root/reference.  Run in the build container:

    python tests/golden/make_golden_refsrc.py

Inputs: 6 synthetic ragged scans + 6 real staged scans (KITTI / RobotCar, from real_scans_multi.npz).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import refsrc as R  # noqa: E402


def main():
    R.build()
    rng = np.random.default_rng(20261017)
    sizes = rng.integers(300, 2500, 6)
    xyz = [rng.normal(size=(n, 3)) * np.array([22.0, 9.0, 2.5]) for n in sizes]
    inten = [rng.integers(0, 256, n).astype(np.float32) for n in sizes]
    g = np.load(os.path.join(HERE, "real_scans_multi.npz"))
    for s in range(0, 36, 6):
        a, b = g["m2dp_off"][s], g["m2dp_off"][s + 1]
        xyz.append(g["m2dp_xyz"][a:b])
        inten.append(g["m2dp_inten"][a:b])
    off = np.concatenate([[0], np.cumsum([len(i) for i in inten])]).astype(np.int64)
    xyz = np.concatenate(xyz)
    inten = np.concatenate(inten)
    out = dict(xyz=xyz, inten=inten, off=off,
               sc_hist=R.sc_generate(xyz, inten, off, 45.0), m2dp_hist=R.m2dp_generate(xyz, inten, off, 45.0),
               delight_hist=R.delight_generate(xyz, inten, off),
               sc_hist_rho12=R.sc_generate(xyz, inten, off, 12.0),
               m2dp_hist_rho12=R.m2dp_generate(xyz, inten, off, 12.0))
    path = os.path.join(HERE, "refsrc_pin.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes;", len(off) - 1, "scans,", len(inten), "points")


if __name__ == "__main__":
    main()
