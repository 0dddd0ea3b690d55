"""Pins the full-size decisions of BASELINE.json configs[2] (Scan Context) and configs[4] (M2DP) to the ORACLE.

The 5 000-scan synthetic set (so_dso_place_recognition_b200/synth.py, seeds fixed) is pushed through the CPU
restatement of the whole path -- test_sc.cpp:36-57 / test_m2dp.cpp:37-67 (generation), processSC.m / processM2DP.m
(all 25 x 10^6 pairs), run_test.m:38-57 (fusion, mask 100, first-index arg-min) -- and the decision is stored:

  config2_oracle_decision.npz   sc_idx int32[5000], sc_score f64[5000], and 2 x 10^5 sampled pairs (i, j, d_p, d_i)
                                + the top-1 pair of every query; m2dp_idx / m2dp_score / m2dp samples likewise

Takes a few minutes of CPU (all cores); run HERE, the GPU box only sees the .npz.  tests/test_gpu_full_size.py and
bench.py (`top1_identical_to_oracle`) compare the GPU path with it.
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from so_dso_place_recognition_b200 import synth  # noqa: E402

N, NPTS, MASK, NSAMP = 5000, 4096, 100, 200_000


def m2dp_match_blas(h1, h2):
    """processM2DP.m:15-21 with BLAS: (1 - H1 H2') / 2 on the variant rows, then the 4 x 4 block minimum."""
    out = []
    for ch in range(2):
        a, b = h1[:, ch * 192:(ch + 1) * 192], h2[:, ch * 192:(ch + 1) * 192]
        m, n = a.shape[0] // 4, b.shape[0] // 4
        res = np.empty((m, n))
        for i0 in range(0, m, 250):
            d = (1.0 - a[4 * i0:4 * (i0 + 250)] @ b.T) / 2.0
            mm = d.shape[0] // 4
            res[i0:i0 + mm] = d.reshape(mm, 4, n, 4).min(axis=(1, 3))
        out.append(res)
    return out


def main():
    nt = os.cpu_count() or 1
    t0 = time.time()
    xyz, inten, off = synth.make_scan_set(N, NPTS, planted_loops=True)
    print(f"scans {time.time() - t0:.0f}s")
    rng = np.random.default_rng(4242)
    si, sj = rng.integers(0, N, NSAMP).astype(np.int32), rng.integers(0, N, NSAMP).astype(np.int32)
    out = dict(n=N, npts=NPTS, mask=MASK, samp_i=si, samp_j=sj)

    hist = O.sc_generate(xyz, inten, off, nthreads=nt)
    print(f"sc signatures {time.time() - t0:.0f}s")
    dp, di = O.sc_match_numpy(hist, hist)
    print(f"sc match {time.time() - t0:.0f}s")
    idx, score = O.fuse_top1(dp, di, MASK)
    out.update(sc_idx=idx.astype(np.int32), sc_score=score, sc_samp_dp=dp[si, sj], sc_samp_di=di[si, sj],
               sc_top_dp=dp[np.arange(N), idx], sc_top_di=di[np.arange(N), idx],
               sc_hist_sha=np.frombuffer(__import__("hashlib").sha256(hist.tobytes()).digest(), dtype=np.uint8))
    # the margin between the best and the second best fused score: how far the decision is from a tie
    _, _, fused = O.fuse_top1(dp[:200], di[:200], MASK, want_fused=True)
    part = np.partition(fused, 1, axis=1)
    out["sc_margin_first200"] = part[:, 1] - part[:, 0]
    print("sc planted loops recovered:", float((idx == (np.arange(N) + N // 2) % N).mean()),
          "min margin (first 200):", float(out["sc_margin_first200"].min()))
    del dp, di, fused

    h4 = O.m2dp_generate(xyz, inten, off, nthreads=nt)
    print(f"m2dp signatures {time.time() - t0:.0f}s")
    mp, mi = m2dp_match_blas(h4, h4)
    # the BLAS block-min against the oracle's own loop on a few rows
    rp, ri = O.m2dp_match(h4[:4 * 8], h4, nthreads=nt)
    assert np.abs(rp - mp[:8]).max() < 1e-12 and np.abs(ri - mi[:8]).max() < 1e-12
    midx, mscore = O.fuse_top1(mp, mi, MASK)
    out.update(m2dp_idx=midx.astype(np.int32), m2dp_score=mscore, m2dp_samp_dp=mp[si, sj], m2dp_samp_di=mi[si, sj])
    _, _, fused = O.fuse_top1(mp[:200], mi[:200], MASK, want_fused=True)
    part = np.partition(fused, 1, axis=1)
    out["m2dp_margin_first200"] = part[:, 1] - part[:, 0]
    part_all = None
    print("m2dp planted loops recovered:", float((midx == (np.arange(N) + N // 2) % N).mean()),
          "min margin (first 200):", float(out["m2dp_margin_first200"].min()), f"{time.time() - t0:.0f}s")
    np.savez_compressed(os.path.join(HERE, "config2_oracle_decision.npz"), **out)
    print("written", os.path.getsize(os.path.join(HERE, "config2_oracle_decision.npz")) / 1e6, "MB")


if __name__ == "__main__":
    main()
