"""Real-data fixtures from the reference's committed SO-DSO outputs (place_recognition/results/), all ORACLE output
(parity unpinned: the reference ships no signatures and cannot be run here).  Run HERE; the GPU box sees the .npz.

  real_scans_multi.npz     36 real scans per descriptor (12 each from KITTI seq00, KITTI seq07 and RobotCar
                           2015-08-13-16-02-58) as staged by the oracle's pts_preprocess restatement (grid filter for
                           Scan Context, polar filter for M2DP / DELIGHT) + the oracle's SC and M2DP signatures.
  robotcar_cross_sc.npz    cross-sequence case of test_robotcar.m:26-40 (run_seq(1,:) = [5 6]: 2015-05-19-14-06-38 vs
                           2015-05-22-11-14-30, mask_width = 0, loop_diff = 25, hist1 != hist2, m != n): Scan Context
                           signatures of the first 1500 / 1400 scans reduced like the text hand-over (6 significant
                           digits), GPS positions, the oracle's decision and evaluation.
  seq06_m2dp_eval.npz      run_test('m2dp', hist, hist, gt, gt, 10, 100) (run_test.m:28, test_kitti.m:19-28) on KITTI
                           seq06: M2DP signatures (6 significant digits), the oracle's decision and evaluation.
  sign_flip_experiment.json  what changes on KITTI seq06 when the implementation-defined signs of Eigen are flipped per
                           scan (pts_align.h:31-39 eigenvectors; M2DP.cpp:94-103 dominant singular pair): fraction of
                           top-1 decisions that change, AUC / top recall -- the only available bound on divergence
                           from a real Eigen build.  tests/test_sign_conventions.py checks and (here) recomputes it.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

RES = "/root/reference/place_recognition/results"
NT = os.cpu_count() or 1


def sig6(a):
    """6 significant digits like Eigen's operator<< at default stream precision (test_sc.cpp:63-66), kept as float32"""
    return np.array([float("%.6g" % v) for v in np.asarray(a).reshape(-1)]).reshape(np.shape(a)).astype(np.float32)


def evaluate(idx, score, gt1, gt2, loop_diff, mask):
    lp, total_lp = O.gt_loops(gt1, gt2, loop_diff, mask)
    ev = O.pr_eval(score, idx, gt1, gt2, total_lp, loop_diff)
    return lp, ev


def real_scans_multi():
    out = {}
    seqs = ["KITTI/seq00", "KITTI/seq07", "RobotCar/2015-08-13-16-02-58"]
    for tag, polar in (("sc", False), ("m2dp", True)):
        xs, its, off, src = [], [], [0], []
        for name in seqs:
            d = f"{RES}/{name}/"
            st = O.stage(d + "poses_history_file.txt", d + "pts_history_file.txt", 45.0, polar)
            ns = len(st["ids"])
            for s in np.linspace(0, ns - 1, 12).astype(int):
                a, b = st["off"][s], st["off"][s + 1]
                xs.append(st["xyz"][a:b])
                its.append(st["inten"][a:b])
                off.append(off[-1] + (b - a))
                src.append(f"{name}#{s}")
        xyz, inten, off = np.concatenate(xs), np.concatenate(its), np.array(off, dtype=np.int64)
        out[tag + "_xyz"], out[tag + "_inten"], out[tag + "_off"] = xyz, inten, off
        out[tag + "_src"] = np.array(src)
        if tag == "sc":
            out["sc_hist"] = O.sc_generate(xyz, inten, off, nthreads=NT)
        else:
            out["m2dp_hist"] = O.m2dp_generate(xyz, inten, off, nthreads=NT)
            out["delight_hist"] = O.delight_generate(xyz, inten, off, nthreads=NT)
        print(tag, len(src), "scans, points per scan", int(np.diff(off).min()), "...", int(np.diff(off).max()))
    np.savez_compressed(os.path.join(HERE, "real_scans_multi.npz"), **out)


def robotcar_cross():
    runs = ["2015-05-19-14-06-38", "2015-05-22-11-14-30"]     # dates(5), dates(6): run_seq(1,:) of test_robotcar.m:9
    take = [1500, 1400]
    H, G, I = [], [], []
    for run, k in zip(runs, take):
        d = f"{RES}/RobotCar/{run}/"
        st = O.stage(d + "poses_history_file.txt", d + "pts_history_file.txt", 45.0, False)
        ids = np.loadtxt(d + "incoming_id_file.txt", dtype=np.int64)
        assert np.array_equal(ids, st["ids"])
        sub_off = st["off"][:k + 1]
        hist = O.sc_generate(st["xyz"][:sub_off[-1]], st["inten"][:sub_off[-1]], sub_off, nthreads=NT)
        gps = np.loadtxt(d + "gps.txt")
        G.append(gps[ids[:k], :3])                     # gt_full(incoming_id + 1, :) in MATLAB's 1-based rows
        H.append(hist)
        I.append(ids[:k])
    s6 = [sig6(h[:, :1200]) for h in H]
    bits = [np.packbits(h[:, 1200:].astype(np.uint8), axis=1) for h in H]
    h = [np.concatenate([a.astype(np.float64), np.unpackbits(b, axis=1)[:, :1200].astype(np.float64)], axis=1)
         for a, b in zip(s6, bits)]
    dp, di = O.sc_match_numpy(h[0], h[1])
    idx, score = O.fuse_top1(dp, di, 0)
    lp, ev = evaluate(idx, score, G[0], G[1], 25.0, 0)
    rng = np.random.default_rng(7)
    si, sj = rng.integers(0, take[0], 20000), rng.integers(0, take[1], 20000)
    print("robotcar cross: gt loops", lp.shape[0], "AUC", ev["AUC"], "top recall", ev["top_recall"])
    np.savez_compressed(os.path.join(HERE, "robotcar_cross_sc.npz"), structure6_1=s6[0], intensity_bits_1=bits[0],
                        structure6_2=s6[1], intensity_bits_2=bits[1], gt1=G[0], gt2=G[1], ids1=I[0], ids2=I[1],
                        idx=idx.astype(np.int32), score=score, n_gt_loops=np.int32(lp.shape[0]), lp_gt=lp,
                        AUC=np.float64(ev["AUC"]), top_recall=np.float64(ev["top_recall"]),
                        top_count=np.int32(ev["top_count"]), samp_i=si.astype(np.int32), samp_j=sj.astype(np.int32),
                        samp_dp=dp[si, sj], samp_di=di[si, sj], runs=np.array(runs))


def seq06_stage(polar):
    d = RES + "/KITTI/seq06/"
    st = O.stage(d + "poses_history_file.txt", d + "pts_history_file.txt", 45.0, polar)
    gt = np.loadtxt(d + "gt.txt")[st["ids"], :][:, [3, 7, 11]]
    return st, gt


def m2dp_match_blas(h1, h2):
    out = []
    for ch in range(2):
        a, b = h1[:, ch * 192:(ch + 1) * 192], h2[:, ch * 192:(ch + 1) * 192]
        m, n = a.shape[0] // 4, b.shape[0] // 4
        d = (1.0 - a @ b.T) / 2.0                                   # processM2DP.m:15
        out.append(d.reshape(m, 4, n, 4).min(axis=(1, 3)))          # :17-21
    return out


def seq06_m2dp():
    st, gt = seq06_stage(True)
    hist = O.m2dp_generate(st["xyz"], st["inten"], st["off"], nthreads=NT)
    h6 = sig6(hist)
    h = h6.astype(np.float64)
    dp, di = m2dp_match_blas(h, h)
    idx, score = O.fuse_top1(dp, di, 100)
    lp, ev = evaluate(idx, score, gt, gt, 10.0, 100)
    _, _, fused = O.fuse_top1(dp, di, 100, want_fused=True)
    part = np.partition(np.where(np.isfinite(fused), fused, np.inf), 1, axis=1)
    print("seq06 m2dp: gt loops", lp.shape[0], "AUC", ev["AUC"], "top recall", ev["top_recall"], "min margin",
          float((part[:, 1] - part[:, 0]).min()))
    np.savez_compressed(os.path.join(HERE, "seq06_m2dp_eval.npz"), hist6=h6, gt=gt, ids=st["ids"], idx=idx.astype(np.int32),
                        score=score, margin=part[:, 1] - part[:, 0], n_gt_loops=np.int32(lp.shape[0]),
                        AUC=np.float64(ev["AUC"]), top_recall=np.float64(ev["top_recall"]),
                        top_count=np.int32(ev["top_count"]))


def sign_flip_experiment(write=True):
    """-> dict (also written to sign_flip_experiment.json)"""
    res = {"sequence": "KITTI/seq06", "mask_width": 100, "loop_diff": 10.0, "seed": 20261017}
    rng = np.random.default_rng(20261017)
    # ---- Scan Context: PCA eigenvector signs (pts_align.h:31-39)
    st, gt = seq06_stage(False)
    ns = len(st["ids"])
    lp, total_lp = O.gt_loops(gt, gt, 10.0, 100)
    is_loop = np.zeros(ns, dtype=bool)
    is_loop[lp[:, 0]] = True

    def sc_run(flip):
        h = O.sc_generate_flip(st["xyz"], st["inten"], st["off"], flip, nthreads=NT)
        dp, di = O.sc_match_numpy(h, h)
        idx, score = O.fuse_top1(dp, di, 100)
        ev = O.pr_eval(score, idx, gt, gt, total_lp, 10.0)
        return h, idx, ev

    h0, idx0, ev0 = sc_run(np.zeros(ns, dtype=np.int32))
    sc = {"n_scans": ns, "n_gt_loops": int(lp.shape[0]), "baseline": {"AUC": ev0["AUC"], "top_recall": ev0["top_recall"]}}
    cases = {"v0_up_axis": 1, "v1_middle_axis": 2, "v2_major_axis": 4, "all_three": 7}
    for name, bits in cases.items():
        flip = (rng.integers(0, 2, ns) * bits if bits != 7 else rng.integers(0, 8, ns)).astype(np.int32)
        h, idx, ev = sc_run(flip)
        changed = idx != idx0
        sc[name] = {"scans_flipped": int((flip != 0).sum()),
                    "top1_changed_frac": float(changed.mean()),
                    "top1_changed_frac_gt_loop_queries": float(changed[is_loop].mean()),
                    "signature_rows_changed_frac": float((np.abs(h - h0).max(axis=1) > 0).mean()),
                    "AUC": ev["AUC"], "top_recall": ev["top_recall"]}
        print("sc", name, sc[name])
    res["scan_context"] = sc
    # ---- M2DP: PCA signs (covered by the 4 variants, test_m2dp.cpp:47-57) and the SVD pair sign (M2DP.cpp:94-103)
    st, gt = seq06_stage(True)
    lp, total_lp = O.gt_loops(gt, gt, 10.0, 100)
    is_loop = np.zeros(ns, dtype=bool)
    is_loop[lp[:, 0]] = True
    zero = np.zeros(ns, dtype=np.int32)

    def m2_run(pf, sf):
        h = O.m2dp_generate_flip(st["xyz"], st["inten"], st["off"], pf, sf, nthreads=NT)
        dp, di = m2dp_match_blas(h, h)
        idx, score = O.fuse_top1(dp, di, 100)
        ev = O.pr_eval(score, idx, gt, gt, total_lp, 10.0)
        return dp, di, idx, ev

    dp0, di0, idx0, ev0 = m2_run(zero, zero)
    m2 = {"n_scans": ns, "n_gt_loops": int(lp.shape[0]), "baseline": {"AUC": ev0["AUC"], "top_recall": ev0["top_recall"]}}
    for name, pf, sf in (("pca_all_three_random", rng.integers(0, 8, ns).astype(np.int32), zero),
                         ("svd_all_scans_flipped", zero, np.full(ns, 3, dtype=np.int32)),
                         ("svd_random_per_scan", zero, rng.integers(0, 4, ns).astype(np.int32))):
        dp, di, idx, ev = m2_run(pf, sf)
        changed = idx != idx0
        m2[name] = {"top1_changed_frac": float(changed.mean()),
                    "top1_changed_frac_gt_loop_queries": float(changed[is_loop].mean()),
                    "max_abs_distance_change": float(max(np.abs(dp - dp0).max(), np.abs(di - di0).max())),
                    "AUC": ev["AUC"], "top_recall": ev["top_recall"]}
        print("m2dp", name, m2[name])
    res["m2dp"] = m2
    if write:
        with open(os.path.join(HERE, "sign_flip_experiment.json"), "w") as f:
            json.dump(res, f, indent=1)
    return res


if __name__ == "__main__":
    what = sys.argv[1:] or ["scans", "robotcar", "m2dp", "signs"]
    if "scans" in what:
        real_scans_multi()
    if "robotcar" in what:
        robotcar_cross()
    if "m2dp" in what:
        seq06_m2dp()
    if "signs" in what:
        sign_flip_experiment()
