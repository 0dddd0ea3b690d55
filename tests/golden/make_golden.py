"""Generates the committed fixtures under tests/golden/ from the reference's own data files.

Run HERE (the container that has /root/reference); the GPU box only sees the .npz outputs.

  incoming_id_kat.npz   for each of the 13 sequences under place_recognition/results/: the pose
                        ids + w2c translation columns parsed from poses_history_file.txt and the
                        reference's own committed incoming_id_file.txt (an OUTPUT of pts_preprocess,
                        pts_preprocess.h:181-182,215) -- the only golden vector the reference ships.
  real_scans_seq06.npz  a few real SO-DSO scans of KITTI seq06 as staged by the oracle's restatement of
                        pts_preprocess (grid filter for SC, polar filter for M2DP) and the ORACLE's
                        signatures for them (parity unpinned: the reference ships no signatures).
"""
import glob
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

RES = "/root/reference/place_recognition/results"


def main():
    kat = {}
    names = []
    for d in sorted(glob.glob(RES + "/*/*/")):
        name = "/".join(d.rstrip("/").split("/")[-2:])
        poses = np.loadtxt(d + "poses_history_file.txt")
        ids = np.loadtxt(d + "incoming_id_file.txt", dtype=np.int64)
        key = name.replace("/", "__")
        kat[key + "__pose_id"] = poses[:, 0].astype(np.int64)
        kat[key + "__t"] = poses[:, [4, 8, 12]].astype(np.float64)
        kat[key + "__ids"] = ids
        names.append(name)
        print(name, poses.shape[0], "poses ->", ids.shape[0], "ids")
    kat["names"] = np.array(names)
    np.savez_compressed(os.path.join(HERE, "incoming_id_kat.npz"), **kat)

    d = RES + "/KITTI/seq06/"
    out = {}
    pick = [0, 100, 400, 879]
    for tag, polar in (("sc", False), ("m2dp", True)):
        st = O.stage(d + "poses_history_file.txt", d + "pts_history_file.txt", 45.0, polar)
        xs, its, off = [], [], [0]
        for s in pick:
            a, b = st["off"][s], st["off"][s + 1]
            xs.append(st["xyz"][a:b])
            its.append(st["inten"][a:b])
            off.append(off[-1] + (b - a))
        xyz = np.concatenate(xs)
        inten = np.concatenate(its)
        off = np.array(off, dtype=np.int64)
        out[tag + "_xyz"] = xyz
        out[tag + "_inten"] = inten
        out[tag + "_off"] = off
        if tag == "sc":
            out["sc_hist"] = O.sc_generate(xyz, inten, off)
        else:
            out["m2dp_hist"] = O.m2dp_generate(xyz, inten, off)
        print(tag, "scans", pick, "points", np.diff(off))
    out["pick"] = np.array(pick)
    np.savez_compressed(os.path.join(HERE, "real_scans_seq06.npz"), **out)

    # staging_seq06_head.npz: the first 110 poses of KITTI seq06 and the points they consume, as parsed from the
    # reference's committed poses_history_file.txt / pts_history_file.txt (inputs of pts_preprocess), plus the staged
    # scans of the ORACLE's restatement (both filters) as one SHA-256 per scan over its points sorted lexicographically.
    po = np.loadtxt(d + "poses_history_file.txt")[:110]
    pt = np.loadtxt(d + "pts_history_file.txt")
    pt = pt[pt[:, 0] <= po[-1, 0]]
    st_in = dict(pose_id=po[:, 0].astype(np.int32), w2c=po[:, 1:13].copy(), pt_id=pt[:, 0].astype(np.int32),
                 pt_xyz=pt[:, 1:4].copy(), pt_inten=pt[:, 4].astype(np.float32))
    for tag, polar in (("grid", False), ("polar", True)):
        r = O.stage_arrays(st_in["pose_id"], st_in["w2c"], st_in["pt_id"], st_in["pt_xyz"], st_in["pt_inten"], 45.0, polar)
        rows = np.concatenate([r["xyz"], r["inten"][:, None].astype(np.float64)], axis=1)
        dig = []
        for s in range(len(r["ids"])):
            blk = rows[r["off"][s]:r["off"][s + 1]]
            dig.append(np.frombuffer(hashlib.sha256(np.ascontiguousarray(blk[np.lexsort(blk.T[::-1])]).tobytes()).digest(),
                                     dtype=np.uint8))
        st_in[tag + "_ids"] = r["ids"]
        st_in[tag + "_off"] = r["off"]
        st_in[tag + "_sha256"] = np.stack(dig)
        print("staging", tag, len(r["ids"]), "scans", r["off"][-1], "points")
    np.savez_compressed(os.path.join(HERE, "staging_seq06_head.npz"), **st_in)

    # seq06_sc_eval.npz: Scan Context signatures of the whole KITTI seq06 (880 scans) from the oracle on the oracle's
    # staging, reduced like the reference's text hand-over (history_sc.txt carries 6 significant digits):
    # structure rounded to 6 significant digits and stored as float32, intensity as packed bits; ground-truth
    # positions gt.txt(incoming_id + 1, [4 8 12]) (test_kitti.m:23-25); and the oracle's decision + evaluation for
    # run_test('sc', hist, hist, gt, gt, 10, 100) (test_kitti.m:19-20,28).  Parity unpinned: oracle, not reference, output.
    st = O.stage(d + "poses_history_file.txt", d + "pts_history_file.txt", 45.0, False)
    hist = O.sc_generate(st["xyz"], st["inten"], st["off"])
    struct6 = np.array([float("%.6g" % v) for v in hist[:, :1200].reshape(-1)]).reshape(-1, 1200).astype(np.float32)
    bits = np.packbits(hist[:, 1200:].astype(np.uint8), axis=1)
    gt_full = np.loadtxt(d + "gt.txt")
    gt = gt_full[st["ids"], :][:, [3, 7, 11]]
    h = np.concatenate([struct6.astype(np.float64), np.unpackbits(bits, axis=1)[:, :1200].astype(np.float64)], axis=1)
    dp, di = O.sc_match_numpy(h, h)
    idx, score = O.fuse_top1(dp, di, 100)
    lp, total_lp = O.gt_loops(gt, gt, 10.0, 100)
    ev = O.pr_eval(score, idx, gt, gt, total_lp, 10.0)
    print("seq06 eval: gt loops", lp.shape[0], "AUC", ev["AUC"], "top recall", ev["top_recall"], "top count", ev["top_count"])
    np.savez_compressed(os.path.join(HERE, "seq06_sc_eval.npz"), structure6=struct6, intensity_bits=bits, gt=gt,
                        ids=st["ids"], idx=idx.astype(np.int32), score=score, n_gt_loops=np.int32(lp.shape[0]),
                        lp_gt=lp, AUC=np.float64(ev["AUC"]), top_recall=np.float64(ev["top_recall"]),
                        top_count=np.int32(ev["top_count"]))


if __name__ == "__main__":
    main()
