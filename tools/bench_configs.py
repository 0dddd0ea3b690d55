#!/usr/bin/env python
"""Secondary measurements: the BASELINE.json configs other than the headline one, one JSON line each (same conventions
as bench.py: CUDA events on the library stream after warm-up, inputs resident in HBM, synthetic data).  Single GPU.

  configs[1]  1000 synthetic scans x 4096 points: batched Scan Context generation        -> scans/s, HBM roofline
  configs[3]  50k-scan DB, 1k streaming queries in batches of 128, k = 8 (all on ONE GPU) -> pairs/s
  configs[4]  M2DP path on 5k scans: generate + match + fuse / top-1                      -> pairs/s

  python tools/bench_configs.py [--steps 5] [--warmup 3] [--only gen1k|stream50k|m2dp5k]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B  # noqa: E402  (peaks, clock sampler)


def timed(fn, steps, warmup):
    import torch

    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / steps


def main():
    import torch

    from so_dso_place_recognition_b200 import api, synth

    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    ctx = api.default_context(0)
    peaks = B.load_peaks()
    d = lambda a: torch.from_numpy(a).cuda()
    common = {"n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "higher_is_better": True, "data": "synthetic"}

    if args.only in ("", "gen1k"):
        xyz, inten, off = synth.make_scan_set(1000, 4096)
        dx, di, do = d(xyz), d(inten), d(off)
        kms = []

        def step():
            api.sc_generate(dx, di, do)
            kms.append(ctx.last_kernel_ms)

        s = timed(step, args.steps, args.warmup)
        k = float(np.mean(kms[-args.steps:])) * 1e-3
        print(json.dumps({**common, "metric": "scans/s (Scan Context generation, 4096 pts/scan)", "value": 1000 / s, "unit": "scans/s",
                          "ms_per_step": 1e3 * s, "dtype": "f64",
                          "config": {"workload": "1000 synthetic scans x 4096 pts: batched signature generation (BASELINE configs[1])"},
                          "roofline": {"bound": "hbm", "kernel": "sc_generate_kernel", "achieved": 1000 * B.SC_BYTES_PER_SCAN / k / 1e9,
                                       "peak": peaks["hbm"], "unit": "GB/s", "frac": 1000 * B.SC_BYTES_PER_SCAN / k / 1e9 / peaks["hbm"],
                                       "kernel_ms": 1e3 * k, "note": "algorithmic bytes 133 888 B/scan"}}))

    if args.only in ("", "stream50k"):
        n, m, batch, k = 50000, 1000, 128, 8
        xyz, inten, off = synth.make_scan_set(n // 10, 1024, planted_loops=True)      # 5k distinct places ...
        sig = api.sc_generate(d(xyz), d(inten), d(off))
        db_sig = sig.repeat(10, 1)                                                      # ... tiled to 50k DB rows
        db = api.SignatureDB("sc", db_sig, global_row0=0)
        q = sig[:m]

        def step():
            for b in range(0, m, batch):
                db.match(q[b:b + batch])
                st = db.partial_stats()
                db.topk(st, n, b, 100, 2.0, k)

        s = timed(step, args.steps, args.warmup)
        print(json.dumps({**common, "metric": "query x DB pair-distances/s (resident 50k-scan DB, streaming queries)",
                          "value": m * n / s, "unit": "pairs/s", "ms_per_step": 1e3 * s,
                          "dtype": "match: fp16 3-term split + e2m1, fp32 accumulate",
                          "config": {"workload": "50k-scan DB resident on one GPU, 1k queries in batches of 128, top-8 "
                                                 "(BASELINE configs[3] without the 8-way sharding)", "batch": batch, "topk": k}}))
        db.close()

    if args.only in ("", "m2dp5k"):
        xyz, inten, off = synth.make_scan_set(5000, 4096, planted_loops=True)
        dx, di, do = d(xyz), d(inten), d(off)
        res = {}

        def step():
            h = api.m2dp_generate(dx, di, do)
            res["gen_ms"] = ctx.last_kernel_ms
            res["idx"], _ = api.run_test("m2dp", h, h, 100)
            res["match_ms"] = ctx.last_kernel_ms

        s = timed(step, args.steps, args.warmup)
        rec = float((res["idx"].cpu().numpy() == (np.arange(5000) + 2500) % 5000).mean())
        print(json.dumps({**common, "metric": "query x DB pair-distances/s (M2DP generate+match+fuse)", "value": 25e6 / s,
                          "unit": "pairs/s", "ms_per_step": 1e3 * s, "dtype": "generation f64 (fp32 bin proposals); match fp16 3-term split",
                          "config": {"workload": "M2DP path: 64-plane projection + histogram + SVD signature, 5k scans (BASELINE configs[4])",
                                     "planted_loop_top1_recovered": rec},
                          "kernels_ms": {"m2dp_generate_kernel": res["gen_ms"], "m2dp_match_tc_kernel": res["match_ms"]}}))


if __name__ == "__main__":
    main()
