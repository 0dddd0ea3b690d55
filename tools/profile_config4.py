"""BASELINE configs[3] shape on ONE GPU for ncu: a resident 6 250-row shard, query batches of 128, top-8 through
sodso_db_query_sharded (no communicator: the local part of a sharded batch).
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/x.csv python tools/profile_config4.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from so_dso_place_recognition_b200 import api, synth  # noqa: E402

nl, m, B, k = 6250, 512, 128, 8
xyz, inten, off = synth.make_scan_set(nl, 512, planted_loops=True)
dev = torch.device("cuda", 0)
ctx = api.default_context(0)
hist = api.sc_generate(torch.from_numpy(xyz).to(dev), torch.from_numpy(inten).to(dev), torch.from_numpy(off).to(dev))
hq = hist[nl // 2: nl // 2 + m].clone()
db = api.SignatureDB("sc", hist)
outs = (torch.empty((B, k), dtype=torch.int64, device=dev),) + tuple(torch.empty((B, k), dtype=torch.float64, device=dev) for _ in range(3))
for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    for b in range(0, m, B):
        db.query_sharded(hq[b:b + B], nl + b, 100, 2.0, k, out=outs)
    ctx.sync()
print("done", outs[0][:2].tolist())
