"""Short workload for ncu captures of the HEADLINE step of bench.py at N = 1: sodso_db_scans_query_sharded on 5 000 scans
whose queries are the shard's own scans (every pair computed: the general match kernel), device resident, three times."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from so_dso_place_recognition_b200 import api, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
xyz, inten, off = synth.make_scan_set(n, 4096, planted_loops=True)
dx, di, do = torch.from_numpy(xyz).cuda(), torch.from_numpy(inten).cuda(), torch.from_numpy(off).cuda()
ctx = api.default_context(0)
db = api.SignatureDB("sc", api.sc_generate(dx, di, do))
for it in range(3):
    idx = db.scans_query_sharded(dx, di, do, n, 0, "same", 0, 100, 2.0, 1)[0][:, 0]
    print("match kernel ms", ctx.last_kernel_ms)
print("done", (idx == (np.arange(n) + n // 2) % n).mean())
