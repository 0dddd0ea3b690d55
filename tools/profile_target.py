"""Short workload for ncu captures: the bench step (5k scans: sc_generate -> match -> fuse/top-1), twice."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from so_dso_place_recognition_b200 import api, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
xyz, inten, off = synth.make_scan_set(n, 4096, planted_loops=True)
dx, di, do = torch.from_numpy(xyz).cuda(), torch.from_numpy(inten).cuda(), torch.from_numpy(off).cuda()
for it in range(2):
    h = api.sc_generate(dx, di, do)
    idx, sc = api.run_test("sc", h, h, 100)
torch.cuda.synchronize()
print("done", (idx.cpu().numpy() == (np.arange(n) + n // 2) % n).mean())
