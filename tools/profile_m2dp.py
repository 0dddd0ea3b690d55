"""Short workload for ncu captures of the M2DP path: generate (592 scans x 4096 pts) + match."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from so_dso_place_recognition_b200 import api, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 592
xyz, inten, off = synth.make_scan_set(n, 4096, planted_loops=True)
dx, di, do = torch.from_numpy(xyz).cuda(), torch.from_numpy(inten).cuda(), torch.from_numpy(off).cuda()
for it in range(2):
    h = api.m2dp_generate(dx, di, do)
    idx, sc = api.run_test("m2dp", h, h, 3)
torch.cuda.synchronize()
print("done")
