"""DELIGHT match timing: 2000 x 2000 signatures of synthetic scans (4096 points)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from so_dso_place_recognition_b200 import api, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
xyz, inten, off = synth.make_scan_set(n, 4096, planted_loops=True)
dx, di, do = torch.from_numpy(xyz).cuda(), torch.from_numpy(inten).cuda(), torch.from_numpy(off).cuda()
ctx = api.default_context(0)
h = api.delight_generate(dx, di, do)
print("delight_generate_kernel ms", ctx.last_kernel_ms)
for it in range(3):
    d = api.processDELIGHT(h, h)
    print("delight_match_kernel ms", ctx.last_kernel_ms, "(%d x %d)" % (n, n))
nz = (h != 0).double().mean().item()
print("histogram density %.3f" % nz)
