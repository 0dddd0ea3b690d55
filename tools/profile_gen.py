"""Short workload for ncu captures of sc_generate_kernel: 5 000 scans x 4096 points, device resident."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from so_dso_place_recognition_b200 import api, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
xyz, inten, off = synth.make_scan_set(min(n, 1000), 4096)
reps = (n + 999) // 1000
xyz = torch.from_numpy(xyz).cuda().repeat(reps, 1)[: n * 4096]
inten = torch.from_numpy(inten).cuda().repeat(reps)[: n * 4096]
off = torch.arange(n + 1, dtype=torch.int64, device="cuda") * 4096
ctx = api.default_context(0)
for it in range(3):
    h = api.sc_generate(xyz, inten, off)
    print("sc_generate_kernel ms", ctx.last_kernel_ms)
