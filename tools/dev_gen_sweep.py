"""Developer probe: sc_generate time for grid / prefetch variants (env SODSO_GEN_CTAS, SODSO_GEN_FLAGS)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from so_dso_place_recognition_b200 import api, synth
ctx = api.default_context(0)
ns = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
xyz, inten, off = synth.make_scan_set(ns, 4096)
dx, di, do = torch.from_numpy(xyz).cuda(), torch.from_numpy(inten).cuda(), torch.from_numpy(off).cuda()
for ctas in (2, 3, 4):
    for fl in (0, 2, 4, 6, 1, 7):
        os.environ["SODSO_GEN_CTAS"] = str(ctas); os.environ["SODSO_GEN_FLAGS"] = str(fl)
        t = []
        for it in range(4):
            h = api.sc_generate(dx, di, do); t.append(ctx.last_kernel_ms)
        print(f"ctas/SM={ctas} flags={fl}: {min(t[1:]):.4f} ms -> {ns*133888/min(t[1:])/1e6:.0f} GB/s")
