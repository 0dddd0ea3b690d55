"""Developer timing probe (run under gpurun): kernel times of the hot-path kernels on synthetic data."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from so_dso_place_recognition_b200 import api, synth

ctx = api.default_context(0)
ns = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
xyz, inten, off = synth.make_scan_set(ns, 4096)
dx, di, do = torch.from_numpy(xyz).cuda(), torch.from_numpy(inten).cuda(), torch.from_numpy(off).cuda()
for it in range(5):
    h = api.sc_generate(dx, di, do)
    print("sc_generate", ns, "scans:", ctx.last_kernel_ms, "ms ->", ns * 133888 / ctx.last_kernel_ms / 1e6, "GB/s")
if "--simt" in sys.argv:
    ctx.set_match_algo(api.SODSO_ALGO_SIMT)
    for m in (256, 1000):
        t = time.time(); dp, dq = api.processSC(h[:m], h, f32=True); torch.cuda.synchronize()
        print("simt match", m, "x", ns, ctx.last_kernel_ms, "ms", m * ns / ctx.last_kernel_ms / 1e3, "pairs/s")
