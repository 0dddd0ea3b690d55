"""Developer timing probe (run under gpurun): kernel times of the hot-path kernels on synthetic data."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from so_dso_place_recognition_b200 import api, synth

ctx = api.default_context(0)
ns = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
npts = 4096
xyz, inten, off = synth.make_scan_set(ns, npts)
dx, di, do = torch.from_numpy(xyz).cuda(), torch.from_numpy(inten).cuda(), torch.from_numpy(off).cuda()
for it in range(3):
    h = api.sc_generate(dx, di, do)
    print("sc_generate", ns, "scans:", round(ctx.last_kernel_ms, 4), "ms ->", round(ns * 133888 / ctx.last_kernel_ms / 1e6, 1), "GB/s")
if "--simt" in sys.argv:
    ctx.set_match_algo(api.SODSO_ALGO_SIMT)
    dp, dq = api.processSC(h[:256], h, f32=True)
    print("simt match 256 x", ns, ctx.last_kernel_ms, "ms", 256 * ns / ctx.last_kernel_ms / 1e3, "Mpairs/s")
ctx.set_match_algo(api.SODSO_ALGO_TC)
for m in (128, 1000, ns):
    for it in range(3):
        dp, dq = api.processSC(h[:m], h, f32=True)
        ms = ctx.last_kernel_ms
        print("tc match", m, "x", ns, round(ms, 3), "ms", round(m * ns / ms / 1e3, 1), "Mpairs/s",
              round(m * ns * 576000 / ms / 1e9, 1), "TFLOP/s algorithmic")
t = time.time()
idx, sc = api.run_test("sc", h, h, 100)
torch.cuda.synchronize()
print("run_test e2e (device sigs)", ns, time.time() - t, "s; launches", ctx.launch_count)
if "--m2dp" in sys.argv:
    nm = 592
    for it in range(2):
        hm = api.m2dp_generate(dx[:nm * npts], di[:nm * npts], do[:nm + 1])
        print("m2dp_generate", nm, "scans:", round(ctx.last_kernel_ms, 3), "ms ->", round(ctx.last_kernel_ms / nm * 1e3, 1), "us/scan")
    hh = hm.repeat(8, 1)[:4 * 4000]
    for it in range(2):
        a, b = api.processM2DP(hh, hh, f32=True)
        print("m2dp match 4000 x 4000:", round(ctx.last_kernel_ms, 3), "ms", round(16e6 / ctx.last_kernel_ms / 1e3, 1), "Mpairs/s")
