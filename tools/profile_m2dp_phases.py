"""Phase breakdown of m2dp_generate_kernel (sodso_debug_phase_profile): share of thread 0's clocks per phase."""
import ctypes as C
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from so_dso_place_recognition_b200 import api, synth, _native
n = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
xyz, inten, off = synth.make_scan_set(n, 4096, planted_loops=True)
dx, di, do = torch.from_numpy(xyz).cuda(), torch.from_numpy(inten).cuda(), torch.from_numpy(off).cuda()
ctx = api.default_context(0)
L = _native.lib()
for it in range(3):
    api.m2dp_generate(dx, di, do)
    print("m2dp_generate_kernel ms (no profile)", ctx.last_kernel_ms)
L.sodso_debug_phase_profile(1, None)
api.m2dp_generate(dx, di, do)
print("m2dp_generate_kernel ms (profiled)", ctx.last_kernel_ms)
out = (C.c_ulonglong * 16)()
L.sodso_debug_phase_profile(0, C.cast(out, C.c_void_p))
v = list(out)
names = {10: "moments pass", 0: "eigen-solve (1 thread)", 1: "main bin pass", 2: "twin copy", 3: "p=0 bin pass", 4: "queue replay", 5: "binarise",
         6: "Gram matrices", 7: "squarings", 8: "power iteration", 9: "A^T u + output"}
tot = sum(v[:11])
for k in sorted(names):
    print("%-18s %6.2f %%  %8.0f clk/scan" % (names[k], 100.0 * v[k] / tot, v[k] / n))
print("total clk/scan", tot / n)
print("power iterations per SVD: group 0 %.1f, group 1 %.1f" % (v[11] / (4.0 * n), v[12] / (4.0 * n)))
