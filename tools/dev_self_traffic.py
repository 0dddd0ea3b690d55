"""Developer probe for ncu: the 5 000 x 5 000 self-match (lower block triangle + mirrored stores) under the tc_flags given
as argument (sodso_debug_set_kernel_flags), to read dram__bytes of sc_match_tc_kernel per variant."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from so_dso_place_recognition_b200 import api, synth, _native as N
n = 5000
xyz, inten, off = synth.make_scan_set(n, 4096, planted_loops=True)
dx, di, do = torch.from_numpy(xyz).cuda(), torch.from_numpy(inten).cuda(), torch.from_numpy(off).cuda()
h = api.sc_generate(dx, di, do)
ctx = api.default_context(0)
for fl in [int(a) for a in sys.argv[1:]] or [0]:
    N.lib().sodso_debug_set_kernel_flags(fl, -1, 0)
    for it in range(2):
        idx, sc = api.run_test("sc", h, h, 100)
    print("tc_flags", fl, "match kernel ms", ctx.last_kernel_ms, "planted", (np.asarray(idx.cpu() if hasattr(idx, "cpu") else idx) == (np.arange(n) + n // 2) % n).mean(), flush=True)
N.lib().sodso_debug_set_kernel_flags(0, -1, 0)
