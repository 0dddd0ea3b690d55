"""Developer probe: sc_generate_kernel variants (sodso_debug_set_kernel_flags gen_flags): time for 5 000 and 1 000 scans x
4096 points with the fraction of the measured HBM peak, output compared with the default kernel's; with `pytest` as
first argument, runs the generation parity tests under the given variant flags instead."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from so_dso_place_recognition_b200 import api, synth, _native as N
if len(sys.argv) > 2 and sys.argv[1] == "pytest":
    import pytest
    N.lib().sodso_debug_set_kernel_flags(0, int(sys.argv[2]), 0)
    sys.exit(pytest.main(["-x", "-q", "-m", "gpu", "tests/test_gpu_generate.py", "tests/test_gpu_real_data.py",
                          "tests/test_gpu_scans_to_loops.py", "tests/test_gpu_full_size.py"]))
ctx = api.default_context(0)
variants = [2]   # experiments add their variant bits here (tools/experiments/sc_generate_*.patch)
for ns in (5000, 1000):
    xyz, inten, off = synth.make_scan_set(min(ns, 1000), 4096)
    reps = (ns + 999) // 1000
    dx = torch.from_numpy(xyz).cuda().repeat(reps, 1)[: ns * 4096]
    di = torch.from_numpy(inten).cuda().repeat(reps)[: ns * 4096]
    do = torch.arange(ns + 1, dtype=torch.int64, device="cuda") * 4096
    base = None
    for fl in variants:
        N.lib().sodso_debug_set_kernel_flags(0, fl, 0)
        t = []
        for it in range(6):
            h = api.sc_generate(dx, di, do); t.append(ctx.last_kernel_ms)
        h = h.cpu().numpy() if hasattr(h, "cpu") else np.asarray(h)
        if base is None:
            base = h
        ds = float(np.abs(base[:, :1200] - h[:, :1200]).max())
        same_i = bool(np.array_equal(base[:, 1200:], h[:, 1200:]))
        ms = min(t[1:])
        print(f"scans={ns} flags={fl}: {ms:.4f} ms -> {ns*133888/ms/1e6:.0f} GB/s frac={ns*133888/ms/1e6/6457:.3f} "
              f"max|d structure|={ds:.2e} intensity identical={same_i}", flush=True)
N.lib().sodso_debug_set_kernel_flags(0, -1, 0)
