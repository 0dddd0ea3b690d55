"""Developer probe: GPU staging (sodso_stage_points) vs the CPU oracle on a synthetic sequence."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from so_dso_place_recognition_b200 import api
from oracle import oracle as O
from test_gpu_staging import _synthetic_sequence

n_pose = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
seq = _synthetic_sequence(n_pose=n_pose, pts_per_frame=600, resets=(0, n_pose // 2))
ctx = api.default_context(0)
for polar in (False, True):
    for it in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        got = api.pts_preprocess(*seq, 45.0, polar, device=True)
        torch.cuda.synchronize(); t1 = time.perf_counter()
    t2 = time.perf_counter(); ref = O.stage_arrays(*seq, 45.0, polar); t3 = time.perf_counter()
    ns = len(ref["ids"])
    print(f"polar={polar}: {ns} scans, {ref['off'][-1]} staged points; GPU {1e3*(t1-t0):.1f} ms ({1e3*(t1-t0)/ns:.3f} ms/frame, "
          f"kernels {ctx.last_kernel_ms:.2f} ms), CPU oracle {1e3*(t3-t2):.0f} ms ({1e3*(t3-t2)/ns:.2f} ms/frame); "
          f"off equal {np.array_equal(got['off'].cpu().numpy(), ref['off'])}")
