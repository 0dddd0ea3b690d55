"""Short workload for ncu launch lists of the non-headline kernels: M2DP generate + tcgen05 match, DELIGHT generate +
match, GPU staging, evaluation."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from so_dso_place_recognition_b200 import api, synth
from test_gpu_staging import _synthetic_sequence

n = 592
xyz, inten, off = synth.make_scan_set(n, 4096, planted_loops=True)
dx, di, do = torch.from_numpy(xyz).cuda(), torch.from_numpy(inten).cuda(), torch.from_numpy(off).cuda()
for it in range(2):
    hm = api.m2dp_generate(dx, di, do)
    api.run_test("m2dp", hm, hm, 3)
    hd = api.delight_generate(dx, di, do)
    api.run_test("delight", hd, hd, 3)
seq = _synthetic_sequence(n_pose=600, pts_per_frame=600, resets=(0, 300))
for polar in (False, True):
    st = api.pts_preprocess(*seq, 45.0, polar, device=True)
gt = np.cumsum(np.random.default_rng(0).normal(size=(n, 3)), axis=0)
api.gt_loops(gt, gt, 5.0, 50)
torch.cuda.synchronize()
print("done")
