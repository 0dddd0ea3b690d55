"""Developer probe (torchrun, N ranks): per-phase wall times of the sharded bench step."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from so_dso_place_recognition_b200 import api, sharded, synth

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
ctx = api.default_context(lr)
N = 5000
xq, iq, oq = synth.make_scan_set(N, 4096, planted_loops=True, first=0)
if rank == 0: xd, idn, od = xq, iq, oq
else: xd, idn, od = synth.make_scan_set(N, 4096, planted_loops=False, first=1_000_000 * rank)
d = lambda a: torch.from_numpy(a).to(dev)
dxq, diq, doq, dxd, did = d(xq), d(iq), d(oq), d(xd), d(idn)
def T(): torch.cuda.synchronize(); return time.perf_counter()
for it in range(4):
    t0 = T(); hist_db = api.sc_generate(dxd, did, doq); hist_q = api.sc_generate(dxq, diq, doq)
    t1 = T(); db = api.SignatureDB("sc", hist_db, global_row0=rank * N, ctx=ctx)
    t2 = T(); db.match(hist_q)
    t3 = T(); st = db.partial_stats(); 
    t4 = T(); dist.all_reduce(st)
    t5 = T(); idx, score, dp, di = db.topk(st, N * world, 0, 100, 2.0, 8)
    t6 = T(); pack = torch.stack([idx.double(), score, dp, di]).contiguous(); parts = [torch.empty_like(pack) for _ in range(world)]; dist.all_gather(parts, pack)
    t7 = T(); g = torch.stack(parts).cpu().numpy(); mi, ms, mp_, md = api.topk_merge(g[:, 0].astype(np.int64), g[:, 1], g[:, 2], g[:, 3])
    t8 = T(); db.close()
    t9 = T()
    print(f"rank {rank} it {it}: gen {1e3*(t1-t0):.1f} dbcreate {1e3*(t2-t1):.1f} match {1e3*(t3-t2):.1f} stats {1e3*(t4-t3):.1f} "
          f"allreduce {1e3*(t5-t4):.1f} topk {1e3*(t6-t5):.1f} allgather {1e3*(t7-t6):.1f} merge {1e3*(t8-t7):.1f} close {1e3*(t9-t8):.1f}", flush=True)
exp = (np.arange(N) + N // 2) % N
print(rank, "planted recovered", (mi[:, 0] == exp).mean(), "top1 in own shard", (mi[:, 0] < N).mean())
# single-shard reference on this rank for its own DB
i1, s1 = api.run_test("sc", hist_q, hist_q, 100)
print(rank, "single-GPU self-match recovered", (i1.cpu().numpy() == exp).mean())
dist.destroy_process_group()
