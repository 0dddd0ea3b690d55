"""N > 1 on real GPUs: world_size-2 NCCL run of the row-sharded query protocol in the shape of BASELINE
configs[3] (DB rows block-sharded 6 250 per GPU, queries streamed in batches of 128, k = 8: stats all-reduce +
per-shard top-k all-gather + merge) with the product backend (api.SignatureDB: tcgen05 matcher + fuse_topk),
checked against ONE GPU holding the whole database.  Skipped on a box with fewer than two GPUs."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from so_dso_place_recognition_b200 import api, sharded, synth

pytestmark = pytest.mark.gpu
ROWS_PER_GPU, NQ, BATCH, K, MASK = 6250, 256, 128, 8, 100


def _data(world):
    n = ROWS_PER_GPU * world
    xyz, inten, off = synth.make_scan_set(n, 128, planted_loops=True)
    return xyz, inten, off, n


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    ctx = api.default_context(rank)
    xyz, inten, off, n = _data(world)
    row0, n_local = sharded.shard_rows(n, world, rank)
    # generation is data-parallel over scans: every rank bins its own DB rows (+ the replicated queries)
    p0, p1 = off[row0], off[row0 + n_local]
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    hist_db = api.sc_generate(d(xyz[p0:p1]), d(inten[p0:p1]), d(off[row0:row0 + n_local + 1] - p0), ctx=ctx)
    qsel = np.arange(NQ) + n // 2                          # queries = scans n/2 .. n/2+NQ (their loops are rows 0..NQ)
    q0, q1 = off[qsel[0]], off[qsel[-1] + 1]
    hist_q = api.sc_generate(d(xyz[q0:q1]), d(inten[q0:q1]), d(off[qsel[0]:qsel[-1] + 2] - q0), ctx=ctx)
    db = api.SignatureDB("sc", hist_db, global_row0=row0, ctx=ctx)
    res = []
    for b in range(0, NQ, BATCH):                          # streamed query batches
        res.append(sharded.sharded_query(db, hist_q[b:b + BATCH], n, int(qsel[0]) + b, MASK, 2.0, K, device=dev))
    db.close()
    if rank == 0:
        out["idx"] = np.concatenate([r[0] for r in res])
        out["score"] = np.concatenate([r[1] for r in res])
        out["dp"] = np.concatenate([r[2] for r in res])
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.timeout(600)
def test_world2_nccl_matches_one_gpu(gpu_ctx):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run under gpurun --gpus 2)")
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    # the whole database on one GPU
    xyz, inten, off, n = _data(world)
    hist = api.sc_generate(xyz, inten, off)
    qsel = np.arange(NQ) + n // 2
    db = api.SignatureDB("sc", hist, global_row0=0)
    db.match(hist[qsel])
    st = db.partial_stats()
    idx, score, dp, di = db.topk(st, n, int(qsel[0]), MASK, 2.0, K)
    db.close()
    np.testing.assert_array_equal(out["idx"], idx)
    np.testing.assert_allclose(out["score"], score, rtol=1e-9, atol=1e-12)
    np.testing.assert_array_equal(out["dp"], dp)
    assert (idx[:, 0] == qsel - n // 2).mean() > 0.5       # most planted loops are the top-1 even at 128 points / scan
