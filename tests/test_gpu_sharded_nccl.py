"""N > 1 on real GPUs: world_size-2 run of the row-sharded query through the C ABI (sodso_comm_init +
sodso_db_query_sharded / sodso_db_scans_query_sharded: NCCL inside the library, on the library stream) in the shape
of BASELINE configs[3] (DB rows block-sharded 6 250 per GPU, queries streamed in batches of 128, k = 8), checked
against ONE GPU holding the whole database (run_test.m:38-57 semantics: global z-score before the arg-min, lowest
global index on ties).  Skipped on a box with fewer than two GPUs (run under `gpurun --gpus 2`); bench.py repeats
the same identity check at every N of the scaling run (`config4.sharded_topk_identical_to_single_gpu`)."""
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
# (query whose own signature is copied, shard, local row it is planted at): exact duplicates of a query have the
# minimal distance, so they tie for the top-1 -- across shards (query 5) and inside one shard (query 9)
DUP = ((5, 0, 3000), (5, 1, 3001), (9, 1, 3002), (9, 1, 3003))


def _data(world):
    n = ROWS_PER_GPU * world
    xyz, inten, off = synth.make_scan_set(n, 128, planted_loops=True)
    return xyz, inten, off, n


def _plant_duplicates(hist, world):
    """rows that are bit-identical copies of an earlier row: cross-shard ties that the merge must break by the
    lowest GLOBAL index (MATLAB's first minimum, run_test.m:57)"""
    hist = hist.copy()
    q0 = ROWS_PER_GPU * world // 2
    for q, shard, row in DUP:
        hist[shard * ROWS_PER_GPU + row] = hist[q0 + q]
    return hist


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)    # control plane only: the id broadcast
    ctx = api.default_context(rank)
    sharded.init_comm(ctx)                                          # sodso_comm_init: NCCL inside the library
    assert ctx.comm_nranks == world and ctx.comm_rank == rank
    exchange = ctx.comm_exchange                                    # 'peer-memory' where the GPUs can map each other
    xyz, inten, off, n = _data(world)
    row0, n_local = sharded.shard_rows(n, world, rank)
    hist_all = _plant_duplicates(api.sc_generate(xyz, inten, off, ctx=ctx), world)
    qsel = np.arange(NQ) + n // 2                          # queries = scans n/2 .. n/2+NQ (their loops are rows 0..NQ)
    hist_q = hist_all[qsel]
    db = api.SignatureDB("sc", hist_all[row0:row0 + n_local], global_row0=row0, ctx=ctx)
    res = []
    for b in range(0, NQ, BATCH):                          # streamed query batches, host buffers in and out
        res.append(db.query_sharded(hist_q[b:b + BATCH], int(qsel[0]) + b, MASK, 2.0, K))
    # the same batches with device buffers: only enqueued, both batches in flight, one synchronisation
    hq_dev = torch.from_numpy(hist_q).to(dev)
    outs = [db.query_sharded(hq_dev[b:b + BATCH], int(qsel[0]) + b, MASK, 2.0, K) for b in range(0, NQ, BATCH)]
    ctx.sync()
    for a, b_ in zip(res, outs):
        for x, y in zip(a, b_):
            assert np.array_equal(x, y.cpu().numpy(), equal_nan=True)
    # from POINTS: every rank bins its slice of the query scans, signatures are exchanged by NCCL, the resident
    # operand is used for the shard
    qa, na = sharded.shard_rows(NQ, world, rank)
    s0, s1 = qsel[0] + qa, qsel[0] + qa + na
    p0, p1 = off[s0], off[s1]
    r2 = db.scans_query_sharded(xyz[p0:p1], inten[p0:p1], off[s0:s1 + 1] - p0, NQ, qa, None, int(qsel[0]), MASK, 2.0, K,
                                want_hist=True)
    # ... and with the shard itself rebuilt from its scans in the same call (no planted duplicates in that one)
    d0, d1 = off[row0], off[row0 + n_local]
    r3 = db.scans_query_sharded(xyz[p0:p1], inten[p0:p1], off[s0:s1 + 1] - p0, NQ, qa,
                                (xyz[d0:d1], inten[d0:d1], off[row0:row0 + n_local + 1] - d0), int(qsel[0]), MASK, 2.0, K)
    # incremental growth: the second half of every shard is appended in two steps to a DB created with the first half
    half = n_local // 2
    db2 = api.SignatureDB("sc", hist_all[row0:row0 + half], global_row0=row0, ctx=ctx)
    db2.append(hist_all[row0 + half:row0 + half + 1000])
    db2.append(hist_all[row0 + half + 1000:row0 + n_local])
    r4 = db2.query_sharded(hist_q[:BATCH], int(qsel[0]), MASK, 2.0, K)
    db.close()
    db2.close()
    # the same first batch with the other transport: a communicator without peer windows (ncclAllReduce + ncclAllGather)
    from so_dso_place_recognition_b200 import _native as N
    ctx.comm_finalize()
    N.lib().sodso_debug_set_peer_exchange(0)
    sharded.init_comm(ctx)
    assert ctx.comm_exchange == "nccl"
    db3 = api.SignatureDB("sc", hist_all[row0:row0 + n_local], global_row0=row0, ctx=ctx)
    r5 = db3.query_sharded(hist_q[:BATCH], int(qsel[0]), MASK, 2.0, K)
    db3.close()
    N.lib().sodso_debug_set_peer_exchange(1)
    if rank == 0:
        out["exchange"] = exchange
        out["nccl_idx"], out["nccl_score"] = r5[0], r5[1]
        out["idx"] = np.concatenate([r[0] for r in res])
        out["score"] = np.concatenate([r[1] for r in res])
        out["dp"] = np.concatenate([r[2] for r in res])
        out["scans_idx"], out["scans_score"], out["scans_hist"] = r2[0], r2[1], r2[4]
        out["rebuilt_idx"] = r3[0]
        out["append_idx"], out["append_score"] = r4[0], r4[1]
    ctx.comm_finalize()
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.timeout(900)
def test_world2_sharded_cabi_matches_one_gpu(gpu_ctx):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run under gpurun --gpus 2)")
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    # the whole database on one GPU, same entry point without a communicator
    xyz, inten, off, n = _data(world)
    hist_clean = api.sc_generate(xyz, inten, off)
    hist = _plant_duplicates(hist_clean, world)
    qsel = np.arange(NQ) + n // 2
    db = api.SignatureDB("sc", hist, global_row0=0)
    idx, score, dp, di = db.query_sharded(hist[qsel], int(qsel[0]), MASK, 2.0, K)
    db.close()
    print("sharded exchange transport:", out["exchange"])
    np.testing.assert_array_equal(out["idx"], idx)
    np.testing.assert_allclose(out["score"], score, rtol=1e-9, atol=1e-12)
    np.testing.assert_array_equal(out["dp"], dp)
    np.testing.assert_array_equal(out["nccl_idx"], idx[:BATCH])             # both transports, same lists
    np.testing.assert_allclose(out["nccl_score"], score[:BATCH], rtol=1e-9, atol=1e-12)
    # the planted duplicates tie exactly and come out in global index order
    assert idx[5, :2].tolist() == [3000, ROWS_PER_GPU + 3001] and score[5, 0] == score[5, 1]
    assert idx[9, :2].tolist() == [ROWS_PER_GPU + 3002, ROWS_PER_GPU + 3003] and score[9, 0] == score[9, 1]
    np.testing.assert_array_equal(out["scans_idx"], idx)
    np.testing.assert_allclose(out["scans_score"], score, rtol=1e-9, atol=1e-12)
    np.testing.assert_array_equal(out["scans_hist"], hist[qsel])
    np.testing.assert_array_equal(out["append_idx"], idx[:BATCH])
    np.testing.assert_allclose(out["append_score"], score[:BATCH], rtol=1e-9, atol=1e-12)
    db = api.SignatureDB("sc", hist_clean, global_row0=0)
    idx_c = db.query_sharded(hist_clean[qsel], int(qsel[0]), MASK, 2.0, K)[0]
    db.close()
    np.testing.assert_array_equal(out["rebuilt_idx"], idx_c)
    assert (idx_c[:, 0] == qsel - n // 2).mean() > 0.5     # most planted loops are the top-1 even at 128 points / scan


def test_single_shard_sharded_api_equals_loop_top1(gpu_ctx, oracle):
    """no communicator: sodso_db_query_sharded on one shard = run_test.m:25-57 on the whole database"""
    xyz, inten, off = synth.make_scan_set(700, 512, planted_loops=True, first=5)
    hist = api.sc_generate(xyz, inten, off)
    db = api.SignatureDB("sc", hist, global_row0=0)
    idx, score, dp, di = db.query_sharded(hist, 0, 50, 2.0, 3)
    db.close()
    ridx, rscore = api.run_test("sc", hist, hist.copy(), 50)        # .copy(): general path, like the DB handle
    np.testing.assert_array_equal(idx[:, 0], ridx)
    np.testing.assert_allclose(score[:, 0], rscore, rtol=1e-12)
    rp, ri = oracle.sc_match_numpy(hist[:64], hist)
    oidx, oscore, fused = oracle.fuse_top1(rp, ri, 50, want_fused=True)
    np.testing.assert_array_equal(idx[:64, 0], oidx)
    order = np.lexsort((np.broadcast_to(np.arange(700), fused.shape), fused), axis=1)[:, :3]
    np.testing.assert_array_equal(idx[:64], order)


def test_db_append_equals_create(gpu_ctx):
    """sodso_db_append: growing a database row block by row block (capacity doubling, re-layout) gives the same
    distances as creating it at once; an empty shard can be grown from nothing"""
    xyz, inten, off = synth.make_scan_set(1100, 256, planted_loops=True, first=9)
    hist = api.sc_generate(xyz, inten, off)
    q = hist[500:540]
    full = api.SignatureDB("sc", hist)
    full.match(q)
    fp, fi = full.distances()
    full.close()
    db = api.SignatureDB("sc", None)
    for a, b in ((0, 1), (1, 300), (300, 301), (301, 1100)):
        db.append(hist[a:b])
    assert db.n == 1100
    db.match(q)
    gp, gi = db.distances()
    np.testing.assert_array_equal(gp, fp)
    np.testing.assert_array_equal(gi, fi)
    db.reload(hist[::-1].copy())            # reload after growth: same size, rows reversed
    db.match(q)
    rp, _ = db.distances()
    np.testing.assert_array_equal(rp, fp[:, ::-1])
    db.close()
    # M2DP shards grow too
    h4 = api.m2dp_generate(xyz[:off[40]], inten[:off[40]], off[:41])
    a = api.SignatureDB("m2dp", h4)
    a.match(h4[:32])
    ap, _ = a.distances()
    a.close()
    b = api.SignatureDB("m2dp", h4[:4 * 10])
    b.append(h4[4 * 10:])
    b.match(h4[:32])
    bp, _ = b.distances()
    b.close()
    np.testing.assert_array_equal(ap, bp)
