"""N > 1 host logic on CPU: world_size-2 gloo run of the row-sharded query protocol
(stats all-reduce + top-k all-gather + merge) with an oracle-backed shard, checked against the
single-process oracle (run_test.m:38-57 on the full matrices)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from so_dso_place_recognition_b200 import sharded, synth

STAT_SHIFT = 0.25


class OracleShard:
    """per-shard compute restated with the oracle (test infrastructure)"""

    def __init__(self, O, hist_db, row0):
        self.O, self.h, self.row0 = O, hist_db, row0

    def match(self, hist_q):
        self.dp, self.di = self.O.sc_match_numpy(hist_q, self.h)

    def partial_stats(self):
        # [sum, sum of squares, count] of the non-NaN entries per channel (MATLAB normalize omits NaN)
        a, b = self.dp - STAT_SHIFT, self.di - STAT_SHIFT
        return np.stack([np.nansum(a, 1), np.nansum(a * a, 1), (~np.isnan(a)).sum(1).astype(np.float64),
                         np.nansum(b, 1), np.nansum(b * b, 1), (~np.isnan(b)).sum(1).astype(np.float64)], axis=1)

    def topk(self, gs, q_row0, mask_width, p_weight, k):
        gs = np.asarray(gs)
        Np, Ni = gs[:, 2], gs[:, 5]
        mu_p, mu_i = STAT_SHIFT + gs[:, 0] / Np, STAT_SHIFT + gs[:, 3] / Ni
        sd_p = np.sqrt((gs[:, 1] - gs[:, 0] ** 2 / Np) / (Np - 1))
        sd_i = np.sqrt((gs[:, 4] - gs[:, 3] ** 2 / Ni) / (Ni - 1))
        f = p_weight * ((self.dp - mu_p[:, None]) / sd_p[:, None]) + (self.di - mu_i[:, None]) / sd_i[:, None]
        f = np.where(np.isnan(f), np.inf, f)
        m, n = f.shape
        jg = self.row0 + np.arange(n)
        qg = q_row0 + np.arange(m)
        f = np.where(np.abs(qg[:, None] - jg[None, :]) < mask_width, np.inf, f)
        order = np.lexsort((np.broadcast_to(jg, f.shape), f), axis=1)[:, :k]
        rows = np.arange(m)[:, None]
        return jg[order].astype(np.int64), f[rows, order], self.dp[rows, order], self.di[rows, order]


def _worker(rank, world, port, hist, mask, k, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as O

    n = hist.shape[0]
    row0, n_local = sharded.shard_rows(n, world, rank)
    shard = OracleShard(O, hist[row0:row0 + n_local], row0)
    idx, score, dp, di = sharded.protocol_reference(shard, hist, 0, mask, 2.0, k)
    if rank == 0:
        out["idx"], out["score"], out["dp"], out["di"] = idx, score, dp, di
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_shard_rows_partition():
    for n, w in ((10, 3), (5000, 8), (7, 8), (50000, 8)):
        spans = [sharded.shard_rows(n, w, r) for r in range(w)]
        assert spans[0][0] == 0 and sum(s[1] for s in spans) == n
        for a, b in zip(spans, spans[1:]):
            assert a[0] + a[1] == b[0]


@pytest.mark.timeout(300)
def test_world2_gloo_matches_single_process(oracle):
    n, mask, k = 61, 4, 5
    xyz, inten, off = synth.make_scan_set(n, 1024, planted_loops=True)
    hist = oracle.sc_generate(xyz, inten, off, nthreads=4)
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), hist, mask, k, out), nprocs=2, join=True)
    dp, di = oracle.sc_match_numpy(hist, hist)
    ridx, rscore, fused = oracle.fuse_top1(dp, di, mask, want_fused=True)
    np.testing.assert_array_equal(out["idx"][:, 0], ridx)
    np.testing.assert_allclose(out["score"][:, 0], rscore, rtol=1e-9)
    # whole top-k equals the k smallest fused scores of the single-process matrix
    order = np.lexsort((np.broadcast_to(np.arange(n), fused.shape), fused), axis=1)[:, :k]
    np.testing.assert_array_equal(out["idx"], order)
    np.testing.assert_allclose(out["dp"], dp[np.arange(n)[:, None], order], rtol=0, atol=1e-15)
