"""Row-sharded signature database over the GPUs of one box.

One process per GPU (torch.distributed, NCCL over NVLink; gloo on CPU for the host-logic tests).
Rank r holds DB rows [row0_r, row0_r + n_r); queries are replicated.  run_test.m:40 z-scores every
query row over the WHOLE database before the argmin (run_test.m:57), so a query batch needs exactly
two small exchanges:
  1. all-reduce (sum) of the per-query partial row sums  [m x 4 fp64]  -> global mean / std
  2. all-gather of the per-shard top-k (fused score, global index, d_p, d_i)  [4 x m x k fp64]
followed by a local k x R merge with lowest-global-index tie-break.  Payloads are KBs: latency-bound.

The per-shard compute is behind a small backend protocol: the product backend is api.SignatureDB
(tcgen05 matcher + fuse_topk kernel); the CPU tests plug in an oracle-based backend.
"""
from __future__ import annotations

import numpy as np


def shard_rows(n_global: int, world: int, rank: int):
    """contiguous block partition: -> (row0, n_local)"""
    base, rem = divmod(n_global, world)
    n_local = base + (1 if rank < rem else 0)
    row0 = rank * base + min(rank, rem)
    return row0, n_local


def _to_torch(a, device):
    import torch

    if isinstance(a, np.ndarray):
        return torch.from_numpy(np.ascontiguousarray(a)).to(device)
    return a.to(device)


def sharded_query(backend, hist_q, n_global, q_global_row0=0, mask_width=100, p_weight=2.0, k=8, group=None,
                  device=None, already_matched=False):
    """One query batch against the sharded DB.  Every rank calls this with the same hist_q and
    gets the same merged result: (idx int64 [m,k] global 0-based, score, d_p, d_i) as numpy arrays."""
    # (generation is data parallel too: see gather_query_signatures)
    import torch
    import torch.distributed as dist

    from . import api

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if not already_matched:   # (the streamed path has matched while the shard's scans were arriving)
        backend.match(hist_q)
    stats = backend.partial_stats()
    if device is None:
        device = stats.device if hasattr(stats, "device") and not isinstance(stats, np.ndarray) else "cpu"
    st = _to_torch(stats, device).contiguous()
    if world > 1:
        dist.all_reduce(st, group=group)                       # exchange 1: global row sums
    idx, score, dp, di = backend.topk(st if not isinstance(stats, np.ndarray) else st.cpu().numpy(), n_global,
                                      q_global_row0, mask_width, p_weight, k)
    # indices travel as fp64 (exact below 2^53) so that one all-gather moves everything
    pack = torch.stack([_to_torch(idx, device).double(), _to_torch(score, device), _to_torch(dp, device),
                        _to_torch(di, device)], dim=0).contiguous()
    if world > 1:
        parts = [torch.empty_like(pack) for _ in range(world)]
        dist.all_gather(parts, pack, group=group)                  # exchange 2: per-shard top-k
        gathered = torch.stack(parts, dim=0)
    else:
        gathered = pack[None]
    if gathered.is_cuda and world <= 16:
        # merge on the GPU, only the merged m x k lists cross PCIe
        oi, os_, op, od = api.topk_merge_device(gathered[:, 0].to(torch.int64), gathered[:, 1], gathered[:, 2],
                                                gathered[:, 3])
        return oi.cpu().numpy(), os_.cpu().numpy(), op.cpu().numpy(), od.cpu().numpy()
    g = gathered.cpu().numpy()
    return api.topk_merge(g[:, 0].astype(np.int64), g[:, 1], g[:, 2], g[:, 3])


def gather_query_signatures(hist_slice, group=None):
    """Generation is data-parallel over scans: every rank bins 1/R of the replicated query scans and
    the SIGNATURES (19 KB per scan instead of 115 KB of points) are all-gathered.  hist_slice: this rank's
    (m/R x 2400) torch tensor -> (m x 2400) on every rank."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return hist_slice
    out = torch.empty((world * hist_slice.shape[0],) + tuple(hist_slice.shape[1:]), dtype=hist_slice.dtype,
                      device=hist_slice.device)
    if hist_slice.is_cuda:
        dist.all_gather_into_tensor(out, hist_slice.contiguous(), group=group)
    else:
        parts = list(out.chunk(world, dim=0))
        dist.all_gather(parts, hist_slice.contiguous(), group=group)
    return out
