"""Row-sharded signature database over the GPUs of one box: host-side plumbing.

The data path lives in the library, behind the C ABI (csrc/sharded.cu): one rank per GPU, an NCCL communicator owned by
the rank's sodso_ctx, and per query batch
    match -> partial row statistics -> ncclAllReduce -> fuse + mask + per-shard top-k -> ncclAllGather -> merge
all on the library stream (sodso_db_query_sharded / sodso_db_scans_query_sharded).  This module only
  * partitions rows (shard_rows: the same contiguous block rule the library checks query slices against),
  * bootstraps the library communicator from an existing torch.distributed group (init_comm: rank 0's NCCL unique id
    is broadcast as 128 bytes; torch.distributed is the control plane, never the data path),
  * keeps `protocol_reference`, a backend-agnostic restatement of the exchange over torch.distributed (gloo) that the
    CPU tests run with an oracle-backed shard: it pins WHAT the library's exchange must compute.
"""
from __future__ import annotations

import numpy as np

STATS_W = 6   # [sum(d_p-c), sum((d_p-c)^2), count_p, sum(d_i-c), sum((d_i-c)^2), count_i], c = 0.25, NaN omitted


def shard_rows(n_global: int, world: int, rank: int):
    """contiguous block partition: -> (row0, n_local)"""
    base, rem = divmod(n_global, world)
    n_local = base + (1 if rank < rem else 0)
    row0 = rank * base + min(rank, rem)
    return row0, n_local


def init_comm(ctx, group=None):
    """Give the library context its own NCCL communicator spanning the ranks of a torch.distributed group
    (sodso_comm_unique_id on rank 0 -> broadcast of the 128 id bytes -> sodso_comm_init on every rank)."""
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        ctx.comm_init(None, 1, 0)
        return
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    box = [ctx.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    ctx.comm_init(box[0], world, rank)


def protocol_reference(backend, hist_q, q_global_row0=0, mask_width=100, p_weight=2.0, k=8, group=None):
    """The sharded query protocol over torch.distributed, for a shard object with match / partial_stats / topk
    (CPU tests: gloo + an oracle-backed shard).  Every rank gets the merged (idx, score, d_p, d_i), m x k each."""
    import torch
    import torch.distributed as dist

    from . import api

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    backend.match(hist_q)
    st = torch.from_numpy(np.ascontiguousarray(backend.partial_stats()))
    if world > 1:
        dist.all_reduce(st, group=group)                       # exchange 1: global row sums and counts
    idx, score, dp, di = backend.topk(st.numpy(), q_global_row0, mask_width, p_weight, k)
    # indices travel as fp64 (exact below 2^53) so that one all-gather moves everything
    pack = torch.from_numpy(np.stack([idx.astype(np.float64), score, dp, di], axis=0))
    if world > 1:
        parts = [torch.empty_like(pack) for _ in range(world)]
        dist.all_gather(parts, pack, group=group)              # exchange 2: per-shard top-k
        g = torch.stack(parts, dim=0).numpy()
    else:
        g = pack[None].numpy()
    return api.topk_merge(g[:, 0].astype(np.int64), g[:, 1], g[:, 2], g[:, 3])
