"""Replica-parallel plumbing: the path shards by request, so the only communication is the
throughput report — max of the per-rank device times and an all-gather of the per-rank token counts."""
from __future__ import annotations

from typing import Tuple

import torch


def aggregate_throughput(ms_local: float, tokens_local: int, device) -> Tuple[float, int, list]:
    """Returns (max device time over ranks [ms], total tokens over ranks, per-rank token list).  Works on
    any initialised process group (NCCL on the GPUs, gloo in the CPU tests); a single process passes through."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(ms_local), int(tokens_local), [int(tokens_local)]
    t = torch.tensor([float(ms_local)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    cnt = torch.tensor([int(tokens_local)], dtype=torch.int64, device=device)
    allc = [torch.zeros_like(cnt) for _ in range(dist.get_world_size())]
    dist.all_gather(allc, cnt)
    per_rank = [int(c) for c in allc]
    return float(t), sum(per_rank), per_rank


def shard_requests(n_requests: int, rank: int, world: int) -> range:
    """Rank r takes requests r, r+world, ... (the reference's --num-chunks/--chunk-idx, eval_spatial.py:72-80)."""
    return range(rank, n_requests, world)
