"""N>1 plumbing on CPU: world_size-2 gloo run of the throughput aggregation used by bench.py."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from spatialrgpt_b200.distributed import aggregate_throughput, shard_requests

    ms, total, per_rank = aggregate_throughput(100.0 + 50.0 * rank, 128 * (rank + 1), torch.device("cpu"))
    mine = list(shard_requests(7, rank, world))
    dist.barrier()
    q.put((rank, ms, total, per_rank, mine))
    dist.destroy_process_group()


def test_aggregate_throughput_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ms, total, per_rank, mine in res:
        assert ms == 150.0          # max over ranks
        assert total == 128 + 256   # sum of tokens
        assert per_rank == [128, 256]
    assert res[0][4] == [0, 2, 4, 6] and res[1][4] == [1, 3, 5]


def test_single_process_passthrough():
    from spatialrgpt_b200.distributed import aggregate_throughput
    assert aggregate_throughput(12.5, 7, torch.device("cpu")) == (12.5, 7, [7])
