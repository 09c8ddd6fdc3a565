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


# ---- tensor-parallel sharding (host logic of spatialrgpt_b200/tensor_parallel.py), world-size-2 gloo ----------------------------
def _tp_worker(rank, world, port, q):
    """Each rank computes its shard's share of one decoder layer in plain fp32 torch from the TPShard slices; the row-parallel partial
    sums are all-reduced over gloo and the vocabulary-parallel (value, index) candidates all-gathered, exactly the collectives of
    TPLlamaDecoder._decode_step_launch."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from types import SimpleNamespace

    from spatialrgpt_b200.config import LlamaDims
    from spatialrgpt_b200.tensor_parallel import TPShard, shard_bounds
    from spatialrgpt_b200.weights import interleave_rows

    d = LlamaDims(hidden_size=64, num_attention_heads=4, num_key_value_heads=2, head_dim=16, intermediate_size=96, vocab_size=101)
    g = torch.Generator().manual_seed(0)
    rn = lambda *s: torch.randn(*s, generator=g)  # noqa: E731  (same stream on every rank -> identical full weights)
    nh, nkv, hd, H, I = 4, 2, 16, 64, 96
    qkv_w, o_w, gate, up, down = rn((nh + 2 * nkv) * hd, H), rn(H, nh * hd), rn(I, H), rn(I, H), rn(H, I)
    lw = SimpleNamespace(qkv_w=qkv_w, o_w=o_w, gateup_w=interleave_rows(gate, up), down_w=down, in_norm=torch.ones(H), post_norm=torch.ones(H))
    sh = TPShard(d, lw, rank, world)
    x = rn(H)
    # column-parallel qkv: the rank's rows are the q heads [q0, q1) and kv heads [k0, k1) of the fused projection
    q0, q1 = shard_bounds(nh, world, rank)
    k0, k1 = shard_bounds(nkv, world, rank)
    full = qkv_w @ x
    mine = sh.qkv_w @ x
    ref = torch.cat([full[q0 * hd:q1 * hd], full[(nh + k0) * hd:(nh + k1) * hd], full[(nh + nkv + k0) * hd:(nh + nkv + k1) * hd]])
    ok = torch.allclose(mine, ref, atol=1e-5)
    # row-parallel o_proj over the rank's head slice of an "attention output", then all-reduce
    attn = rn(nh * hd)
    part = sh.o_w @ attn[q0 * hd:q1 * hd]
    dist.all_reduce(part)
    ok &= torch.allclose(part, o_w @ attn, atol=1e-4)
    # column-parallel gate/up (interleaved rows) + SwiGLU, row-parallel down, all-reduce
    i0, i1 = shard_bounds(I, world, rank)
    gu = sh.gateup_w @ x
    act = torch.nn.functional.silu(gu[0::2]) * gu[1::2]
    ok &= torch.allclose(act, (torch.nn.functional.silu(gate @ x) * (up @ x))[i0:i1], atol=1e-5)
    part = sh.down_w @ act
    dist.all_reduce(part)
    ok &= torch.allclose(part, down @ (torch.nn.functional.silu(gate @ x) * (up @ x)), atol=1e-3)
    # vocabulary-parallel arg max: all-gather (value, index), lowest index on ties
    V = d.vocab_size
    lm = rn(V, H)
    lm[77] = lm[3]
    logits = lm @ x
    per = (V + world - 1) // world
    v0, v1 = min(V, rank * per), min(V, (rank + 1) * per)
    loc = logits[v0:v1]
    cand = torch.tensor([float(loc.max()), float(v0 + int(loc.argmax()))], dtype=torch.float64)
    allc = [torch.zeros(2, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(allc, cand)
    best = max(allc, key=lambda c: (float(c[0]), -float(c[1])))
    ok &= int(best[1]) == int(logits.argmax())
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_tensor_parallel_sharding_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_tp_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [(0, True), (1, True)]
