#!/usr/bin/env python
"""Tensor-parallel decode (BASELINE config c5) under torchrun, one rank per GPU:

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 tools/tp_run.py [--check] [--bench]

--check  tiny GQA model (fixture sizes; widened to 2N / N heads for N > 2): generate() with TP = N must give exactly the ids of the TP-1 decoder on the same weights,
         for the CUDA-graph and the eager decode loop, with EOS handling.
--bench  c5: Llama-3-8B dims, SigLIP@448 px, 8 mask regions, depth ON, 512 greedy tokens: tokens/s at TP = N next to the TP-1
         decoder of the same process, the ids of both (must agree on the TP-1 margin-safe prefix), NVLink bytes per token.
Rank 0 prints one JSON line per mode."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--bench", action="store_true")
    ap.add_argument("--new", type=int, default=512)
    ap.add_argument("--comm", default="", help="p2p (fused NVLink peer-memory collectives, default) or nccl")
    args = ap.parse_args()
    if args.comm:
        os.environ["SRGPT_TP_COMM"] = args.comm
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from spatialrgpt_b200 import baseline_config
    from spatialrgpt_b200.llava_llama import LlavaLlamaModel
    from spatialrgpt_b200.synth import synth_request
    from spatialrgpt_b200.weights import from_state_dicts, random_init

    if args.check:
        from oracle import srgpt_oracle as O  # test infrastructure: seeded tiny weights + request
        from tests.golden.make_golden import CASES
        from tests.test_gpu_pipeline import build_model
        kw, n_regions, t_text, kind, n_new, _ = CASES["tiny_masks_gqa"]
        if world > 2:  # the fixture model has 4 / 2 heads: widen it so that every rank owns at least one KV head (head_dim stays 128)
            kw = {**kw, "hidden": 256 * world, "heads": 2 * world, "kv_heads": world, "inter": 128 * world}
        oc, sd, ref_model = build_model(kw, 7)
        tp_model = LlavaLlamaModel(ref_model.config, ref_model.weights, max_seq_len=512, tensor_parallel=(rank, world))
        ids, im, de, mk = O.synth_request(oc, n_regions, t_text, seed=1234, kind=kind)
        a = dict(images=im.to(dev), depths=de.to(dev), masks=[m.to(dev) for m in mk], do_sample=False, max_new_tokens=24)
        ref = ref_model.generate(ids.to(dev), **a)[0].tolist()
        got_graph = tp_model.generate(ids.to(dev), **a)[0].tolist()
        got_eager = tp_model.generate(ids.to(dev), use_cuda_graph=False, **a)[0].tolist()
        cut = tp_model.generate(ids.to(dev), eos_token_id=ref[5], **a)[0].tolist()
        ok = got_graph == ref and got_eager == ref and cut == ref[: ref.index(ref[5]) + 1]
        flags = torch.tensor([1 if ok else 0], device=dev)
        if world > 1:
            dist.all_reduce(flags, op=dist.ReduceOp.MIN)
        if rank == 0:
            print(json.dumps({"mode": "check", "tp": world, "comm": tp_model.llm.comm, "ok": bool(int(flags)), "ref_ids": ref, "tp_ids": got_graph}), flush=True)
        if not int(flags):
            sys.exit(1)

    if args.bench:
        cfg = baseline_config("c5")
        weights = random_init(cfg, dev, seed=0, n_tower_layers=cfg.vision.num_hidden_layers - 1)  # same seed -> identical replicas
        model = LlavaLlamaModel(cfg, weights, max_seq_len=1024, tensor_parallel=(rank, world))
        ids, im, de, mk = synth_request(cfg, 8, 64, 1234)
        a = dict(images=im.to(dev), depths=de.to(dev), masks=[mk[0].to(dev)], do_sample=False)
        ids = ids.to(dev)

        def timed(m, n_new, reps):
            for _ in range(2):
                m.generate(ids, max_new_tokens=n_new, **a)
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                out = m.generate(ids, max_new_tokens=n_new, **a)
            e1.record()
            torch.cuda.synchronize()
            ms = torch.tensor([e0.elapsed_time(e1) / reps], device=dev)
            if world > 1:
                dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            return float(ms), out[0].tolist()

        ms_full, tp_ids = timed(model, args.new, 3)
        ms_ttft, _ = timed(model, 1, 3)
        step_ms = (ms_full - ms_ttft) / (args.new - 1)
        line = {"mode": "bench", "config": "c5: Llama-3-8B TP decode, SigLIP@448, 8 mask regions, depth ON, 64-token prompt (S=259)", "tp": world, "comm": model.llm.comm if world > 1 else "none",
                "new_tokens": args.new, "tokens_per_s": round(args.new / ms_full * 1e3, 1), "ms_per_request": round(ms_full, 2), "ttft_ms": round(ms_ttft, 2),
                "decode_step_ms": round(step_ms, 4), "allreduce_bytes_per_token_per_rank": getattr(model.llm, "allreduce_bytes_per_token", 0),
                "collectives_per_token": 2 * cfg.llama.num_hidden_layers + 1}
        if rank == 0:
            ref_model = LlavaLlamaModel(cfg, weights, max_seq_len=1024)
            ref = ref_model.generate(ids, max_new_tokens=32, output_logits=True, **a)
            ref_ids, lg = ref[0][0].tolist(), ref[1][0]
            top2 = lg.topk(2, -1).values
            margin = (top2[:, 0] - top2[:, 1])
            tol = 0.06 * float(lg.std())
            safe = int((margin > 2 * tol).long().cumprod(0).sum())
            line.update({"tp1_ids_head": ref_ids[:16], "tp_ids_head": tp_ids[:16], "ids_equal_on_margin_safe_prefix": tp_ids[:safe] == ref_ids[:safe],
                         "margin_safe_prefix": safe, "ids_equal_32": tp_ids[:32] == ref_ids})
            print(json.dumps(line), flush=True)
    # leave without tearing down NCCL / the captured graphs: with collectives captured in a CUDA graph (comm=nccl) the teardown was
    # seen to hang on the 2-GPU box after all results were printed
    sys.stdout.flush()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    os._exit(0)


if __name__ == "__main__":
    main()
