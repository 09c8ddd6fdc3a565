#!/usr/bin/env python
"""Per-stage CUDA-event timing of the prefill path (tower, refinement, pooling, projector, splice, Llama prefill, first
token) for a batch of requests at config-c2 widths.  `python tools/prefill_breakdown.py [batch] [regions]`."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from bench import algorithmic_numbers, load_peaks, make_batch  # noqa: E402
from spatialrgpt_b200 import baseline_config, ops  # noqa: E402
from spatialrgpt_b200.llava_llama import LlavaLlamaModel  # noqa: E402
from spatialrgpt_b200.weights import random_init  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    M = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    dev = torch.device("cuda", 0)
    cfg = baseline_config("c2")
    model = LlavaLlamaModel(cfg, random_init(cfg, dev, seed=0, n_tower_layers=cfg.vision.num_hidden_layers - 1), max_seq_len=1024)
    ids, img, dep, msk = make_batch(cfg, B, M, 4321)
    ids, img, dep = ids.to(dev), img.to(dev), dep.to(dev)
    msk = [m.to(dev) for m in msk]
    v, l = cfg.vision, cfg.llama
    T, Dv, Iv, Lv = v.grid ** 2, v.hidden_size, v.intermediate_size, v.num_hidden_layers - 1
    nums = algorithmic_numbers(cfg)
    _, tensor_peak, _, _ = load_peaks()
    stages = {}

    def ev():
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        return e

    def run(record):
        marks = [("start", ev())]
        both = model.vision_tower(torch.cat([img, dep], 0))
        marks.append(("tower(2B images)", ev()))
        tf, df = both[:B].contiguous(), both[B:].contiguous()
        hres, lres = model.region_extractor.feature_refinement_nested(tf)
        marks.append(("refinement(deconv x2 + LN + avgpool)", ev()))
        me, de = model.region_extractor(hres, df, msk, hres_order=ops.ORDER_NESTED)
        marks.append(("mask pooling + region projectors", ev()))
        feats = model.mm_projector(lres)
        marks.append(("mm_projector", ev()))
        model.generate(ids, images=img, depths=dep, masks=msk, max_new_tokens=1)
        marks.append(("whole generate(max_new_tokens=1)", ev()))
        packed, lens = model._last_packed
        llm = model.llm
        for b in range(len(llm.cache.owned)):
            llm.cache.release(b)
        llm.ensure_capacity(B, max(lens) + 1)
        for b in range(B):
            llm.cache.reserve(b, lens[b] + 1)
        marks.append(("(cache reserve)", ev()))
        if B > 1:
            hid = llm.prefill_packed(packed, lens)
        else:
            hid = llm.prefill_hidden(packed, 0, 0)
        marks.append(("llama prefill layers", ev()))
        llm.first_tokens(hid, lens)
        marks.append(("final norm + lm_head + argmax", ev()))
        torch.cuda.synchronize()
        if record:
            for (n0, e0), (n1, e1) in zip(marks[:-1], marks[1:]):
                stages.setdefault(n1, []).append(e0.elapsed_time(e1))

    for i in range(4):
        run(i >= 1)
    f_vit = 2 * B * (Lv * (2 * T * (4 * Dv * Dv + 2 * Dv * Iv) + 4 * T * T * Dv) + 2 * T * 3 * v.patch_size ** 2 * Dv)
    f_attn_vit = 2 * B * Lv * 4 * T * T * Dv
    f_ref = B * (2 * T * Dv * 4 * Dv + 2 * 4 * T * Dv * 4 * Dv)
    H, I, nh, nkv, hd, V = l.hidden_size, l.intermediate_size, l.num_attention_heads, l.num_key_value_heads, l.head_dim, l.vocab_size
    S = nums["S"]
    f_llm = B * (S * l.num_hidden_layers * 2 * (H * nh * hd + 2 * H * nkv * hd + nh * hd * H + 3 * H * I) + l.num_hidden_layers * 2 * S * S * nh * hd)
    flops = {"tower(2B images)": f_vit, "refinement(deconv x2 + LN + avgpool)": f_ref, "llama prefill layers": f_llm,
             "mm_projector": B * 196 * 2 * (4 * Dv * H + H * H), "whole generate(max_new_tokens=1)": B * nums["flops_ttft"]}
    print(f"batch {B}, {M} regions; tensor peak {tensor_peak} TFLOP/s; ViT attention share of tower FLOPs {f_attn_vit / f_vit:.3f}")
    out = {}
    for k, vs in stages.items():
        ms = sorted(vs)[len(vs) // 2]
        fl = flops.get(k)
        tf_s = None if fl is None else fl / ms / 1e9
        out[k] = {"ms": round(ms, 3), "tflops": None if tf_s is None else round(tf_s, 1),
                  "frac": None if tf_s is None else round(tf_s / tensor_peak, 3)}
        print(f"  {k:42s} {ms:9.3f} ms" + ("" if tf_s is None else f"  {tf_s:7.1f} TFLOP/s  {tf_s / tensor_peak:.3f} of peak"))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
