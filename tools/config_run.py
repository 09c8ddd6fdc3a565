#!/usr/bin/env python
"""Full-size timing of the other BASELINE.json configurations through generate() (synthetic inputs, random-init weights):
  c1  Sheared-LLaMA-2.7B + SigLIP@336, 1 image, 2 box regions, 32 greedy tokens
  c4  Llama-2-7B (MHA, 32 KV heads) + SigLIP@448, depth ON, 16 mask regions, 96-token prompt, 256 greedy tokens (paged-KV stress)
Prints one JSON line per configuration (tokens/s over whole requests, TTFT, decode step).  `python tools/config_run.py c1 c4`"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from spatialrgpt_b200 import baseline_config  # noqa: E402
from spatialrgpt_b200.llava_llama import LlavaLlamaModel  # noqa: E402
from spatialrgpt_b200.synth import synth_request  # noqa: E402
from spatialrgpt_b200.weights import random_init  # noqa: E402

SPECS = {"c1": dict(regions=2, t_text=64, new=32, kind="box"), "c4": dict(regions=16, t_text=96, new=256, kind="mask")}


def run(name):
    sp = SPECS[name]
    dev = torch.device("cuda", 0)
    cfg = baseline_config(name)
    model = LlavaLlamaModel(cfg, random_init(cfg, dev, seed=0, n_tower_layers=cfg.vision.num_hidden_layers - 1), max_seq_len=1024)
    ids, im, de, mk = synth_request(cfg, sp["regions"], sp["t_text"], 1234, kind=sp["kind"])
    a = dict(images=im.to(dev), depths=de.to(dev), masks=[mk[0].to(dev)], do_sample=False)
    ids = ids.to(dev)

    def timed(n_new, reps):
        for _ in range(2):
            model.generate(ids, max_new_tokens=n_new, **a)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            out = model.generate(ids, max_new_tokens=n_new, **a)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps, out

    ms_full, out = timed(sp["new"], 3)
    ms_ttft, _ = timed(1, 5)
    l = cfg.llama
    w_stream = (l.num_hidden_layers * (l.hidden_size * (l.num_attention_heads + 2 * l.num_key_value_heads) * l.head_dim
                                       + l.num_attention_heads * l.head_dim * l.hidden_size + 3 * l.hidden_size * l.intermediate_size
                                       + 2 * l.hidden_size) + l.hidden_size + l.vocab_size * l.hidden_size) * 2
    step_ms = (ms_full - ms_ttft) / (sp["new"] - 1)
    print(json.dumps({"config": name, "new_tokens": sp["new"], "regions": sp["regions"], "tokens_per_s": round(sp["new"] / ms_full * 1e3, 1),
                      "ms_per_request": round(ms_full, 2), "ttft_ms": round(ms_ttft, 2), "decode_step_ms": round(step_ms, 4),
                      "weight_stream_GB": round(w_stream / 1e9, 2), "decode_GBps": round(w_stream / step_ms / 1e6, 1),
                      "ids_head": out[0, :6].tolist()}), flush=True)
    del model
    torch.cuda.empty_cache()


if __name__ == "__main__":
    for n in (sys.argv[1:] or ["c1", "c4"]):
        run(n)
