"""Batched decode step time of 32 c2-shaped requests (differential: generate(65 tokens) - generate(1 token)) / 64, CUDA events,
median of 3.  Prints one JSON line for tools/ab.py."""
import json, os, statistics, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import make_batch
from spatialrgpt_b200 import baseline_config
from spatialrgpt_b200.llava_llama import LlavaLlamaModel
from spatialrgpt_b200.weights import random_init

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
cfg = baseline_config("c2")
dev = torch.device("cuda", 0)
model = LlavaLlamaModel(cfg, random_init(cfg, dev, seed=0, n_tower_layers=cfg.vision.num_hidden_layers - 1), max_seq_len=1024)
ids, img, dep, msk = make_batch(cfg, B, 4, 4321)
a = dict(images=img.to(dev), depths=dep.to(dev), masks=[m.to(dev) for m in msk], do_sample=False)
ids = ids.to(dev)


def timed(n_new):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    model.generate(ids, max_new_tokens=n_new, **a)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1)


timed(65); timed(1)
steps = []
for _ in range(3):
    t1, t65 = timed(1), timed(65)
    steps.append((t65 - t1) / 64)
print(json.dumps({"kernel": f"batched_decode_step B={B}", "ms_median": round(statistics.median(steps), 4), "ms_best": round(min(steps), 4)}))
