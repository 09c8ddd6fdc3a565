"""32 c2-shaped requests through generate() with a few batched decode steps, eager launches (the command profiled by
`ncu --metrics gpu__time_duration.sum` to see where a batched decode step spends its time)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import make_batch
from spatialrgpt_b200 import baseline_config
from spatialrgpt_b200.llava_llama import LlavaLlamaModel
from spatialrgpt_b200.weights import random_init

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
n_new = int(sys.argv[2]) if len(sys.argv) > 2 else 3
cfg = baseline_config("c2")
dev = torch.device("cuda", 0)
model = LlavaLlamaModel(cfg, random_init(cfg, dev, seed=0, n_tower_layers=cfg.vision.num_hidden_layers - 1), max_seq_len=1024)
ids, img, dep, msk = make_batch(cfg, B, 4, 4321)
a = dict(images=img.to(dev), depths=dep.to(dev), masks=[m.to(dev) for m in msk], do_sample=False, use_cuda_graph=False)
model.generate(ids.to(dev), max_new_tokens=n_new, **a)
torch.cuda.synchronize()
torch.cuda.profiler.start()
out = model.generate(ids.to(dev), max_new_tokens=n_new, **a)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ids", out[:2].tolist())
