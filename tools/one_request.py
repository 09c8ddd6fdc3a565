"""One c2-shaped request through generate() with few decode tokens — the command profiled by ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from spatialrgpt_b200 import baseline_config
from spatialrgpt_b200.llava_llama import LlavaLlamaModel
from spatialrgpt_b200.synth import synth_request
from spatialrgpt_b200.weights import random_init

n_new = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 8
use_graph = "--graph" in sys.argv
cfg = baseline_config("c2")
dev = torch.device("cuda", 0)
model = LlavaLlamaModel(cfg, random_init(cfg, dev, seed=0, n_tower_layers=cfg.vision.num_hidden_layers - 1), max_seq_len=1024)
ids, im, de, mk = synth_request(cfg, 8, 64, 1234)
args = dict(images=im.to(dev), depths=de.to(dev), masks=[mk[0].to(dev)], do_sample=False, max_new_tokens=n_new, use_cuda_graph=use_graph)
model.generate(ids.to(dev), **args)  # warm-up (lazy kernel attribute setup, allocator)
torch.cuda.synchronize()
torch.cuda.profiler.start()  # ncu --profile-from-start off: only the request below is profiled
out = model.generate(ids.to(dev), **args)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ids", out[0].tolist())
