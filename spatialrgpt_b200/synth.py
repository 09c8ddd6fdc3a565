"""Synthetic (image, depth, regions, prompt) requests of the shapes SURVEY.md §8d defines, for
benchmarks and smoke runs (there are no datasets or checkpoints offline)."""
from __future__ import annotations

import torch
import torch.nn.functional as F

from .config import LlavaConfig
from .constants import IMAGE_TOKEN_INDEX


def synth_request(cfg: LlavaConfig, n_regions: int, t_text: int, seed: int = 1234, kind: str = "mask"):
    """Returns (input_ids [1,T] int64, images [1,3,R,R] fp32, depths [1,3,R,R] fp32, masks [M,R,R] list) on the CPU."""
    g = torch.Generator().manual_seed(seed)
    R = cfg.vision.image_size
    images = torch.rand(1, 3, R, R, generator=g) * 2 - 1
    depths = (torch.rand(1, 1, R, R, generator=g) * 2 - 1).expand(1, 3, R, R).contiguous()  # grey x3 (eval_spatial.py:105)
    masks = torch.zeros(n_regions, R, R)
    for m in range(n_regions):
        sh = int(torch.randint(R // 8, R // 2 + 1, (1,), generator=g))
        sw = int(torch.randint(R // 8, R // 2 + 1, (1,), generator=g))
        y0 = int(torch.randint(0, R - sh + 1, (1,), generator=g))
        x0 = int(torch.randint(0, R - sw + 1, (1,), generator=g))
        box = torch.zeros(R, R)
        box[y0:y0 + sh, x0:x0 + sw] = 1
        if kind == "mask":
            lo = max(R // 16, 2)
            noise = torch.rand(1, 1, lo, lo, generator=g)
            box = box * (F.interpolate(noise, (R, R), mode="bilinear", align_corners=False)[0, 0] > 0.45).float()
            if box.sum() == 0:
                box[y0, x0] = 1
        masks[m] = box
    V = cfg.llama.vocab_size
    hi = min(30000, V - 8)
    ids = [1] + torch.randint(min(1000, hi // 4), hi, (t_text - 1,), generator=g).tolist()
    ids[8] = IMAGE_TOKEN_INDEX
    p = 10
    for _ in range(n_regions):
        ids[p] = cfg.llm_mask_token_id
        p += 1
        if cfg.enable_depth:
            ids[p] = cfg.llm_depth_token_id
            p += 1
        p += 2
    if p > t_text:
        raise ValueError("t_text too short for the requested regions")
    return torch.tensor([ids], dtype=torch.long), images, depths, [masks]
