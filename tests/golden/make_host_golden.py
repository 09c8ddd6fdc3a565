"""Golden fixtures for the host-side API (chat templates, <image> tokenisation, stopping criterion),
produced by the REFERENCE's own llava.conversation / llava.mm_utils (authoring container only)."""
from __future__ import annotations

import json
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))
from oracle import ref_shim  # noqa: E402


class ToyTokenizer:
    """Deterministic whitespace tokenizer with a BOS token (stands in for the Llama tokenizers offline)."""
    bos_token_id = 1

    def __init__(self):
        self.vocab = {}

    def _id(self, w):
        return self.vocab.setdefault(w, 10 + len(self.vocab))

    def __call__(self, text):
        from types import SimpleNamespace
        return SimpleNamespace(input_ids=[self.bos_token_id] + [self._id(w) for w in text.split()])

    def batch_decode(self, ids, skip_special_tokens=True):
        inv = {v: k for k, v in self.vocab.items()}
        return [" ".join(inv.get(int(i), "") for i in row if int(i) != self.bos_token_id) for row in ids]


CONVERSATIONS = [
    [("U", "What is in <mask> <depth>?"), ("A", None)],
    [("U", "<image>\nHow far is <mask> <depth> from <mask> <depth>?"), ("A", "About two meters."), ("U", "And which is taller?"), ("A", None)],
    [("U", ("describe <image> this", "IMG", "Resize")), ("A", None)],
    [("U", "hello"), ("A", "hi"), ("U", "bye"), ("A", "see you")],
]
PROMPTS = ["<image>\nWhat is <mask> <depth>?", "no image here", "a <image> b <image> c", "<image>"]


def preproc_inputs():
    """Seeded synthetic PIL images / uint8 region masks shared by the generator and tests/test_host_api.py."""
    import numpy as np
    from PIL import Image

    rng = np.random.RandomState(7)
    # smooth-ish content (random low-res noise upsampled) so that bicubic resampling is exercised away from saturation
    imgs = []
    for (h, w) in ((40, 70), (64, 64), (90, 33)):
        low = rng.randint(0, 255, (h // 4 + 1, w // 4 + 1, 3), dtype=np.uint8)
        imgs.append(Image.fromarray(low).resize((w, h), Image.BILINEAR))
    masks = []
    for (h, w) in ((40, 70), (40, 70), (90, 33)):
        m = np.zeros((h, w), dtype=np.uint8)
        y0, x0 = rng.randint(0, h // 2), rng.randint(0, w // 2)
        m[y0:y0 + h // 3 + 1, x0:x0 + w // 3 + 1] = 1
        m[rng.rand(h, w) > 0.97] ^= 1
        masks.append(m)
    return imgs, masks


def make_preproc_golden(RM):
    """process_images / process_regions of the REFERENCE (llava/mm_utils.py:421-542) with the SigLIP processor, both aspect modes."""
    import numpy as np
    from types import SimpleNamespace
    from transformers import SiglipImageProcessor

    imgs, masks = preproc_inputs()
    out = {}
    for mode in ("resize", "pad"):
        proc = SiglipImageProcessor(size={"height": 56, "width": 56})
        # transformers 4.37.2 (the reference's pin): the SigLIP processor has NO crop_size attribute, which is how
        # mm_utils.py:434-441 tells it from CLIP; 5.5.0 adds crop_size=None -> restore the pinned behaviour for the reference run
        if getattr(proc, "crop_size", None) is None and "crop_size" in vars(proc):
            delattr(proc, "crop_size")
        cfg = SimpleNamespace(image_aspect_ratio=mode, image_processor=proc)
        out[f"images_{mode}"] = RM.process_images(imgs, proc, cfg).numpy()
        out[f"regions_{mode}_a"] = RM.process_regions(masks[:2], proc, cfg).numpy()
        out[f"regions_{mode}_b"] = RM.process_regions(masks[2:], proc, cfg).numpy()
    np.savez_compressed(os.path.join(HERE, "host_preproc.npz"), **out)
    print("wrote host_preproc.npz:", {k: v.shape for k, v in out.items()})


def masks_depth_inputs():
    """Region sources in the dataset's three forms (one per call so that random.choice has a single option) and a depth image."""
    import numpy as np
    from PIL import Image

    from spatialrgpt_b200.eval_spatial import rle_encode_counts
    rng = np.random.RandomState(11)
    H, W = 45, 80
    info = {"height": H, "width": W}
    boxes = [[3, 4, 40, 30], [-5, 10, 90, 44], [60, 0, 70, 45]]
    rles = []
    for _ in range(2):
        m = np.zeros((H, W), dtype=np.uint8)
        y0, x0 = rng.randint(0, H // 2), rng.randint(0, W // 2)
        m[y0:y0 + 17, x0:x0 + 23] = 1
        flat = m.flatten(order="F")  # COCO RLE is column-major
        counts, cur, run = [], 0, 0
        for v in flat:
            if v == cur:
                run += 1
            else:
                counts.append(run); cur, run = v, 1
        counts.append(run)
        rles.append({"size": [H, W], "counts": rle_encode_counts(counts)})
    low = rng.randint(0, 255, (H // 5 + 1, W // 5 + 1), dtype=np.uint8)
    depth = Image.fromarray(np.repeat(np.asarray(Image.fromarray(low).resize((W, H), Image.BILINEAR))[:, :, None], 3, axis=2))
    return info, boxes, rles, depth


def make_masks_depth_golden(RM):
    """process_masks (bbox and rle modalities) and process_depth of the REFERENCE (llava/mm_utils.py:279-418), both aspect modes.
    pycocotools is not installed here: for the rle modality the reference's ``cocomask.decode`` is served by this repo's COCO RLE
    decoder, so that case pins everything AFTER the decode (nearest resize / padding / mask processor), not the decoder itself
    (tests/test_eval_driver_cpu.py checks the decoder against hand-built masks)."""
    import numpy as np
    from types import SimpleNamespace
    from transformers import SiglipImageProcessor

    from spatialrgpt_b200.eval_spatial import rle_decode
    sys.modules["pycocotools"].mask.decode = rle_decode
    RM.cocomask.decode = rle_decode
    info, boxes, rles, depth = masks_depth_inputs()
    out = {}
    for mode in ("resize", "pad"):
        proc = SiglipImageProcessor(size={"height": 56, "width": 56})
        if getattr(proc, "crop_size", None) is None and "crop_size" in vars(proc):
            delattr(proc, "crop_size")
        cfg = SimpleNamespace(image_aspect_ratio=mode, image_processor=proc)
        out[f"masks_bbox_{mode}"] = RM.process_masks([{"bbox": boxes, "image_info": info}], cfg).numpy()
        out[f"masks_bbox_info_{mode}"] = RM.process_masks([{"bbox": boxes}], cfg, image_info=info).numpy()
        out[f"masks_rle_{mode}"] = RM.process_masks([{"rle": rles}], cfg).numpy()
        out[f"depth_{mode}"] = RM.process_depth(depth, cfg, None).numpy()
    np.savez_compressed(os.path.join(HERE, "host_masks_depth.npz"), **out)
    print("wrote host_masks_depth.npz:", {k: v.shape for k, v in out.items()})


def main():
    ref_shim.install()
    from llava import conversation as RC
    from llava import mm_utils as RM

    out = {"prompts": {}, "tokenize": [], "stopping": []}
    for name, tmpl in RC.conv_templates.items():
        rows = []
        for conv_msgs in CONVERSATIONS:
            c = tmpl.copy()
            try:
                for who, msg in conv_msgs:
                    c.append_message(c.roles[0] if who == "U" else c.roles[1], msg)
                rows.append(c.get_prompt())
            except Exception as e:  # the reference itself raises for some template/conversation combinations
                rows.append(f"raises {type(e).__name__}")
        out["prompts"][name] = rows
    for lstrip in (False, True):
        for p in PROMPTS:
            tok = ToyTokenizer()
            out["tokenize"].append({"prompt": p, "lstrip": lstrip, "ids": RM.tokenizer_image_token(p, tok, lstrip=lstrip)})
    tok = ToyTokenizer()
    base = tok("the red chair is left of the table </s> extra").input_ids
    crit = RM.KeywordsStoppingCriteria(["</s>"], tok, torch.zeros(1, 0, dtype=torch.long))
    for n in range(1, len(base) + 1):
        out["stopping"].append(bool(crit(torch.tensor([base[:n]]), None)))
    out["stopping_ids"] = base
    out["model_names"] = {p: RM.get_model_name_from_path(p) for p in ["a/b/SpatialRGPT-VILA1.5-8B", "x/run1/checkpoint-500/", "solo"]}
    make_preproc_golden(RM)
    make_masks_depth_golden(RM)
    with open(os.path.join(HERE, "host_api.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote host_api.json:", {k: len(v) for k, v in out.items()})


if __name__ == "__main__":
    main()
