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


def main():
    ref_shim.install()
    from llava import conversation as RC
    from llava import mm_utils as RM

    out = {"prompts": {}, "tokenize": [], "stopping": []}
    for name, tmpl in RC.conv_templates.items():
        if name in ("default", "v0"):  # few-shot template: only system/separators are mirrored (see conversation.py)
            continue
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
    with open(os.path.join(HERE, "host_api.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote host_api.json:", {k: len(v) for k, v in out.items()})


if __name__ == "__main__":
    main()
