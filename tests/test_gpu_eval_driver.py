"""The SpatialRGPT-Bench driver end to end on the GPU with a tiny random-weight model: annotation -> regions -> depth
post-processing kernel -> generate() -> JSONL (the reference's llava/eval/eval_spatial.py flow)."""
import json
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from spatialrgpt_b200 import eval_spatial as E
from tests.golden.make_golden import CASES
from tests.golden.make_host_golden import ToyTokenizer
from tests.test_gpu_pipeline import build_model

pytestmark = pytest.mark.gpu


def test_driver_runs_the_cuda_path(tmp_path):
    from PIL import Image
    from transformers import SiglipImageProcessor
    kw = CASES["tiny_boxes"][0]
    oc, sd, model = build_model(kw, 17)
    proc = SiglipImageProcessor(size={"height": oc.image_size, "width": oc.image_size})
    model.config.image_aspect_ratio = "resize"
    tok = ToyTokenizer()
    tok.vocab.update({"<mask>": oc.mask_token_id, "<depth>": oc.depth_token_id})  # the region tokens map to the model's special ids
    tok.batch_decode = lambda ids, skip_special_tokens=True: [" ".join(str(int(i)) for i in ids[0])]
    Image.fromarray(np.random.RandomState(1).randint(0, 255, (40, 60, 3), dtype=np.uint8)).save(tmp_path / "a.jpg")
    ann = [{"id": 7, "image_info": {"file_path": "a.jpg", "height": 40, "width": 60}, "text_q": "q", "qa_info": {},
            "bbox": [[2, 3, 30, 30], [10, 5, 55, 38]],
            "conversations": [{"from": "human", "value": "<image>\n how far is <mask> from <mask> ?"}, {"from": "gpt", "value": "gt"}]}]
    (tmp_path / "ann.json").write_text(json.dumps(ann))
    args = SimpleNamespace(model_path="m/tiny", model_base=None, image_folder=str(tmp_path), annotation_file=str(tmp_path / "ann.json"),
                           answers_file=str(tmp_path / "ans.jsonl"), conv_mode="llava_v1", num_chunks=1, chunk_idx=0, temperature=0.0,
                           top_p=None, num_beams=1, use_mask=False)
    depth_calls = []

    def depth_predictor(rgb):
        depth_calls.append(rgb.shape)
        g = torch.Generator().manual_seed(0)
        return torch.rand(1, 24, 36, generator=g)  # a low-resolution depth map, like the depth network's output

    n = E.eval_model(args, depth_predictor=depth_predictor, loader=lambda p, name, base: (tok, model, proc, 4096))
    rec = [json.loads(l) for l in open(args.answers_file)]
    assert n == 1 and rec[0]["question_id"] == 7 and rec[0]["gt"] == "gt" and depth_calls == [(40, 60, 3)]
    ids = [int(x) for x in rec[0]["pred"].split()]
    assert 1 <= len(ids) <= 128 and all(0 <= i < oc.vocab for i in ids)
    # the same request through generate() directly gives the same tokens (the driver adds nothing to the hot path)
    line = ann[0]
    masks = torch.vstack([E._mask_processor(proc).preprocess(m[None, ...], return_tensors="pt")["pixel_values"][0]
                          for m in E.regions_for_line(json.loads(json.dumps(line)), False, False)]).float()
    image = Image.open(tmp_path / "a.jpg").convert("RGB")
    again = E.answer_questions(line, model, tok, proc, image, E.depth_image(np.array(image), depth_predictor), masks, "llava_v1", "tiny", "a.jpg")
    assert again[0]["pred"] == rec[0]["pred"]
