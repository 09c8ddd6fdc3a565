"""The SpatialRGPT-Bench driver (spatialrgpt_b200/eval_spatial.py, mirror of llava/eval/eval_spatial.py) on the CPU with a stub
model: chunking, region construction (run-length masks, box fallback, clamping, padding), prompt assembly, JSONL records."""
import json
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from spatialrgpt_b200 import eval_spatial as E
from spatialrgpt_b200.constants import IMAGE_TOKEN_INDEX
from tests.golden.make_host_golden import ToyTokenizer


def test_chunking_matches_the_reference_rule():
    items = list(range(10))
    assert [list(c) for c in E.split_list(items, 3)] == [[0, 1, 2, 3], [4, 5, 6, 7], [8, 9]]  # ceil(10/3) = 4 per chunk
    assert list(E.get_chunk(items, 4, 3)) == [9] and list(E.get_chunk(items, 1, 0)) == items
    assert sum(len(c) for c in E.split_list(items, 8)) == 10


def test_rle_roundtrip_and_column_major_order():
    rng = np.random.RandomState(0)
    for h, w in [(5, 7), (64, 48), (1, 9), (33, 1)]:
        m = (rng.rand(h, w) > 0.6).astype(np.uint8)
        flat = m.T.reshape(-1)  # column-major
        counts, cur, run = [], 0, 0
        for v in flat:
            if v == cur:
                run += 1
            else:
                counts.append(run); cur ^= 1; run = 1
        counts.append(run)
        assert np.array_equal(E.rle_decode({"size": [h, w], "counts": counts}), m)
        s = E.rle_encode_counts(counts)
        assert all(48 <= ord(ch) < 48 + 64 for ch in s)
        assert np.array_equal(E.rle_decode({"size": [h, w], "counts": s}), m)
        assert np.array_equal(E.rle_decode({"size": [h, w], "counts": s.encode()}), m)
    # a hand-checked case: 2x3 mask, column-major runs 1 zero, 2 ones, 3 zeros -> [[0,1,0],[1,0,0]]
    assert E.rle_decode({"size": [2, 3], "counts": [1, 2, 3]}).tolist() == [[0, 1, 0], [1, 0, 0]]
    with pytest.raises(ValueError):
        E.rle_decode({"size": [2, 3], "counts": [1, 2]})
    # long runs need several characters and the delta coding of the compressed form
    counts = [0, 1000, 70000, 3, 5, 1200]
    assert E._rle_counts_from_string(E.rle_encode_counts(counts)) == counts


def test_regions_boxes_clamp_pad_and_fallback():
    info = {"height": 20, "width": 30, "file_path": "x.jpg"}
    line = {"image_info": info, "bbox": [[-5.0, 2.0, 10.0, 50.0], [3.2, 4.9, 8.1, 9.0]], "rle": [{"size": [20, 30], "counts": "garbage!"}]}
    boxes = E.regions_for_line(json.loads(json.dumps(line)), use_mask=False, pad=False)
    assert boxes[0].shape == (20, 30) and int(boxes[0].sum()) == 10 * 18 and int(boxes[1].sum()) == 5 * 5  # clamped; int() truncation
    fallback = E.regions_for_line(json.loads(json.dumps(line)), use_mask=True, pad=False)  # undecodable rle -> boxes, like the reference
    assert all(np.array_equal(a, b) for a, b in zip(boxes, fallback))
    padded = E.regions_for_line(json.loads(json.dumps(line)), use_mask=False, pad=True)
    assert padded[0].shape == (30, 30) and int(padded[0].sum()) == int(boxes[0].sum()) and padded[0][:5].sum() == 0
    good = dict(line, rle=[{"size": [20, 30], "counts": [40, 20, 540]}])
    m = E.regions_for_line(good, use_mask=True, pad=False)
    assert int(m[0].sum()) == 20 and m[0][0, 2] == 1  # the run starts at column 2, row 0


class _StubModel:
    def __init__(self):
        self.device = torch.device("cpu")
        self.config = SimpleNamespace(image_aspect_ratio="resize")
        self.calls = []
        self.dtype = torch.float16  # what the loader hands out (builder.py:62)

    def to(self, dtype=None, **kw):
        self.dtype = dtype or self.dtype
        return self

    def generate(self, input_ids, images=None, depths=None, masks=None, **kw):
        self.calls.append(dict(ids=input_ids.clone(), images=images, depths=depths, masks=masks, kw=kw))
        return torch.tensor([[5, 6, 7]])


def test_driver_writes_the_reference_records(tmp_path):
    from PIL import Image
    from transformers import SiglipImageProcessor
    proc = SiglipImageProcessor(size={"height": 28, "width": 28})
    tok = ToyTokenizer()
    tok.batch_decode = lambda ids, skip_special_tokens=True: ["  the answer is 2 m</s>"]
    model = _StubModel()
    Image.fromarray(np.random.RandomState(1).randint(0, 255, (20, 30, 3), dtype=np.uint8)).save(tmp_path / "a.jpg")
    ann = [{"id": i, "image_info": {"file_path": "a.jpg", "height": 20, "width": 30}, "text_q": "how far?", "qa_info": {"type": "dist"},
            "bbox": [[1, 1, 10, 10], [5, 5, 25, 18]],
            "conversations": [{"from": "human", "value": "<image>\nDistance between <mask> and <mask>?"}, {"from": "gpt", "value": "2 m"},
                              {"from": "human", "value": "And is <mask> closer?"}, {"from": "gpt", "value": "yes"}]} for i in range(3)]
    (tmp_path / "ann.json").write_text(json.dumps(ann))
    args = SimpleNamespace(model_path="ckpt/SpatialRGPT-VILA1.5-8B", model_base=None, image_folder=str(tmp_path), annotation_file=str(tmp_path / "ann.json"),
                           answers_file=str(tmp_path / "out" / "answers.jsonl"), conv_mode="llava_v1", num_chunks=2, chunk_idx=0, temperature=0.0,
                           top_p=None, num_beams=1, use_mask=False)
    n = E.eval_model(args, depth_predictor=None, loader=lambda p, name, base: (tok, model, proc, 4096))
    recs = [json.loads(l) for l in open(args.answers_file)]
    assert n == len(recs) == 4  # chunk 0 of 2 holds ceil(3/2) = 2 annotations x 2 question turns
    assert recs[0] == {"question_id": 0, "image": "a.jpg", "question": "how far?", "pred": "the answer is 2 m", "gt": "2 m",
                       "model_id": "SpatialRGPT-VILA1.5-8B", "qa_info": {"type": "dist"}}
    assert recs[1]["gt"] == "yes" and recs[2]["question_id"] == 1
    c0, c1 = model.calls[0], model.calls[1]
    assert c0["images"].shape == (1, 3, 28, 28) and c0["images"].dtype == torch.bfloat16 and c0["depths"] is None
    assert c0["masks"][0].shape == (2, 28, 28) and c0["kw"]["do_sample"] is False and c0["kw"]["max_new_tokens"] == 128
    assert int((c0["ids"] == IMAGE_TOKEN_INDEX).sum()) == 1
    assert c1["ids"].shape[1] > c0["ids"].shape[1]  # the second turn carries the first question and an empty answer slot
    assert E.question_with_depth_tokens("a <mask> b <mask>") == "a <mask> <depth> b <mask> <depth>"
    assert E.stop_string("llava_v1") == "</s>" and E.clean_output(" x </s>", "</s>") == "x"


# ---- region classification driver (spatialrgpt_b200/eval_region_cls.py, mirror of llava/eval/eval_region_cls.py) -------------------
def test_region_cls_crop_box_matches_the_reference_rule():
    from spatialrgpt_b200 import eval_region_cls as R
    info = {"height": 100, "width": 160}
    assert R.get_crop_box([[0, 0, 150, 40]], info) == [0, 0, 160, 100]            # wider than the short side: whole image
    assert R.get_crop_box([[10, 10, 30, 40]], info) == [0, 0, 100, 100]           # window clipped at the left / top edge
    # expected values = outputs of the reference's own get_crop_box (eval_region_cls.py:49-72) on these inputs; note its quirk: the
    # right edge is compared with the SHORT side, so any window reaching past x = 100 snaps to the right border
    assert R.get_crop_box([[60, 30, 100, 70]], info) == [60, 0, 160, 100]
    assert R.get_crop_box([[120, 30, 150, 70]], info) == [60, 0, 160, 100]


def test_region_cls_driver_end_to_end_with_a_stub_model(tmp_path):
    from PIL import Image
    from transformers import SiglipImageProcessor

    from spatialrgpt_b200 import eval_region_cls as R
    Image.fromarray(np.random.RandomState(2).randint(0, 255, (60, 90, 3), dtype=np.uint8)).save(tmp_path / "img1.jpg")
    os.makedirs(tmp_path / "coco" / "val2017")
    os.replace(tmp_path / "img1.jpg", tmp_path / "coco" / "val2017" / "img1.jpg")
    coco = {"images": [{"id": 5, "height": 60, "width": 90, "coco_url": "http://x/val2017/img1.jpg"}],
            "categories": [{"id": 1, "name": "Dog"}, {"id": 2, "name": "cat"}],
            "annotations": [{"id": 1, "image_id": 5, "category_id": 1, "iscrowd": 0, "bbox": [10, 5, 30, 40], "segmentation": [[12, 8, 38, 8, 38, 40, 12, 40]]},
                            {"id": 2, "image_id": 5, "category_id": 2, "iscrowd": 1, "bbox": [0, 0, 5, 5], "segmentation": [[0, 0, 4, 0, 4, 4]]},
                            {"id": 3, "image_id": 5, "category_id": 2, "iscrowd": 0, "bbox": [50, 20, 20, 20],
                             "segmentation": {"size": [60, 90], "counts": [60 * 50 + 20, 20, 60 * 90 - 60 * 50 - 40]}}]}
    (tmp_path / "ann.json").write_text(json.dumps(coco))
    data = R.generate_data_list(str(tmp_path / "ann.json"))
    assert [d["category_name"] for d in data] == ["dog", "cat"] and data[0]["bbox"] == [[10, 5, 40, 45]] and data[0]["image"] == os.path.join("coco", "val2017", "img1.jpg")
    poly = R.segmentation_to_mask(data[0]["segmentation"][0], 60, 90)
    assert poly.shape == (60, 90) and poly[20, 20] == 1 and poly[2, 2] == 0 and 800 < int(poly.sum()) < 1000
    assert int(R.segmentation_to_mask(data[1]["segmentation"][0], 60, 90).sum()) == 20

    proc = SiglipImageProcessor(size={"height": 28, "width": 28})
    tok = ToyTokenizer()
    seen = []

    class Stub:
        device = torch.device("cpu")
        dtype = torch.float16  # this script runs the model as loaded (eval_region_cls.py:316-317)
        config = SimpleNamespace(image_aspect_ratio="resize", mm_use_im_start_end=False)

        def generate(self, input_ids, images=None, masks=None, **kw):
            assert images.dtype == torch.float16 and masks[0].dtype == torch.float16
            seen.append((input_ids.clone(), tuple(images.shape), tuple(masks[0].shape), kw))
            return torch.tensor([[tok._id("dog"), tok._id("</s>")]])

    args = SimpleNamespace(model_path="m/tiny-cls", model_base=None, image_folder=str(tmp_path), annotation_file=str(tmp_path / "ann.json"),
                           answers_file=str(tmp_path / "out" / "ans.jsonl"), conv_mode="llava_v1", num_chunks=1, chunk_idx=0, temperature=0.0, top_p=None,
                           num_beams=1, dataset="coco", prompt_type="seg")
    n = R.eval_model(args, loader=lambda p, name, base: (tok, Stub(), proc, 2048), seed=0)
    rec = [json.loads(l) for l in open(args.answers_file)]
    assert n == 2 and [r["gt_name"] for r in rec] == ["dog", "cat"] and rec[0]["text"] == "dog" and rec[0]["model_id"] == "tiny-cls"
    assert rec[0]["question_id"] == data[0]["image"] and rec[0]["image_id"] == 5 and rec[1]["bbox"] == [[50, 20, 70, 40]]
    ids, img_shape, mask_shape, kw = seen[0]
    assert img_shape == (1, 3, 28, 28) and mask_shape == (1, 28, 28) and int((ids == IMAGE_TOKEN_INDEX).sum()) == 1
    assert kw["max_new_tokens"] == 64 and kw["do_sample"] is False
    # the box prompt type rasterises the box inside the same crop window
    args.prompt_type = "box"
    assert R.eval_model(args, loader=lambda p, name, base: (tok, Stub(), proc, 2048), seed=0) == 2


# ---- multi-turn region chat (spatialrgpt_b200/chat.py, the follow-up flow of demo/gradio_web_server_multi.py:137-236) ---------------
def test_region_chat_follow_up_flow_with_a_stub_model():
    from PIL import Image
    from transformers import SiglipImageProcessor

    from spatialrgpt_b200.chat import RegionChat
    proc = SiglipImageProcessor(size={"height": 28, "width": 28})
    tok = ToyTokenizer()
    calls = []
    answers = iter(["Region [0] is behind Region [1] </s>", "It is [0] </s>"])

    class Stub:
        device = torch.device("cpu")
        dtype = torch.bfloat16
        config = SimpleNamespace(image_aspect_ratio="resize", mm_use_im_start_end=False)

        def generate(self, input_ids, images=None, depths=None, masks=None, **kw):
            calls.append((input_ids.clone(), masks[0].clone(), depths is not None, kw))
            return torch.tensor([tok(next(answers)).input_ids[1:]])

    img = Image.fromarray(np.random.RandomState(0).randint(0, 255, (40, 50, 3), dtype=np.uint8))
    depth = Image.fromarray(np.random.RandomState(1).randint(0, 255, (40, 50, 3), dtype=np.uint8))
    segs = [np.zeros((40, 50), dtype=np.uint8) for _ in range(4)]
    for i, s in enumerate(segs):
        s[5 * i:5 * i + 8, 4:20] = 1
    chat = RegionChat(Stub(), tok, proc, conv_mode="llava_v1")
    a1 = chat.ask("Is <region2> behind <region0> ?", img, segs, depth_image=depth)
    assert a1 == "Region [2] is behind Region [0]"            # [k] = k-th region of the turn -> the user's region number
    ids1, masks1, had_depth, kw = calls[0]
    assert had_depth and masks1.shape[0] == 2 and int((ids1 == IMAGE_TOKEN_INDEX).sum()) == 1
    assert kw["stopping_criteria"] and kw["do_sample"] is False
    a2 = chat.ask("And how wide is <region3> ?", img, segs, depth_image=depth, follow_up=True)
    assert a2 == "It is [3]"
    ids2, masks2, _, _ = calls[1]
    assert masks2.shape[0] == 3                                # the masks of ALL region references so far: regions 2, 0, 3
    full = __import__("spatialrgpt_b200.mm_utils", fromlist=["process_regions"]).process_regions(segs, proc, Stub.config)
    assert torch.equal(masks2.float(), full[[2, 0, 3]].to(torch.bfloat16).float())
    assert ids2.shape[1] > ids1.shape[1] and int((ids2 == IMAGE_TOKEN_INDEX).sum()) == 1  # the conversation grew, one image token
    assert chat.conv.messages[1][1] == "Region [0] is behind Region [1]" and len(chat.conv.messages) == 4
    # a new first turn resets the session
    answers2 = iter(["ok </s>"])
    Stub.generate = lambda self, input_ids, images=None, depths=None, masks=None, **kw: torch.tensor([tok(next(answers2)).input_ids[1:]])
    assert chat.ask("What is <region1> ?", img, segs) == "ok" and len(chat.user_turns) == 1
