"""Host-side API surface vs fixtures produced by the reference's llava.conversation / llava.mm_utils."""
import json
import os

import numpy as np
import pytest
import torch

from spatialrgpt_b200 import conversation as C
from spatialrgpt_b200 import mm_utils as M
from spatialrgpt_b200.constants import IMAGE_TOKEN_INDEX
from tests.golden.make_host_golden import CONVERSATIONS, PROMPTS, ToyTokenizer


@pytest.fixture(scope="module")
def gold(golden_dir):
    return json.load(open(os.path.join(golden_dir, "host_api.json")))


def test_every_reference_template_renders_identically(gold):
    assert set(gold["prompts"]) <= set(C.conv_templates)
    for name, rows in gold["prompts"].items():
        for conv_msgs, expect in zip(CONVERSATIONS, rows):
            c = C.conv_templates[name].copy()
            try:
                for who, msg in conv_msgs:
                    c.append_message(c.roles[0] if who == "U" else c.roles[1], tuple(msg) if isinstance(msg, list) else msg)
                got = c.get_prompt()
            except Exception as e:
                got = f"raises {type(e).__name__}"
            assert got == expect, f"template {name}: {got!r} != {expect!r}"


def test_llama3_and_v1_stop_strings():
    # eval_spatial.py:215-219: stop_str = sep unless SeparatorStyle.TWO -> sep2
    l3, v1 = C.conv_templates["llama_3"], C.conv_templates["v1"]
    assert l3.sep == "<|eot_id|>" and l3.sep_style == C.SeparatorStyle.LLAMA_3
    assert v1.sep2 == "</s>" and v1.sep_style == C.SeparatorStyle.TWO
    c = l3.copy()
    c.append_message(c.roles[0], "hi")
    c.append_message(c.roles[1], None)
    assert c.get_prompt().endswith("<|start_header_id|>assistant<|end_header_id|>\n\n")
    assert C.conv_templates["llama_3"].messages == []  # copy() does not alias the template's history


def test_tokenizer_image_token(gold):
    for row in gold["tokenize"]:
        ids = M.tokenizer_image_token(row["prompt"], ToyTokenizer(), lstrip=row["lstrip"])
        assert ids == row["ids"], row
    t = M.tokenizer_image_token(PROMPTS[0], ToyTokenizer(), return_tensors="pt")
    assert t.dtype == torch.long and int((t == IMAGE_TOKEN_INDEX).sum()) == 1
    with pytest.raises(ValueError):
        M.tokenizer_image_token("x", ToyTokenizer(), return_tensors="np")


def test_keywords_stopping_criteria(gold):
    tok = ToyTokenizer()
    base = tok("the red chair is left of the table </s> extra").input_ids
    assert base == gold["stopping_ids"]
    crit = M.KeywordsStoppingCriteria(["</s>"], tok, torch.zeros(1, 0, dtype=torch.long))
    got = [bool(crit(torch.tensor([base[:n]]), None)) for n in range(1, len(base) + 1)]
    assert got == gold["stopping"]


def test_model_name_from_path(gold):
    for p, name in gold["model_names"].items():
        assert M.get_model_name_from_path(p) == name


def test_process_images_and_regions_shapes():
    from PIL import Image
    from transformers import SiglipImageProcessor
    from types import SimpleNamespace

    proc = SiglipImageProcessor(size={"height": 56, "width": 56})
    cfg = SimpleNamespace(image_aspect_ratio="resize", image_processor=proc)
    rng = np.random.RandomState(0)
    imgs = [Image.fromarray(rng.randint(0, 255, (40, 70, 3), dtype=np.uint8)) for _ in range(2)]
    px = M.process_images(imgs, proc, cfg)
    assert px.shape == (2, 3, 56, 56) and px.dtype == torch.float32
    assert float(px.min()) >= -1.0 and float(px.max()) <= 1.0  # rescale 1/255, mean = std = 0.5
    masks = M.boxes_to_masks([[5, 5, 30, 20], [-3, 0, 100, 100]], 40, 70)
    assert masks[1].sum() == 40 * 70 and masks[0].sum() == 25 * 15
    reg = M.process_regions(masks, proc, cfg)
    assert reg.shape == (2, 56, 56) and reg.dtype == torch.float32
    assert 0.9 < float(reg[1].mean()) <= 1.01  # resampled floats, not rescaled by 1/255 (mm_utils.py:479-482)


@pytest.mark.parametrize("mode", ["resize", "pad"])
def test_process_images_and_regions_match_reference(mode):
    """a1: process_images / process_regions against outputs of the REFERENCE's llava/mm_utils.py:421-542 on the same seeded
    inputs (tests/golden/host_preproc.npz, written by make_host_golden.py).  Same PIL / cv2 / HF processor underneath -> exact."""
    from types import SimpleNamespace
    from transformers import SiglipImageProcessor

    from tests.golden.make_host_golden import preproc_inputs

    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "host_preproc.npz"))
    imgs, masks = preproc_inputs()
    proc = SiglipImageProcessor(size={"height": 56, "width": 56})
    cfg = SimpleNamespace(image_aspect_ratio=mode, image_processor=proc)
    px = M.process_images(imgs, proc, cfg)
    assert px.dtype == torch.float32 and np.array_equal(px.numpy(), gold[f"images_{mode}"])
    for tag, sl in (("a", slice(0, 2)), ("b", slice(2, 3))):
        reg = M.process_regions(masks[sl], proc, cfg)
        assert reg.dtype == torch.float32 and np.array_equal(reg.numpy(), gold[f"regions_{mode}_{tag}"])


def test_read_checkpoint_roundtrip(tmp_path):
    """A synthetic checkpoint written in the reference's four-directory layout parses back (CPU only)."""
    from safetensors.torch import save_file

    from oracle import srgpt_oracle as O
    from spatialrgpt_b200 import builder
    from tests.golden.make_golden import CASES

    oc = O.OracleConfig(**CASES["tiny_boxes"][0])
    sd = O.make_weights(oc, seed=1)
    root = str(tmp_path / "ckpt")
    top = {"architectures": ["LlavaLlamaModel"], "model_type": "llava_llama", "enable_region": True, "enable_depth": True,
           "mm_vision_select_layer": -2, "mm_vision_select_feature": "cls_patch", "mm_use_im_patch_token": False,
           "image_aspect_ratio": "resize", "llm_cfg": {}, "vision_tower_cfg": {}, "mm_projector_cfg": {}, "region_extractor_cfg": {}}
    subs = {
        "llm": {"hidden_size": oc.hidden, "num_hidden_layers": oc.layers, "num_attention_heads": oc.heads,
                "num_key_value_heads": oc.kv_heads, "head_dim": oc.head_dim, "intermediate_size": oc.inter, "vocab_size": oc.vocab,
                "rope_theta": oc.rope_theta, "rms_norm_eps": oc.rms_eps, "max_position_embeddings": 4096},
        "vision_tower": {"image_size": oc.image_size, "patch_size": 14, "hidden_size": oc.v_hidden, "num_hidden_layers": oc.v_layers,
                         "num_attention_heads": oc.v_heads, "intermediate_size": oc.v_inter, "layer_norm_eps": 1e-6},
        "mm_projector": {"mm_projector_type": "mlp_downsample"},
        "region_extractor": {"region_extractor_type": "regiongpt"},
    }
    os.makedirs(root)
    json.dump(top, open(os.path.join(root, "config.json"), "w"))
    for name, cfg in subs.items():
        d = os.path.join(root, name)
        os.makedirs(d)
        json.dump(cfg, open(os.path.join(d, "config.json"), "w"))
        tensors = {k: v.contiguous() for k, v in sd[name].items()}
        if name == "llm":  # sharded, like a real 8B checkpoint
            keys = sorted(tensors)
            save_file({k: tensors[k] for k in keys[::2]}, os.path.join(d, "model-00001-of-00002.safetensors"))
            save_file({k: tensors[k] for k in keys[1::2]}, os.path.join(d, "model-00002-of-00002.safetensors"))
        else:
            save_file(tensors, os.path.join(d, "model.safetensors"))
    assert builder.is_mm_model(root)
    cfg, got, tok, proc = builder.read_checkpoint(root, load_tokenizer=False)
    assert cfg.enable_region and cfg.enable_depth and not cfg.mm_use_im_patch_token
    assert cfg.llama.hidden_size == oc.hidden and cfg.llama.num_key_value_heads == oc.kv_heads and cfg.llama.rope_theta == oc.rope_theta
    assert cfg.vision.image_size == oc.image_size and cfg.vision.intermediate_size == oc.v_inter
    for part in ("llm", "vision_tower", "mm_projector", "region_extractor"):
        assert set(got[part]) == set(sd[part])
        for k in sd[part]:
            assert torch.equal(got[part][k], sd[part][k])
    # resize_token_embeddings semantics (builder.py:199)
    builder._resize_token_embeddings(cfg, got["llm"], oc.vocab + 3)
    assert got["llm"]["model.embed_tokens.weight"].shape[0] == oc.vocab + 3 and cfg.llama.vocab_size == oc.vocab + 3
    assert torch.equal(got["llm"]["lm_head.weight"][: oc.vocab], sd["llm"]["lm_head.weight"])
    with pytest.raises(NotImplementedError):
        builder.load_pretrained_model(root, "x", load_4bit=True)


def test_read_checkpoint_with_tokenizer_registers_special_tokens(tmp_path):
    """builder.py:186-199 on a synthetic four-directory checkpoint WITH tokenizer / processor files: <mask> and <depth> are added as
    special tokens and their ids recorded on the config, the token tables are resized to len(tokenizer), the image processor comes
    from vision_tower/, and the stop ids come from llm/generation_config.json (a list for Llama-3 style checkpoints)."""
    from oracle import srgpt_oracle as O
    from spatialrgpt_b200 import builder
    from tests.golden.make_golden import CASES
    from tests.util import write_synthetic_checkpoint

    oc = O.OracleConfig(**CASES["tiny_boxes"][0])
    sd = O.make_weights(oc, seed=1)
    root = str(tmp_path / "ckpt")
    n_vocab = write_synthetic_checkpoint(root, oc, sd, generation_eos=[2, 7])
    cfg, got, tok, proc = builder.read_checkpoint(root)
    assert tok is not None and proc is not None and proc.size["height"] == oc.image_size
    assert cfg.llm_mask_token_id == n_vocab and cfg.llm_depth_token_id == n_vocab + 1 and len(tok) == n_vocab + 2
    assert tok.convert_tokens_to_ids("<mask>") == n_vocab and tok("<mask> <depth>").input_ids[-2:] == [n_vocab, n_vocab + 1]
    assert cfg.llama.vocab_size == len(tok) and got["llm"]["model.embed_tokens.weight"].shape[0] == len(tok)
    assert torch.equal(got["llm"]["lm_head.weight"], sd["llm"]["lm_head.weight"][: len(tok)])  # the tables shrink from 512 rows
    assert cfg.llama.eos_token_id == [2, 7]
    assert not cfg.mm_use_im_patch_token and tok.convert_tokens_to_ids("<im_patch>") in (None, tok.unk_token_id)


def test_process_masks_and_process_depth_match_reference():
    """mm_utils.py:279-418 (the dataset-side region / depth preparation) against outputs of the reference's own functions
    (tests/golden/host_masks_depth.npz by make_host_golden.py): box, box + external image_info and run-length regions, a depth image,
    both aspect modes; the modality draw uses the global ``random`` generator like the reference."""
    import random
    from types import SimpleNamespace

    from transformers import SiglipImageProcessor

    from spatialrgpt_b200 import mm_utils as M
    from tests.golden.make_host_golden import masks_depth_inputs
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "host_masks_depth.npz"))
    info, boxes, rles, depth = masks_depth_inputs()
    for mode in ("resize", "pad"):
        proc = SiglipImageProcessor(size={"height": 56, "width": 56})
        if getattr(proc, "crop_size", None) is None and "crop_size" in vars(proc):
            delattr(proc, "crop_size")
        cfg = SimpleNamespace(image_aspect_ratio=mode, image_processor=proc)
        assert np.array_equal(M.process_masks([{"bbox": boxes, "image_info": info}], cfg).numpy(), g[f"masks_bbox_{mode}"])
        assert np.array_equal(M.process_masks([{"bbox": boxes}], cfg, image_info=info).numpy(), g[f"masks_bbox_info_{mode}"])
        assert np.array_equal(M.process_masks([{"rle": rles}], cfg).numpy(), g[f"masks_rle_{mode}"])
        assert np.array_equal(M.process_depth(depth, cfg, None).numpy(), g[f"depth_{mode}"])  # same PIL / HF processor underneath -> exact
        # both modalities present: one is drawn with random.choice (seeded here); either way the result is one of the two fixtures
        random.seed(3)
        both = M.process_masks([{"bbox": boxes[:2], "rle": rles, "image_info": info}], cfg).numpy()
        assert np.array_equal(both, g[f"masks_rle_{mode}"]) or np.array_equal(both, g[f"masks_bbox_{mode}"][:2])
    with pytest.raises(NotImplementedError):
        M.process_masks([{"segmentation": [[[1, 1, 5, 1, 5, 5]]], "image_info": info}], cfg)


def test_small_image_helpers():
    """expand2square (mm_utils.py:249-276, incl. mode "L"), load_image_from_base64 (245-246), is_gemma_tokenizer (573-574)."""
    import base64
    from io import BytesIO

    from PIL import Image

    from llava.mm_utils import expand2square, is_gemma_tokenizer, load_image_from_base64
    im = Image.fromarray(np.full((4, 10, 3), 200, dtype=np.uint8))
    sq = expand2square(im, (10, 20, 30))
    a = np.asarray(sq)
    assert sq.size == (10, 10) and (a[:3] == (10, 20, 30)).all() and (a[3:7] == 200).all() and (a[7:] == (10, 20, 30)).all()
    tall = np.asarray(expand2square(Image.fromarray(np.full((9, 4), 7, dtype=np.uint8), mode="L"), (99, 1, 1)))
    assert tall.shape == (9, 9) and (tall[:, :2] == 99).all() and (tall[:, 2:6] == 7).all() and (tall[:, 6:] == 99).all()
    assert expand2square(sq, (0, 0, 0)) is sq
    buf = BytesIO()
    im.save(buf, format="PNG")
    back = load_image_from_base64(base64.b64encode(buf.getvalue()))
    assert np.array_equal(np.asarray(back), np.asarray(im))
    assert is_gemma_tokenizer(type("GemmaTokenizerFast", (), {})()) and not is_gemma_tokenizer(ToyTokenizer())


def test_rope_scaling_config_is_read_like_the_reference():
    """modeling_llama.py:267-292 (rope_scaling type linear / dynamic / unknown) and the loader's own context extension
    (language_model/builder.py:31-38: model_max_length > max_position_embeddings -> linear, factor = ceil(ratio))."""
    from spatialrgpt_b200.builder import _llama_dims
    base = {"hidden_size": 64, "num_hidden_layers": 1, "num_attention_heads": 2, "intermediate_size": 128, "vocab_size": 100,
            "max_position_embeddings": 4096}
    assert _llama_dims(base).rope_scaling_factor == 1.0
    assert _llama_dims({**base, "rope_scaling": {"type": "linear", "factor": 2.0}}).rope_scaling_factor == 2.0
    assert _llama_dims({**base, "rope_scaling": {"type": "dynamic", "factor": 2.0}}).rope_scaling_factor == 1.0
    assert _llama_dims({**base, "model_max_length": 10000}).rope_scaling_factor == 3.0
    assert _llama_dims({**base, "model_max_length": 4096}).rope_scaling_factor == 1.0
    with pytest.raises(ValueError):
        _llama_dims({**base, "rope_scaling": {"type": "yarn", "factor": 2.0}})


def test_unsupported_tower_families_raise_at_the_loader():
    from spatialrgpt_b200.builder import _vision_config
    base = {"image_size": 448, "patch_size": 14, "hidden_size": 64, "num_hidden_layers": 2, "num_attention_heads": 2, "intermediate_size": 128}
    assert not _vision_config({**base, "architectures": ["SiglipVisionModel"]}).is_clip
    assert _vision_config({**base, "architectures": ["CLIPVisionModel"]}).is_clip
    for arch in ("InternVisionModel", "RADIOModel"):
        with pytest.raises(NotImplementedError):
            _vision_config({**base, "architectures": [arch]})


def test_clip_checkpoint_reads_back_as_a_clip_tower(tmp_path):
    """multimodal_encoder/builder.py:38-47 picks the tower class from the vision config's architecture name; a checkpoint whose
    vision_tower/ holds a CLIPVisionModel parses as a CLIP tower ("patch" select, quick_gelu, class token + pre_layrnorm weights)."""
    from oracle import srgpt_oracle as O
    from spatialrgpt_b200 import builder
    from spatialrgpt_b200.weights import from_state_dicts
    from tests.golden.make_golden import CLIP_CASE
    from tests.util import write_synthetic_checkpoint

    oc = O.OracleConfig(**CLIP_CASE)
    sd = O.make_weights(oc, seed=2)
    root = str(tmp_path / "ckpt_clip")
    write_synthetic_checkpoint(root, oc, sd)
    cfg, got, tok, proc = builder.read_checkpoint(root)
    v = cfg.vision
    assert v.is_clip and v.hidden_act == "quick_gelu" and v.layer_norm_eps == 1e-5 and v.tokens == v.grid ** 2 + 1
    assert cfg.mm_vision_select_feature == "patch"
    w = from_state_dicts(cfg, got, "cpu")
    assert w.vision.patch_b is None and w.vision.cls_emb.shape == (oc.v_hidden,) and w.vision.pos_emb.shape == (v.tokens, oc.v_hidden)
    assert torch.equal(w.vision.pre_ln_w, sd["vision_tower"]["vision_model.pre_layrnorm.weight"])
    assert len(w.vision.layers) == oc.v_layers


def test_bench_algorithmic_numbers_match_the_survey():
    """bench.py's FLOP / byte model of config c2 equals SURVEY.md §8(d): 5.58 TFLOP per request to the first token,
    15.01 GB streamed per decoded token, 131072 B of KV per cached token, 234.9 MB for the gate/up GEMV."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    from spatialrgpt_b200 import baseline_config
    n = bench.algorithmic_numbers(baseline_config("c2"))
    assert n["S"] == 259
    assert abs(n["flops_ttft"] / 1e12 - 5.58) < 0.01
    assert abs(n["w_stream"] / 1e9 - 15.01) < 0.01
    assert n["kv_per_tok"] == 131072 and n["gateup_bytes"] == 2 * 14336 * 4096 * 2
    hbm, tensor, src, tensor_sustained = bench.load_peaks()
    assert 0 < tensor_sustained <= tensor
    assert hbm > 1000 and tensor > 100 and isinstance(src, str)


def test_ab_tool_parsing_and_summary():
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("ab_tool", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "ab.py"))
    ab = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ab)
    assert ab.parse_variant("B:SRGPT_X=2,Y=z") == ("B", {"SRGPT_X": "2", "Y": "z"}) and ab.parse_variant("A:") == ("A", {})
    out = 'noise\n{"kernel": "k1", "ms_median": 1.5}\n{"other": 1}\n{"kernel": "k2", "ms_median": 0.5, "x": 1}\n{broken'
    assert ab.collect(out) == {"k1": 1.5, "k2": 0.5}
    rows = ab.summarise({"A": {"k1": [1.0, 3.0, 2.0]}, "B": {"k1": [1.0, 1.0, 4.0]}})
    assert rows[0][0] == "k1" and rows[0][1] == {"A": 2.0, "B": 1.0} and rows[0][2]["B"] == 0.5
