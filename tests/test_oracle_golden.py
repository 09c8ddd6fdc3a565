"""Pin the CPU oracle (oracle/srgpt_oracle.py) against the committed golden fixtures, which were
produced by the REFERENCE's own modules (tests/golden/make_golden.py).  CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle import srgpt_oracle as O
from tests.golden.make_golden import CASES
from tests.util import BF16_STAGE, assert_close, load_npz


def _load(golden_dir, name):
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    return {k: torch.from_numpy(z[k]) for k in z.files}


def _close(a, b, rtol=2e-4, atol=2e-5):
    a, b = a.float(), b.float()
    assert a.shape == b.shape, (a.shape, b.shape)
    err = (a - b).abs().max().item()
    assert torch.allclose(a, b, rtol=rtol, atol=atol), f"max abs err {err}"


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_matches_reference_pipeline(golden_dir, name):
    kw, n_regions, t_text, kind, n_new, depth_on = CASES[name]
    g = _load(golden_dir, name)
    cfg = O.OracleConfig(**kw)
    weights = O.make_weights(cfg, seed=int(g["weight_seed"]))
    input_ids, images, depths, masks = O.synth_request(cfg, n_regions, t_text, seed=1234, kind=kind)
    # the synthetic request itself must be reproducible (it is regenerated on the GPU box)
    assert torch.equal(input_ids, g["input_ids"])
    _close(images, g["images"], 0, 0)
    _close(masks[0], g["masks"], 0, 0)
    if not depth_on:
        depths = None
    ids, enc = O.generate(cfg, weights, input_ids, images, depths, masks, n_new, return_all=True)
    _close(enc["tower_features"], g["tower_features"])
    _close(enc["hres"], g["hres"])
    _close(enc["lres"], g["lres"])
    _close(enc["mask_embeds"][0], g["mask_embeds"])
    if depth_on:
        _close(enc["depth_features"], g["depth_features"])
        _close(enc["depth_embeds"][0], g["depth_embeds"])
    _close(enc["image_features"], g["image_features"])
    _close(enc["inputs_embeds"], g["inputs_embeds"][0])
    _close(enc["logits"], g["logits"], rtol=1e-3, atol=1e-4)
    assert ids.tolist() == g["new_ids"].tolist()


@pytest.mark.parametrize("tag", ["rgb448", "depth448", "rgb384", "odd336"])
def test_mask_pooling_kats(golden_dir, tag):
    g = _load(golden_dir, "op_kats")
    x = g[f"{tag}_x"]
    masks = g[f"{tag}_masks"].float()
    out32 = O.mask_pooling(x.float(), [masks])[0]
    _close(out32, g[f"{tag}_out_f32"], rtol=1e-5, atol=1e-6)
    out16 = O.mask_pooling(x.to(torch.bfloat16), [masks])[0]
    assert out16.dtype == torch.bfloat16
    assert torch.equal(out16.float(), g[f"{tag}_out_bf16"])  # bit-exact: same torch ops, same roundings
    assert O.mask_pooling(x.float(), [None]) == [None]
    assert O.mask_pooling(x.float(), None) == [None]


def test_downsample_and_ln2d_kats(golden_dir):
    g = _load(golden_dir, "op_kats")
    assert torch.equal(O.downsample_block(g["downsample_x"]), g["downsample_out"])
    _close(O.layernorm2d(g["ln2d_x"], g["ln2d_w"], g["ln2d_b"]), g["ln2d_out"], 1e-6, 1e-6)


def test_bf16_mode_tracks_fp32():
    """The rounding-faithful bf16 mode stays within bf16 noise of the fp32 ground truth."""
    kw = CASES["tiny_boxes"][0]
    cfg = O.OracleConfig(**kw)
    w = O.make_weights(cfg, seed=3)
    ids, im, de, ma = O.synth_request(cfg, 2, 24, kind="box")
    a = O.encode_multimodal(cfg, w, im, de, ma, torch.float32)
    b = O.encode_multimodal(cfg, w, im, de, ma, torch.bfloat16)
    for k in ("tower_features", "image_features"):
        ref = a[k].float()
        err = (b[k].float() - ref).abs().max() / ref.abs().max()
        assert err < 0.05, (k, err)


def test_depth_to_u8x3():
    d = torch.linspace(0, 1, 12).view(1, 3, 4)
    out = O.depth_to_u8x3(d, 6, 8)
    assert out.shape == (6, 8, 3) and out.dtype == torch.uint8
    assert out.min() == 0 and out.max() == 255
    assert torch.equal(out[..., 0], out[..., 2])


def test_beam_search_matches_hf_generate(golden_dir):
    """The oracle's restatement of HF beam search against ``LlamaForCausalLM.generate(inputs_embeds=..., num_beams=k)`` on the same
    seeded weights (fixture by ``make_golden.py beam``): plain beams, EOS-closed hypotheses, length penalties, early stopping."""
    from tests.golden.make_golden import BEAM_CASES, CASES as C2
    g = load_npz(os.path.join(golden_dir, "beam_kats.npz"))
    oc = O.OracleConfig(**C2["tiny_masks_gqa"][0])
    sd = O.make_weights(oc, seed=int(g["weight_seed"]))
    for i, (nb, eos, n_new, lp, es) in enumerate(BEAM_CASES):
        ids = O.beam_search_generate(oc, sd["llm"], g["inputs_embeds"], nb, n_new, eos_token_id=eos, length_penalty=lp, early_stopping=es)
        ref = g[f"case{i}"].tolist()
        assert ids.tolist() == ref[: len(ids)], (i, ids.tolist(), ref)
        assert all(t == 0 for t in ref[len(ids):])  # HF pads a shorter best hypothesis with pad_token_id (0 in the fixture)
    assert O.beam_search_generate(oc, sd["llm"], g["inputs_embeds"], 1, 10).tolist() == O.greedy_generate(oc, sd["llm"], g["inputs_embeds"], 10).tolist()


def test_rope_tables_match_reference_rotary_classes(golden_dir):
    """cos / sin of LlamaRotaryEmbedding and LlamaLinearScalingRotaryEmbedding (modeling_llama.py:81-141; fixture by
    ``make_golden.py rope``): the oracle's restatement and the host-side table builder of the product path, fp32 and bf16."""
    from spatialrgpt_b200.config import LlamaDims
    from spatialrgpt_b200.llama_decoder import build_rope_tables
    g = load_npz(os.path.join(golden_dir, "rope_kats.npz"))
    pos = g["positions"].long()
    for name, factor in (("plain", 1.0), ("linear4", 4.0)):
        oc = O.OracleConfig(head_dim=128, rope_theta=500000.0, rope_scaling_factor=factor)
        cos, sin = O.rope_cos_sin(oc, pos, torch.float32)
        assert torch.equal(cos, g[name + "_cos"]) and torch.equal(sin, g[name + "_sin"])
        tc, ts = build_rope_tables(LlamaDims(head_dim=128, rope_theta=500000.0, rope_scaling_factor=factor), 4096, "cpu", torch.bfloat16)
        assert torch.equal(tc[pos].float(), g[name + "_cos"][:, :64].to(torch.bfloat16).float())
        assert torch.equal(ts[pos].float(), g[name + "_sin"][:, :64].to(torch.bfloat16).float())


def test_clip_tower_matches_reference_module(golden_dir):
    """The oracle's CLIP branch (clip_encoder.py:8-13 over HF CLIPVisionModel) against the reference's VisionTower.forward +
    feature_select("patch") run on the same seeded weights (fixture by ``make_golden.py clip``): the class token is dropped,
    hidden_states[-2] is selected, fp32 agrees to rounding and the bf16 / fp16 modes track it."""
    from tests.golden.make_golden import CLIP_CASE
    g = load_npz(os.path.join(golden_dir, "clip_tower.npz"))
    oc = O.OracleConfig(**CLIP_CASE)
    sd = O.make_weights(oc, seed=int(g["weight_seed"]))
    out = O.vision_tower_forward(oc, sd["vision_tower"], g["images"])
    assert out.shape == g["tower_features"].shape == (3, oc.grid ** 2, oc.v_hidden)
    assert torch.allclose(out, g["tower_features"], rtol=1e-5, atol=1e-5)
    assert_close(O.vision_tower_forward(oc, sd["vision_tower"], g["images"], torch.bfloat16), g["tower_features"], **BF16_STAGE, what="clip bf16")
    assert_close(O.vision_tower_forward(oc, sd["vision_tower"], g["images"], torch.float16), g["tower_features"], rel_rms=1.25e-2, rel_max=1.5e-1,
                 what="clip fp16")
    with pytest.raises(ValueError):
        O.vision_tower_forward(O.OracleConfig(**{**CLIP_CASE, "select_feature": "nope"}), sd["vision_tower"], g["images"])


@pytest.mark.parametrize("ptype", ["linear", "mlp2x_gelu", "mlp3x_gelu", "identity"])
def test_projector_types_match_reference_module(golden_dir, ptype):
    """The non-default mm_projector types against the reference's MultimodalProjector (fixture by make_golden.py proj)."""
    g = load_npz(os.path.join(golden_dir, "proj_kats.npz"))
    w = {k.split("__w__")[1]: v for k, v in g.items() if k.startswith(ptype + "__w__")}
    out = O.mm_projector_forward(O.OracleConfig(), w, g["x"], ptype=ptype)
    assert torch.allclose(out, g[ptype + "__out"], rtol=1e-5, atol=1e-6)
