"""Generate the committed golden fixtures by running the REFERENCE's own modules.

Run in the authoring container only (needs /root/reference):

    python tests/golden/make_golden.py

What runs (all unmodified reference / third-party code, fp32 on CPU):
  * ``llava.model.llava_arch.LlavaMetaForCausalLM.prepare_inputs_labels_for_multimodal``
  * ``llava.model.region_extractor.base_extractor.RegionExtractor`` (+ MaskPooling, LayerNorm2d)
  * ``llava.model.multimodal_projector.base_projector.MultimodalProjector`` (mlp_downsample)
  * ``llava.model.multimodal_encoder.vision_encoder.VisionTower`` over stock HF ``SiglipVisionModel``
  * stock HF ``LlamaForCausalLM.generate`` (greedy) started from ``inputs_embeds``
with the seeded synthetic weights of ``oracle.srgpt_oracle.make_weights``.  Outputs go to
``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` pins the oracle against them.
"""
from __future__ import annotations

import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))

from oracle import ref_shim  # noqa: E402
from oracle import srgpt_oracle as O  # noqa: E402

CASES = {
    # name: (OracleConfig kwargs, n_regions, t_text, kind, max_new_tokens, depth_on)
    "tiny_boxes": (dict(image_size=56, v_hidden=144, v_layers=3, v_heads=2, v_inter=296, hidden=256, layers=2,
                        heads=2, kv_heads=1, inter=384, vocab=512, rope_theta=10000.0, mask_token_id=510,
                        depth_token_id=511), 2, 24, "box", 8, True),
    "tiny_masks_gqa": (dict(image_size=112, v_hidden=144, v_layers=4, v_heads=2, v_inter=296, hidden=512, layers=3,
                            heads=4, kv_heads=2, inter=640, vocab=1003, rope_theta=500000.0, mask_token_id=1001,
                            depth_token_id=1002), 3, 32, "mask", 12, True),
    "tiny_nodepth": (dict(image_size=56, v_hidden=144, v_layers=3, v_heads=2, v_inter=296, hidden=256, layers=2,
                          heads=2, kv_heads=1, inter=384, vocab=512, rope_theta=10000.0, mask_token_id=510,
                          depth_token_id=511, enable_depth=False), 2, 24, "mask", 6, False),
}


def build_reference_vlm(cfg: O.OracleConfig, weights):
    """the reference classes wired together without from_pretrained."""
    ref_shim.install()
    import torch.nn as nn
    from transformers import LlamaConfig, LlamaForCausalLM, SiglipVisionConfig, SiglipVisionModel

    from llava.model.llava_arch import LlavaMetaForCausalLM
    from llava.model.multimodal_encoder.vision_encoder import VisionTower
    from llava.model.multimodal_projector.base_projector import MultimodalProjector, MultimodalProjectorConfig
    from llava.model.region_extractor.base_extractor import RegionExtractor, RegionExtractorConfig

    vcfg = SiglipVisionConfig(hidden_size=cfg.v_hidden, intermediate_size=cfg.v_inter, num_hidden_layers=cfg.v_layers,
                              num_attention_heads=cfg.v_heads, image_size=cfg.image_size, patch_size=cfg.patch_size,
                              layer_norm_eps=cfg.v_eps, hidden_act="gelu_pytorch_tanh")
    vcfg._attn_implementation = "eager"
    siglip = SiglipVisionModel(vcfg).float().eval()
    missing, unexpected = siglip.load_state_dict({k: v.float() for k, v in weights["vision_tower"].items()}, strict=False)
    assert not unexpected, unexpected
    assert all("head" in k or "post_layernorm" in k for k in missing), missing

    lcfg = LlamaConfig(hidden_size=cfg.hidden, intermediate_size=cfg.inter, num_hidden_layers=cfg.layers,
                       num_attention_heads=cfg.heads, num_key_value_heads=cfg.kv_heads, vocab_size=cfg.vocab,
                       rms_norm_eps=cfg.rms_eps, rope_theta=cfg.rope_theta, max_position_embeddings=4096,
                       tie_word_embeddings=False, head_dim=cfg.head_dim, attention_bias=False, mlp_bias=False,
                       bos_token_id=1, eos_token_id=None, pad_token_id=None)
    lcfg._attn_implementation = "eager"
    llm = LlamaForCausalLM(lcfg).float().eval()
    llm.load_state_dict({k: v.float() for k, v in weights["llm"].items()}, strict=True)

    mm_cfg = SimpleNamespace(mm_hidden_size=cfg.v_hidden, hidden_size=cfg.hidden)
    projector = MultimodalProjector(MultimodalProjectorConfig("mlp_downsample"), mm_cfg).float().eval()
    projector.load_state_dict({k: v.float() for k, v in weights["mm_projector"].items()}, strict=True)
    extractor = RegionExtractor(RegionExtractorConfig("regiongpt"), mm_cfg).float().eval()
    extractor.load_state_dict({k: v.float() for k, v in weights["region_extractor"].items()}, strict=True)

    class Tower(VisionTower):
        def __init__(self):
            super().__init__("synthetic", SimpleNamespace(mm_vision_select_layer=cfg.select_layer,
                                                          mm_vision_select_feature="cls_patch"))
            self.vision_tower = siglip
            self.is_loaded = True

        @property
        def dtype(self):
            return torch.float32

        @property
        def device(self):
            return torch.device("cpu")

        @property
        def config(self):
            return self._cfg

    tower = Tower()
    tower._cfg = SimpleNamespace(llm_mask_token_id=cfg.mask_token_id, llm_depth_token_id=cfg.depth_token_id)

    class VLM(nn.Module, LlavaMetaForCausalLM):
        def __init__(self):
            super().__init__()
            self.llm = llm
            self.vt = tower
            self.mm_projector = projector
            self.region_extractor = extractor
            self.config = SimpleNamespace(enable_region=cfg.enable_region, enable_depth=cfg.enable_depth,
                                          mm_hidden_size=cfg.v_hidden, hidden_size=cfg.hidden)

        def get_vision_tower(self):
            return self.vt

        def get_mm_projector(self):
            return self.mm_projector

        def get_region_extractor(self):
            return self.region_extractor

        def get_llm(self):
            return self.llm

        @property
        def device(self):
            return torch.device("cpu")

    return VLM().eval()


@torch.no_grad()
def run_case(name: str):
    kw, n_regions, t_text, kind, n_new, depth_on = CASES[name]
    cfg = O.OracleConfig(**kw)
    input_ids, images, depths, masks = O.synth_request(cfg, n_regions, t_text, seed=1234, kind=kind)
    if not depth_on:
        depths = None
    # pick the weight seed whose greedy continuation has the widest top-1/top-2 logit margin, so the
    # committed token ids are robust to bf16 rounding on the GPU path (SURVEY.md §7 "hard parts")
    best = (-1.0, None)
    for seed in range(24):
        w = O.make_weights(cfg, seed=seed)
        _, enc = O.generate(cfg, w, input_ids, images, depths, masks, n_new, return_all=True)
        top2 = enc["logits"].topk(2, -1).values
        margin = float((top2[:, 0] - top2[:, 1]).min())
        if margin > best[0]:
            best = (margin, seed)
    print(f"{name}: weight seed {best[1]} (min margin {best[0]:.3f})")
    weight_seed = best[1]
    weights = O.make_weights(cfg, seed=weight_seed)
    vlm = build_reference_vlm(cfg, weights)

    tower_features = vlm.get_vision_tower()(images)
    hres, lres = vlm.get_region_extractor().feature_refinement(tower_features)
    depth_features = vlm.get_vision_tower()(depths) if depths is not None else None
    mask_embeds, depth_embeds = vlm.get_region_extractor()(hres, depth_features, masks)
    image_features = vlm.get_mm_projector()(lres)

    (_, _, attn, _, inputs_embeds, _) = vlm.prepare_inputs_labels_for_multimodal(
        input_ids, None, None, None, None, images, masks, depths)
    assert attn is None
    out = vlm.llm.generate(inputs_embeds=inputs_embeds, do_sample=False, max_new_tokens=n_new, use_cache=True,
                           eos_token_id=None, pad_token_id=0, output_logits=True, return_dict_in_generate=True)
    new_ids = out.sequences[0]
    logits = torch.stack([l[0] for l in out.logits])

    arrays = dict(
        input_ids=input_ids.numpy(), images=images.numpy(), masks=masks[0].numpy(),
        tower_features=tower_features.numpy(), hres=hres.numpy(), lres=lres.numpy(),
        mask_embeds=mask_embeds[0].numpy(), image_features=image_features.numpy(),
        inputs_embeds=inputs_embeds.numpy(), new_ids=new_ids.numpy(), logits=logits.numpy(),
        weight_seed=np.array(weight_seed),
    )
    if depths is not None:
        arrays["depths"] = depths.numpy()
        arrays["depth_features"] = depth_features.numpy()
        arrays["depth_embeds"] = depth_embeds[0].numpy()
    path = os.path.join(HERE, f"{name}.npz")
    np.savez_compressed(path, **{k: (v.astype(np.float32) if v.dtype == np.float64 else v) for k, v in arrays.items()})
    print(f"{name}: S={inputs_embeds.shape[1]} new_ids={new_ids.tolist()} -> {path} ({os.path.getsize(path)/1e6:.2f} MB)")


@torch.no_grad()
def run_maskpool_kats():
    """Op-level known answers from the reference MaskPooling / DownSampleBlock / LayerNorm2d in
    bf16 (the rounding-faithful comparator), at the real 448->128 / 448->32 / 384->108 geometries
    but a narrow channel count so the fixture stays small."""
    be = ref_shim.load_standalone("llava/model/region_extractor/base_extractor.py", "ref_base_extractor")
    bp = ref_shim.load_standalone("llava/model/multimodal_projector/base_projector.py", "ref_base_projector")
    g = torch.Generator().manual_seed(99)
    arrays = {}
    pool = be.MaskPooling()
    for tag, R, side, C, M in (("rgb448", 448, 128, 32, 5), ("depth448", 448, 32, 32, 5), ("rgb384", 384, 108, 24, 3),
                               ("odd336", 336, 24, 16, 2)):
        x = torch.randn(1, side * side, C, generator=g).to(torch.bfloat16)
        masks = (torch.rand(M, R, R, generator=g) > 0.6).float()
        masks[0] = 0  # an all-zero mask exercises the +1e-8 denominator
        masks[1, : R // 3, : R // 2] = 1
        soft = torch.rand(R, R, generator=g)  # non-binary mask (bicubic-resized masks are floats, mm_utils.py:479-482)
        masks[-1] = soft
        masks = masks.half().float()  # stored as fp16 to keep the fixture small; exact for the values used
        out_bf16 = pool(x, [masks], return_list=True)[0]
        out_f32 = pool(x.float(), [masks], return_list=True)[0]
        arrays[f"{tag}_x"] = x.float().numpy()
        arrays[f"{tag}_masks"] = masks.numpy().astype(np.float16)
        arrays[f"{tag}_out_bf16"] = out_bf16.float().numpy()
        arrays[f"{tag}_out_f32"] = out_f32.numpy()
    ds = bp.DownSampleBlock()
    x = torch.randn(2, 27 * 27, 8, generator=g)
    arrays["downsample_x"] = x.numpy()
    arrays["downsample_out"] = ds(x).numpy()
    ln = be.LayerNorm2d(12)
    ln.weight.data = torch.randn(12, generator=g)
    ln.bias.data = torch.randn(12, generator=g)
    x = torch.randn(2, 12, 5, 7, generator=g)
    arrays["ln2d_x"], arrays["ln2d_w"], arrays["ln2d_b"] = x.numpy(), ln.weight.data.numpy(), ln.bias.data.numpy()
    arrays["ln2d_out"] = ln(x).numpy()
    path = os.path.join(HERE, "op_kats.npz")
    np.savez_compressed(path, **arrays)
    print(f"op_kats -> {path} ({os.path.getsize(path)/1e6:.2f} MB)")


PROJECTOR_TYPES = ("linear", "mlp2x_gelu", "mlp3x_gelu", "identity")


@torch.no_grad()
def run_projector_kats():
    """The reference MultimodalProjector (base_projector.py:55-94) for its non-default types, fp32, seeded weights that are
    stored in the fixture (keys = the reference module's own state-dict names)."""
    ref_shim.install()
    from llava.model.multimodal_projector.base_projector import MultimodalProjector, MultimodalProjectorConfig

    g = torch.Generator().manual_seed(321)
    C, H = 48, 64
    arrays = {}
    x = torch.randn(2, 9, C, generator=g).to(torch.bfloat16).float()
    arrays["x"] = x.numpy()
    for t in PROJECTOR_TYPES:
        m = MultimodalProjector(MultimodalProjectorConfig(t), SimpleNamespace(mm_hidden_size=C, hidden_size=H)).float().eval()
        sd = {k: (torch.randn(v.shape, generator=g) * 0.1).to(torch.bfloat16).float() for k, v in m.state_dict().items()}
        m.load_state_dict(sd, strict=True)
        arrays[f"{t}__out"] = m(x).numpy()
        for k, v in sd.items():
            arrays[f"{t}__w__{k}"] = v.numpy()
    path = os.path.join(HERE, "proj_kats.npz")
    np.savez_compressed(path, **arrays)
    print(f"proj_kats -> {path}: {[k for k in arrays if '__w__' in k]}")


CLIP_CASE = dict(image_size=56, v_hidden=128, v_layers=4, v_heads=2, v_inter=256, v_eps=1e-5, v_type="clip", v_act="quick_gelu",
                 select_feature="patch", hidden=256, layers=2, heads=2, kv_heads=1, inter=384, vocab=512, rope_theta=10000.0,
                 mask_token_id=510, depth_token_id=511)
CLIP_WEIGHT_SEED = 11


@torch.no_grad()
def run_clip_kat():
    """The reference's ``VisionTower.forward`` + ``feature_select("patch")`` (vision_encoder.py:26-34,115-132) over stock HF
    ``CLIPVisionModel`` - what ``CLIPVisionTower`` (clip_encoder.py:8-13) wraps - in fp32, on the oracle's seeded CLIP weights."""
    ref_shim.install()
    from transformers import CLIPVisionConfig, CLIPVisionModel

    from llava.model.multimodal_encoder.vision_encoder import VisionTower

    cfg = O.OracleConfig(**CLIP_CASE)
    sd = O.make_weights(cfg, seed=CLIP_WEIGHT_SEED)
    vcfg = CLIPVisionConfig(hidden_size=cfg.v_hidden, intermediate_size=cfg.v_inter, num_hidden_layers=cfg.v_layers,
                            num_attention_heads=cfg.v_heads, image_size=cfg.image_size, patch_size=cfg.patch_size,
                            layer_norm_eps=cfg.v_eps, hidden_act="quick_gelu")
    vcfg._attn_implementation = "eager"
    clip = CLIPVisionModel(vcfg).float().eval()
    missing, unexpected = clip.load_state_dict({k: v.float() for k, v in sd["vision_tower"].items()}, strict=False)
    assert not unexpected, unexpected
    assert all("post_layernorm" in k or "position_ids" in k for k in missing), missing

    class Tower(VisionTower):
        def __init__(self):
            super().__init__("synthetic", SimpleNamespace(mm_vision_select_layer=cfg.select_layer, mm_vision_select_feature="patch"))
            self.vision_tower = clip
            self.is_loaded = True

        @property
        def dtype(self):
            return torch.float32

        @property
        def device(self):
            return torch.device("cpu")

    g = torch.Generator().manual_seed(77)
    images = torch.randn(3, 3, cfg.image_size, cfg.image_size, generator=g).to(torch.bfloat16).float()
    feats = Tower()(images)
    assert feats.shape == (3, cfg.grid ** 2, cfg.v_hidden)
    path = os.path.join(HERE, "clip_tower.npz")
    np.savez_compressed(path, images=images.numpy(), tower_features=feats.numpy(), weight_seed=np.int64(CLIP_WEIGHT_SEED))
    print(f"clip_tower -> {path}: features {tuple(feats.shape)}, rms {float(feats.pow(2).mean().sqrt()):.4f}")


@torch.no_grad()
def run_rope_kat():
    """cos / sin of the reference's rotary classes (modeling_llama.py:81-141), plain and with linear scaling, fp32."""
    # the file is written to REPLACE transformers/models/llama/modeling_llama.py (relative imports): load it inside that package
    import importlib.util

    import transformers.models.llama  # noqa: F401
    spec = importlib.util.spec_from_file_location("transformers.models.llama._srgpt_ref_modeling",
                                                  "/root/reference/llava/train/transformers_replace/models/llama/modeling_llama.py")
    m = importlib.util.module_from_spec(spec)
    m.__package__ = "transformers.models.llama"
    sys.modules[spec.name] = m
    spec.loader.exec_module(m)
    LlamaLinearScalingRotaryEmbedding, LlamaRotaryEmbedding = m.LlamaLinearScalingRotaryEmbedding, m.LlamaRotaryEmbedding

    pos = torch.tensor([[0, 1, 2, 3, 17, 258, 259, 1000, 4095]])
    x = torch.zeros(1, 1, pos.shape[1], 128)
    arrays = {"positions": pos[0].numpy()}
    for name, emb in (("plain", LlamaRotaryEmbedding(128, max_position_embeddings=8192, base=500000.0)),
                      ("linear4", LlamaLinearScalingRotaryEmbedding(128, max_position_embeddings=8192, base=500000.0, scaling_factor=4.0))):
        cos, sin = emb(x, pos)
        arrays[name + "_cos"], arrays[name + "_sin"] = cos[0].numpy(), sin[0].numpy()
    path = os.path.join(HERE, "rope_kats.npz")
    np.savez_compressed(path, **arrays)
    print(f"rope_kats -> {path}")


BEAM_CASES = [  # (num_beams, eos_token_id, max_new_tokens, length_penalty, early_stopping)
    (1, None, 10, 1.0, False), (3, None, 10, 1.0, False), (3, [460], 10, 1.0, False), (4, [38, 97], 12, 1.0, False),
    (2, [886], 10, 1.0, False), (4, [134, 562], 16, 1.0, False), (3, [764], 16, 2.0, False), (3, [764], 16, 0.5, True), (5, [303], 14, 1.0, True),
]
BEAM_WEIGHT_SEED = 7


@torch.no_grad()
def run_beam_kats():
    """HF ``generate(inputs_embeds=..., num_beams=k)`` (the call llava_llama.py:212 makes when the eval scripts pass --num_beams) on a stock
    LlamaForCausalLM holding the oracle's seeded weights.  Run with the transformers of this image (5.5; the reference pins 4.37.2,
    whose beam search is the same algorithm for these settings) - the generated ids pin the oracle's restatement."""
    from transformers import LlamaConfig, LlamaForCausalLM

    cfg = O.OracleConfig(**CASES["tiny_masks_gqa"][0])
    sd = O.make_weights(cfg, seed=BEAM_WEIGHT_SEED)
    lcfg = LlamaConfig(hidden_size=cfg.hidden, intermediate_size=cfg.inter, num_hidden_layers=cfg.layers, num_attention_heads=cfg.heads,
                       num_key_value_heads=cfg.kv_heads, vocab_size=cfg.vocab, rms_norm_eps=cfg.rms_eps, rope_theta=cfg.rope_theta,
                       max_position_embeddings=4096, tie_word_embeddings=False, head_dim=cfg.head_dim, attention_bias=False, mlp_bias=False,
                       bos_token_id=1, eos_token_id=None, pad_token_id=None)
    lcfg._attn_implementation = "eager"
    llm = LlamaForCausalLM(lcfg).float().eval()
    llm.load_state_dict({k: v.float() for k, v in sd["llm"].items()}, strict=True)
    g = torch.Generator().manual_seed(0)
    emb = (torch.randn(20, cfg.hidden, generator=g) * 0.3).to(torch.bfloat16).float()
    arrays = {"inputs_embeds": emb.numpy(), "weight_seed": np.int64(BEAM_WEIGHT_SEED)}
    for i, (nb, eos, n_new, lp, es) in enumerate(BEAM_CASES):
        out = llm.generate(inputs_embeds=emb[None], num_beams=nb, do_sample=False, max_new_tokens=n_new, eos_token_id=eos, pad_token_id=0,
                           early_stopping=es, length_penalty=lp, num_return_sequences=1)[0]
        arrays[f"case{i}"] = out.numpy()
        print(i, nb, eos, n_new, lp, es, out.tolist())
    path = os.path.join(HERE, "beam_kats.npz")
    np.savez_compressed(path, **arrays)
    print(f"beam_kats -> {path}")


if __name__ == "__main__":
    torch.manual_seed(0)
    torch.set_num_threads(8)
    if sys.argv[1:] == ["beam"]:
        run_beam_kats()
        sys.exit(0)
    if sys.argv[1:] == ["rope"]:
        run_rope_kat()
        sys.exit(0)
    if sys.argv[1:] == ["clip"]:
        run_clip_kat()
        sys.exit(0)
    if sys.argv[1:] == ["proj"]:  # only the projector-type fixture (the others regenerate bit-identically but take minutes)
        run_projector_kats()
        sys.exit(0)
    for n in CASES:
        run_case(n)
    run_maskpool_kats()
    run_projector_kats()
    run_clip_kat()
    run_rope_kat()
    run_beam_kats()
