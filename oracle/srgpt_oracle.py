"""CPU oracle for SpatialRGPT's multimodal generate() hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in ``spatialrgpt_b200/`` may import this module; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs use it, and only as the checker / the timed CPU baseline.

It is a plain-torch *restatement* (no nn.Module classes, no HF model classes) of the arithmetic
the reference executes for one (image, regions, prompt) request.  Each function cites the
reference file:line it follows (paths relative to the reference checkout):

* SigLIP vision tower — arithmetic lives in the third-party dependency ``transformers==4.37.2``
  (``pyproject.toml:17``; call sites ``llava/model/multimodal_encoder/siglip_encoder.py:11-16``,
  ``vision_encoder.py:115-132``).  Restated from the published SigLIP algorithm (pre-LN ViT,
  tanh-GELU, learned position embedding, no CLS token).
* Region extractor, MaskPooling, LayerNorm2d — ``llava/model/region_extractor/base_extractor.py``.
* mm_projector (mlp_downsample) — ``llava/model/multimodal_projector/base_projector.py``.
* Embedding splice — ``llava/model/llava_arch.py:333-650``.
* Llama decoder — ``llava/train/transformers_replace/models/llama/modeling_llama.py``.
* Greedy loop — HF ``GenerationMixin`` (third party, call site ``llava_llama.py:212``).

Pinning: the reference ships no tests / golden vectors.  The restatement is
pinned instead against outputs of the reference's *own modules* imported in the authoring
container (``tests/golden/make_golden.py`` — the shimmed reference ``RegionExtractor``,
``MultimodalProjector``, ``prepare_inputs_labels_for_multimodal`` driving stock HF
``SiglipVisionModel`` / ``LlamaForCausalLM``); the resulting fixtures are committed under
``tests/golden/`` and ``tests/test_oracle_golden.py`` checks this file against them.

All functions take ``dtype``: ``torch.float32`` is the ground truth; ``torch.bfloat16`` reproduces
the reference's rounding points (every torch op rounds its output to bf16).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F

IMAGE_TOKEN_INDEX = -200  # llava/constants.py:26
IGNORE_INDEX = -100  # llava/constants.py:25


@dataclass
class OracleConfig:
    # vision tower (SigLIP)
    image_size: int = 448
    patch_size: int = 14
    v_hidden: int = 1152
    v_layers: int = 27
    v_heads: int = 16
    v_inter: int = 4304
    v_eps: float = 1e-6
    select_layer: int = -2  # scripts/srgpt/*/3_sft.sh:29
    v_type: str = "siglip"  # "clip": HF CLIPVisionModel behind CLIPVisionTower (clip_encoder.py:8-13)
    v_act: str = "gelu_pytorch_tanh"  # CLIP: "quick_gelu"
    select_feature: str = "cls_patch"  # CLIP: "patch" (vision_encoder.py:28-31)
    # llm (Llama)
    hidden: int = 4096
    layers: int = 32
    heads: int = 32
    kv_heads: int = 8
    head_dim: int = 128
    inter: int = 14336
    vocab: int = 128259
    rope_theta: float = 500000.0
    rope_scaling_factor: float = 1.0  # "linear" scaling (LlamaLinearScalingRotaryEmbedding, modeling_llama.py:133-141)
    rms_eps: float = 1e-5
    # multimodal
    enable_region: bool = True
    enable_depth: bool = True
    mask_token_id: int = 128257
    depth_token_id: int = 128258
    ada_pool: int = 27  # base_extractor.py:123 (hard-coded)

    @property
    def grid(self) -> int:
        return self.image_size // self.patch_size

    @property
    def v_head_dim(self) -> int:
        return self.v_hidden // self.v_heads

    @property
    def n_tower_layers(self) -> int:
        """Layers whose output ``hidden_states[select_layer]`` depends on."""
        # hidden_states = (embeddings, layer_1_out, ..., layer_L_out); index -2 -> layer L-1.
        return self.v_layers + 1 + self.select_layer if self.select_layer < 0 else self.select_layer


# ----------------------------------------------------------------------------------------------
# synthetic weights (names = the reference checkpoint's state-dict keys)
# ----------------------------------------------------------------------------------------------

def make_weights(cfg: OracleConfig, seed: int = 0, std: float = 0.02, nontrivial_norms: bool = True,
                 dtype: torch.dtype = torch.bfloat16, embed_std: float = 0.3) -> Dict[str, Dict[str, torch.Tensor]]:
    """Seeded random weights keyed like the reference's four-directory checkpoint
    (``llava_arch.py:181-250``).  Values are rounded to ``dtype`` so that the GPU path and the
    fp32 oracle see bit-identical parameters."""
    g = torch.Generator().manual_seed(seed)

    def rn(*shape, s=std):
        return (torch.randn(*shape, generator=g) * s).to(dtype)

    def norm_w(n):
        if nontrivial_norms:
            return (1.0 + 0.1 * torch.randn(n, generator=g)).to(dtype)
        return torch.ones(n, dtype=dtype)

    def norm_b(n):
        if nontrivial_norms:
            return (0.05 * torch.randn(n, generator=g)).to(dtype)
        return torch.zeros(n, dtype=dtype)

    D, I = cfg.v_hidden, cfg.v_inter
    vt: Dict[str, torch.Tensor] = {}
    vt["vision_model.embeddings.patch_embedding.weight"] = rn(D, 3, cfg.patch_size, cfg.patch_size)
    if cfg.v_type == "clip":  # HF CLIPVisionEmbeddings / CLIPVisionTransformer parameter names (bias-free conv, "pre_layrnorm" sic)
        vt["vision_model.embeddings.class_embedding"] = rn(D)
        vt["vision_model.embeddings.position_embedding.weight"] = rn(cfg.grid * cfg.grid + 1, D)
        vt["vision_model.pre_layrnorm.weight"] = norm_w(D)
        vt["vision_model.pre_layrnorm.bias"] = norm_b(D)
    else:
        vt["vision_model.embeddings.patch_embedding.bias"] = rn(D)
        vt["vision_model.embeddings.position_embedding.weight"] = rn(cfg.grid * cfg.grid, D)
    for i in range(cfg.v_layers):
        p = f"vision_model.encoder.layers.{i}."
        vt[p + "layer_norm1.weight"] = norm_w(D)
        vt[p + "layer_norm1.bias"] = norm_b(D)
        for n in ("q_proj", "k_proj", "v_proj", "out_proj"):
            vt[p + f"self_attn.{n}.weight"] = rn(D, D, s=std * 2)
            vt[p + f"self_attn.{n}.bias"] = rn(D)
        vt[p + "layer_norm2.weight"] = norm_w(D)
        vt[p + "layer_norm2.bias"] = norm_b(D)
        vt[p + "mlp.fc1.weight"] = rn(I, D, s=std * 2)
        vt[p + "mlp.fc1.bias"] = rn(I)
        vt[p + "mlp.fc2.weight"] = rn(D, I, s=std * 2)
        vt[p + "mlp.fc2.bias"] = rn(D)

    H = cfg.hidden
    re: Dict[str, torch.Tensor] = {}
    re["feature_refinement_module.0.weight"] = rn(D, D, 2, 2, s=std * 2)
    re["feature_refinement_module.0.bias"] = rn(D)
    re["feature_refinement_module.1.weight"] = norm_w(D)
    re["feature_refinement_module.1.bias"] = norm_b(D)
    re["feature_refinement_module.3.weight"] = rn(D, D, 2, 2, s=std * 2)
    re["feature_refinement_module.3.bias"] = rn(D)
    re["rgb_projector.weight"] = rn(H, D, s=std * 2)
    re["rgb_projector.bias"] = rn(H)
    re["depth_projector.weight"] = rn(H, D, s=std * 2)
    re["depth_projector.bias"] = rn(H)

    mp: Dict[str, torch.Tensor] = {}
    mp["layers.1.weight"] = norm_w(4 * D)
    mp["layers.1.bias"] = norm_b(4 * D)
    mp["layers.2.weight"] = rn(H, 4 * D)
    mp["layers.2.bias"] = rn(H)
    mp["layers.4.weight"] = rn(H, H)
    mp["layers.4.bias"] = rn(H)

    llm: Dict[str, torch.Tensor] = {}
    # token embeddings dominate the residual stream so that greedy decoding does not collapse onto
    # one repeated token (it does with std 0.02), which would make id-parity a weak test
    llm["model.embed_tokens.weight"] = rn(cfg.vocab, H, s=embed_std)
    qd, kd = cfg.heads * cfg.head_dim, cfg.kv_heads * cfg.head_dim
    for i in range(cfg.layers):
        p = f"model.layers.{i}."
        llm[p + "input_layernorm.weight"] = norm_w(H)
        llm[p + "self_attn.q_proj.weight"] = rn(qd, H)
        llm[p + "self_attn.k_proj.weight"] = rn(kd, H)
        llm[p + "self_attn.v_proj.weight"] = rn(kd, H)
        llm[p + "self_attn.o_proj.weight"] = rn(H, qd)
        llm[p + "post_attention_layernorm.weight"] = norm_w(H)
        llm[p + "mlp.gate_proj.weight"] = rn(cfg.inter, H)
        llm[p + "mlp.up_proj.weight"] = rn(cfg.inter, H)
        llm[p + "mlp.down_proj.weight"] = rn(H, cfg.inter)
    llm["model.norm.weight"] = norm_w(H)
    # a "peaky" lm_head keeps greedy argmax margins well above bf16 noise (SURVEY.md §7 hard parts)
    llm["lm_head.weight"] = rn(cfg.vocab, H, s=std * 4)
    return {"vision_tower": vt, "region_extractor": re, "mm_projector": mp, "llm": llm}


# ----------------------------------------------------------------------------------------------
# vision tower  (third-party transformers SiglipVisionModel; vision_encoder.py:115-132)
# ----------------------------------------------------------------------------------------------

def gelu_tanh(x: torch.Tensor) -> torch.Tensor:
    return F.gelu(x, approximate="tanh")


def quick_gelu(x: torch.Tensor) -> torch.Tensor:
    """HF ``QuickGELUActivation`` (transformers 4.37.2 activations.py): three ops in the tensor's dtype."""
    return x * torch.sigmoid(1.702 * x)


def clip_tower_forward(cfg: OracleConfig, w: Dict[str, torch.Tensor], images: torch.Tensor,
                       dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """``CLIPVisionTower`` (clip_encoder.py:8-13) through ``VisionTower.forward`` + ``feature_select`` (vision_encoder.py:26-34,
    115-132).  The arithmetic is third-party: HF ``CLIPVisionModel`` (transformers 4.37.2, modeling_clip.py) - bias-free patch
    convolution, class token, position embedding, ``pre_layrnorm``, pre-LN encoder layers whose attention scales q BEFORE q k^T
    (``q_proj(x) * scale``) and takes the softmax in the tensor's dtype, ``quick_gelu`` MLP; ``hidden_states[select_layer]`` with
    the class-token row dropped for "patch"."""
    W = lambda k: w[k].to(dtype)  # noqa: E731
    x = images.to(dtype)
    N = x.shape[0]
    D, ps = cfg.v_hidden, cfg.patch_size
    x = F.conv2d(x, W("vision_model.embeddings.patch_embedding.weight"), None, stride=ps)
    x = x.flatten(2).transpose(1, 2)
    cls = W("vision_model.embeddings.class_embedding").expand(N, 1, -1)
    x = torch.cat([cls, x], dim=1) + W("vision_model.embeddings.position_embedding.weight")[None]
    x = F.layer_norm(x, (D,), W("vision_model.pre_layrnorm.weight"), W("vision_model.pre_layrnorm.bias"), cfg.v_eps)
    nh, hd = cfg.v_heads, cfg.v_head_dim
    scale = hd ** -0.5
    act = {"quick_gelu": quick_gelu, "gelu": F.gelu, "gelu_pytorch_tanh": gelu_tanh}[cfg.v_act]
    for i in range(cfg.n_tower_layers):
        p = f"vision_model.encoder.layers.{i}."
        r = x
        h = F.layer_norm(x, (D,), W(p + "layer_norm1.weight"), W(p + "layer_norm1.bias"), cfg.v_eps)
        q = F.linear(h, W(p + "self_attn.q_proj.weight"), W(p + "self_attn.q_proj.bias")) * scale
        k = F.linear(h, W(p + "self_attn.k_proj.weight"), W(p + "self_attn.k_proj.bias"))
        v = F.linear(h, W(p + "self_attn.v_proj.weight"), W(p + "self_attn.v_proj.bias"))
        q = q.view(N, -1, nh, hd).transpose(1, 2)
        k = k.view(N, -1, nh, hd).transpose(1, 2)
        v = v.view(N, -1, nh, hd).transpose(1, 2)
        att = F.softmax(torch.matmul(q, k.transpose(-1, -2)), dim=-1)
        o = torch.matmul(att, v).transpose(1, 2).reshape(N, -1, D)
        o = F.linear(o, W(p + "self_attn.out_proj.weight"), W(p + "self_attn.out_proj.bias"))
        x = r + o
        r = x
        h = F.layer_norm(x, (D,), W(p + "layer_norm2.weight"), W(p + "layer_norm2.bias"), cfg.v_eps)
        h = act(F.linear(h, W(p + "mlp.fc1.weight"), W(p + "mlp.fc1.bias")))
        h = F.linear(h, W(p + "mlp.fc2.weight"), W(p + "mlp.fc2.bias"))
        x = r + h
    if cfg.select_feature == "patch":
        x = x[:, 1:]
    elif cfg.select_feature != "cls_patch":
        raise ValueError(f"Unexpected select feature: {cfg.select_feature}")  # vision_encoder.py:33
    return x


def vision_tower_forward(cfg: OracleConfig, w: Dict[str, torch.Tensor], images: torch.Tensor,
                         dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """``VisionTower.forward`` + ``feature_select`` (vision_encoder.py:26-34,115-132):
    SigLIP forward, returns ``hidden_states[select_layer]`` with all T tokens ("cls_patch")."""
    if cfg.v_type == "clip":
        return clip_tower_forward(cfg, w, images, dtype)
    W = lambda k: w[k].to(dtype)  # noqa: E731
    x = images.to(dtype)
    N = x.shape[0]
    D, P, ps = cfg.v_hidden, cfg.grid, cfg.patch_size
    # patch embedding: Conv2d(k=ps, s=ps) == unfold + linear
    x = F.conv2d(x, W("vision_model.embeddings.patch_embedding.weight"),
                 W("vision_model.embeddings.patch_embedding.bias"), stride=ps)
    x = x.flatten(2).transpose(1, 2)  # [N, T, D], row-major (H W)
    x = x + W("vision_model.embeddings.position_embedding.weight")[None]
    nh, hd = cfg.v_heads, cfg.v_head_dim
    scale = hd ** -0.5
    for i in range(cfg.n_tower_layers):
        p = f"vision_model.encoder.layers.{i}."
        r = x
        h = F.layer_norm(x, (D,), W(p + "layer_norm1.weight"), W(p + "layer_norm1.bias"), cfg.v_eps)
        q = F.linear(h, W(p + "self_attn.q_proj.weight"), W(p + "self_attn.q_proj.bias"))
        k = F.linear(h, W(p + "self_attn.k_proj.weight"), W(p + "self_attn.k_proj.bias"))
        v = F.linear(h, W(p + "self_attn.v_proj.weight"), W(p + "self_attn.v_proj.bias"))
        q = q.view(N, -1, nh, hd).transpose(1, 2)
        k = k.view(N, -1, nh, hd).transpose(1, 2)
        v = v.view(N, -1, nh, hd).transpose(1, 2)
        att = torch.matmul(q, k.transpose(-1, -2)) * scale
        att = F.softmax(att, dim=-1, dtype=torch.float32).to(dtype)
        o = torch.matmul(att, v).transpose(1, 2).reshape(N, -1, D)
        o = F.linear(o, W(p + "self_attn.out_proj.weight"), W(p + "self_attn.out_proj.bias"))
        x = r + o
        r = x
        h = F.layer_norm(x, (D,), W(p + "layer_norm2.weight"), W(p + "layer_norm2.bias"), cfg.v_eps)
        h = F.linear(h, W(p + "mlp.fc1.weight"), W(p + "mlp.fc1.bias"))
        h = gelu_tanh(h)
        h = F.linear(h, W(p + "mlp.fc2.weight"), W(p + "mlp.fc2.bias"))
        x = r + h
    return x


# ----------------------------------------------------------------------------------------------
# region extractor (base_extractor.py)
# ----------------------------------------------------------------------------------------------

def layernorm2d(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    """``LayerNorm2d.forward`` (base_extractor.py:19-24): per-pixel norm over channels, NCHW."""
    u = x.mean(1, keepdim=True)
    s = (x - u).pow(2).mean(1, keepdim=True)
    x = (x - u) / torch.sqrt(s + eps)
    return weight[:, None, None] * x + bias[:, None, None]


def feature_refinement(cfg: OracleConfig, w: Dict[str, torch.Tensor], tower_features: torch.Tensor,
                       dtype: torch.dtype = torch.float32):
    """``RegionExtractor.feature_refinement`` (base_extractor.py:137-147) with the ``deconv2x``
    module (87-101): ConvT(k2,s2) -> LayerNorm2d -> GELU(erf) -> ConvT(k2,s2) -> GELU(erf);
    returns (hres [N,(4P)^2,C], lres [N,27*27,C]), both flattened row-major (H W)."""
    W = lambda k: w[k].to(dtype)  # noqa: E731
    x = tower_features.to(dtype)
    N, HW, C = x.shape
    P = int(HW ** 0.5)
    x = x.view(N, P, P, C).permute(0, 3, 1, 2)
    x = F.conv_transpose2d(x, W("feature_refinement_module.0.weight"), W("feature_refinement_module.0.bias"), stride=2)
    x = layernorm2d(x, W("feature_refinement_module.1.weight"), W("feature_refinement_module.1.bias"))
    x = F.gelu(x)
    x = F.conv_transpose2d(x, W("feature_refinement_module.3.weight"), W("feature_refinement_module.3.bias"), stride=2)
    x = F.gelu(x)
    hres = x.flatten(2).transpose(1, 2)
    lres = F.adaptive_avg_pool2d(x, cfg.ada_pool).flatten(2).transpose(1, 2)
    return hres, lres


def mask_pooling(x: torch.Tensor, mask_list: Optional[Sequence[Optional[torch.Tensor]]]) -> List[Optional[torch.Tensor]]:
    """``MaskPooling.forward(return_list=True)`` (base_extractor.py:32-84).

    x: [B, L, C] in the compute dtype; mask_list[i]: [M, IH, IW] or None.
    scale = sqrt(L/(IH*IW)); bilinear (align_corners=False, no antialias) resize in fp32;
    cast to x.dtype BEFORE the sum and the divide; pooled = einsum('lc,ml->mc')."""
    B = x.shape[0]
    if mask_list is None:
        mask_list = [None] * B
    out: List[Optional[torch.Tensor]] = []
    for i in range(B):
        mask = mask_list[i]
        if mask is None:
            out.append(None)
            continue
        L = x.shape[1]
        scale = (L / (mask.shape[-1] * mask.shape[-2])) ** 0.5
        m = F.interpolate(mask.float()[None], scale_factor=scale, mode="bilinear")[0]
        m = m.to(x.dtype)
        denorm = (m.sum(dim=(-1, -2)) + 1e-8).unsqueeze(-1)
        m = m.flatten(1)
        out.append(torch.einsum("lc,ml->mc", x[i], m / denorm))
    return out


def region_extractor_forward(cfg: OracleConfig, w: Dict[str, torch.Tensor], hres: torch.Tensor,
                             depth_features: Optional[torch.Tensor], masks, dtype: torch.dtype = torch.float32):
    """``RegionExtractor.forward`` (base_extractor.py:149-173): rgb_projector(mask_pool(hres)),
    depth_projector(mask_pool(depth tower features))."""
    W = lambda k: w[k].to(dtype)  # noqa: E731

    def branch(feat, name):
        pooled = mask_pooling(feat.to(dtype), masks)
        return [None if p is None else F.linear(p, W(name + ".weight"), W(name + ".bias")) for p in pooled]

    mask_embeds = branch(hres, "rgb_projector")
    depth_embeds = branch(depth_features, "depth_projector") if depth_features is not None else None
    return mask_embeds, depth_embeds


# ----------------------------------------------------------------------------------------------
# mm_projector (base_projector.py:32-52,73-80)
# ----------------------------------------------------------------------------------------------

def downsample_block(x: torch.Tensor) -> torch.Tensor:
    """``DownSampleBlock`` (base_projector.py:32-52): [N, h*w, C] -> [N, ceil(h/2)*ceil(w/2), 4C]
    with zero padding of odd sides and the reference's view/permute (spatially transposed) order."""
    n, hw, c = x.shape
    h = w = int(hw ** 0.5)
    x = x.reshape(n, h, w, c)
    # flat_square names dim1 "w" and dim2 "h"; keep its exact order of operations
    n, d1, d2, c = x.shape
    if d1 % 2 == 1:
        x = torch.cat([x, torch.zeros(n, 1, d2, c, dtype=x.dtype)], dim=1)
        d1 += 1
    if d2 % 2 == 1:
        x = torch.cat([x, torch.zeros(n, d1, 1, c, dtype=x.dtype)], dim=2)
        d2 += 1
    x = x.contiguous().view(n, d1, d2 // 2, c * 2)
    x = x.permute(0, 2, 1, 3).contiguous()
    x = x.view(n, d2 // 2, d1 // 2, c * 4)
    return x.reshape(n, -1, c * 4)


def mm_projector_forward(cfg: OracleConfig, w: Dict[str, torch.Tensor], lres: torch.Tensor,
                         dtype: torch.dtype = torch.float32, ptype: str = "mlp_downsample") -> torch.Tensor:
    """``MultimodalProjector.forward`` (base_projector.py:69-94).  ``mlp_downsample`` (:73-80): DownSample -> LN(4C) -> Linear ->
    GELU(erf) -> Linear; ``identity`` (:69-70); ``linear`` (:71-72); ``mlpNx_gelu`` (:81-88): Linear, then N-1 x (GELU, Linear)."""
    W = lambda k: w[k].to(dtype)  # noqa: E731
    if ptype == "identity":
        return lres.to(dtype)
    if ptype == "linear":
        return F.linear(lres.to(dtype), W("layers.weight"), W("layers.bias"))
    if ptype != "mlp_downsample":
        import re
        depth = int(re.match(r"^mlp(\d+)x_gelu$", ptype).group(1))
        x = F.linear(lres.to(dtype), W("layers.0.weight"), W("layers.0.bias"))
        for i in range(1, depth):
            x = F.linear(F.gelu(x), W(f"layers.{2 * i}.weight"), W(f"layers.{2 * i}.bias"))
        return x
    x = downsample_block(lres.to(dtype))
    x = F.layer_norm(x, (x.shape[-1],), W("layers.1.weight"), W("layers.1.bias"), 1e-5)
    x = F.linear(x, W("layers.2.weight"), W("layers.2.bias"))
    x = F.gelu(x)
    x = F.linear(x, W("layers.4.weight"), W("layers.4.bias"))
    return x


# ----------------------------------------------------------------------------------------------
# encode + splice (llava_arch.py:333-650)
# ----------------------------------------------------------------------------------------------

def encode_multimodal(cfg: OracleConfig, weights, images: torch.Tensor, depths: Optional[torch.Tensor], masks,
                      dtype: torch.dtype = torch.float32):
    """llava_arch.py:398-411: tower(images) -> refinement -> (tower(depths)) -> region extractor ->
    mm_projector.  Returns dict of every intermediate (used as stage goldens)."""
    out = {}
    tf = vision_tower_forward(cfg, weights["vision_tower"], images, dtype)
    out["tower_features"] = tf
    if cfg.enable_region:
        hres, lres = feature_refinement(cfg, weights["region_extractor"], tf, dtype)
        out["hres"], out["lres"] = hres, lres
        if cfg.enable_depth and depths is not None:
            df = vision_tower_forward(cfg, weights["vision_tower"], depths, dtype)
            out["depth_features"] = df
            me, de = region_extractor_forward(cfg, weights["region_extractor"], hres, df, masks, dtype)
        else:
            me, de = region_extractor_forward(cfg, weights["region_extractor"], hres, None, masks, dtype)
        out["mask_embeds"], out["depth_embeds"] = me, de
    else:
        lres = tf
        out["mask_embeds"] = out["depth_embeds"] = None
    out["image_features"] = mm_projector_forward(cfg, weights["mm_projector"], lres, dtype)
    return out


def splice_embeddings(cfg: OracleConfig, embed_tokens: torch.Tensor, input_ids: torch.Tensor,
                      image_features: torch.Tensor, mask_embeds, depth_embeds,
                      attention_mask: Optional[torch.Tensor] = None, depths_given: bool = True) -> List[torch.Tensor]:
    """llava_arch.py:434-539 for one batch: embed text (image slots -> id 0), overwrite <mask> /
    <depth> rows in token order with mask_embed[:num_mask] / depth_embed[:num_depth], replace each
    IMAGE_TOKEN_INDEX by that image's feature rows.  Returns the per-sample un-padded embeddings."""
    B = input_ids.shape[0]
    if attention_mask is None:
        attention_mask = torch.ones_like(input_ids, dtype=torch.bool)
    attention_mask = attention_mask.bool()
    ids_copy = input_ids.clone()
    ids_copy[ids_copy == IMAGE_TOKEN_INDEX] = 0
    emb = embed_tokens[ids_copy]
    outs = []
    cur_image_idx = 0
    for b in range(B):
        ids = input_ids[b][attention_mask[b]]
        e = emb[b][attention_mask[b]].clone()
        n_img = int((ids == IMAGE_TOKEN_INDEX).sum())
        if n_img == 0:
            outs.append(e)
            continue
        if cfg.enable_region and mask_embeds is not None:
            pos = ids == cfg.mask_token_id
            me = mask_embeds[cur_image_idx]
            if me is not None:
                e[pos] = me[: int(pos.sum())].to(e.dtype)
        if cfg.enable_depth and depths_given and depth_embeds is not None:
            pos = ids == cfg.depth_token_id
            de = depth_embeds[cur_image_idx]
            if de is not None:
                e[pos] = de[: int(pos.sum())].to(e.dtype)
        pieces = []
        start = 0
        img_pos = torch.where(ids == IMAGE_TOKEN_INDEX)[0].tolist()
        for p in img_pos:
            pieces.append(e[start:p])
            pieces.append(image_features[cur_image_idx].to(e.dtype))
            cur_image_idx += 1
            start = p + 1
        pieces.append(e[start:])
        outs.append(torch.cat(pieces, dim=0))
    return outs


# ----------------------------------------------------------------------------------------------
# Llama decoder (modeling_llama.py)
# ----------------------------------------------------------------------------------------------

def rms_norm(x: torch.Tensor, weight: torch.Tensor, eps: float) -> torch.Tensor:
    """``LlamaRMSNorm.forward`` (modeling_llama.py:70-75): fp32 normalise, cast, THEN * weight."""
    dt = x.dtype
    h = x.to(torch.float32)
    var = h.pow(2).mean(-1, keepdim=True)
    h = h * torch.rsqrt(var + eps)
    return weight * h.to(dt)


def rope_cos_sin(cfg: OracleConfig, positions: torch.Tensor, dtype: torch.dtype):
    """``LlamaRotaryEmbedding.forward`` (modeling_llama.py:117-130): fp32 freqs, cast to dtype."""
    hd = cfg.head_dim
    inv_freq = 1.0 / (cfg.rope_theta ** (torch.arange(0, hd, 2, dtype=torch.int64).float() / hd))
    pos = positions.float()
    if cfg.rope_scaling_factor != 1.0:
        pos = pos / cfg.rope_scaling_factor  # modeling_llama.py:138
    freqs = pos[:, None] * inv_freq[None, :]
    emb = torch.cat((freqs, freqs), dim=-1)
    return emb.cos().to(dtype), emb.sin().to(dtype)


def rotate_half(x: torch.Tensor) -> torch.Tensor:
    x1, x2 = x[..., : x.shape[-1] // 2], x[..., x.shape[-1] // 2:]
    return torch.cat((-x2, x1), dim=-1)


def llama_forward(cfg: OracleConfig, w: Dict[str, torch.Tensor], inputs_embeds: torch.Tensor,
                  kv_cache: Optional[list], dtype: torch.dtype = torch.float32, return_hidden: bool = False):
    """One ``LlamaForCausalLM.forward`` (modeling_llama.py:824-936,1011-1045) for a single
    un-padded sequence: inputs_embeds [S, H]; kv_cache = list of (k, v) per layer with shapes
    [kvh, S_past, hd] or None.  Causal softmax attention with GQA (== FA2 causal, 564-566).
    Returns (fp32 logits [S, V], new kv_cache)."""
    W = lambda k: w[k].to(dtype)  # noqa: E731
    x = inputs_embeds.to(dtype)
    S = x.shape[0]
    past = 0 if kv_cache is None else kv_cache[0][0].shape[1]
    pos = torch.arange(past, past + S)
    cos, sin = rope_cos_sin(cfg, pos, dtype)
    nh, nkv, hd = cfg.heads, cfg.kv_heads, cfg.head_dim
    new_cache = []
    hiddens = []
    for i in range(cfg.layers):
        p = f"model.layers.{i}."
        r = x
        h = rms_norm(x, W(p + "input_layernorm.weight"), cfg.rms_eps)
        q = F.linear(h, W(p + "self_attn.q_proj.weight")).view(S, nh, hd).transpose(0, 1)
        k = F.linear(h, W(p + "self_attn.k_proj.weight")).view(S, nkv, hd).transpose(0, 1)
        v = F.linear(h, W(p + "self_attn.v_proj.weight")).view(S, nkv, hd).transpose(0, 1)
        q = (q * cos[None]) + (rotate_half(q) * sin[None])
        k = (k * cos[None]) + (rotate_half(k) * sin[None])
        if kv_cache is not None:
            k = torch.cat([kv_cache[i][0], k], dim=1)
            v = torch.cat([kv_cache[i][1], v], dim=1)
        new_cache.append((k, v))
        g = nh // nkv
        kk = k.repeat_interleave(g, dim=0)
        vv = v.repeat_interleave(g, dim=0)
        att = torch.matmul(q, kk.transpose(-1, -2)).float() * (hd ** -0.5)
        Sk = kk.shape[1]
        causal = torch.arange(Sk)[None, :] <= (past + torch.arange(S))[:, None]
        att = att.masked_fill(~causal[None], float("-inf"))
        att = F.softmax(att, dim=-1).to(dtype)
        o = torch.matmul(att, vv).transpose(0, 1).reshape(S, nh * hd)
        o = F.linear(o, W(p + "self_attn.o_proj.weight"))
        x = r + o
        r = x
        h = rms_norm(x, W(p + "post_attention_layernorm.weight"), cfg.rms_eps)
        h = F.linear(F.silu(F.linear(h, W(p + "mlp.gate_proj.weight"))) * F.linear(h, W(p + "mlp.up_proj.weight")),
                     W(p + "mlp.down_proj.weight"))
        x = r + h
        if return_hidden:
            hiddens.append(x)
    x = rms_norm(x, W("model.norm.weight"), cfg.rms_eps)
    logits = F.linear(x, W("lm_head.weight")).float()
    if return_hidden:
        return logits, new_cache, hiddens
    return logits, new_cache


def greedy_generate(cfg: OracleConfig, w_llm: Dict[str, torch.Tensor], inputs_embeds: torch.Tensor,
                    max_new_tokens: int, eos_token_id=None, dtype: torch.dtype = torch.float32,
                    return_logits: bool = False):
    """HF ``GenerationMixin`` greedy search started from ``inputs_embeds`` (llava_llama.py:212):
    returns only the NEW ids.  ``eos_token_id`` may be an int, a list, or None (never stop)."""
    eos = set()
    if eos_token_id is not None:
        eos = set(eos_token_id) if isinstance(eos_token_id, (list, tuple, set)) else {int(eos_token_id)}
    logits, cache = llama_forward(cfg, w_llm, inputs_embeds, None, dtype)
    out, all_logits = [], []
    emb = w_llm["model.embed_tokens.weight"]
    for _ in range(max_new_tokens):
        last = logits[-1]
        all_logits.append(last)
        nxt = int(torch.argmax(last))
        out.append(nxt)
        if nxt in eos:
            break
        logits, cache = llama_forward(cfg, w_llm, emb[nxt][None].to(dtype), cache, dtype)
    ids = torch.tensor(out, dtype=torch.long)
    if return_logits:
        return ids, torch.stack(all_logits)
    return ids


def beam_search_generate(cfg: OracleConfig, w_llm: Dict[str, torch.Tensor], inputs_embeds: torch.Tensor, num_beams: int,
                         max_new_tokens: int, eos_token_id=None, length_penalty: float = 1.0, early_stopping: bool = False,
                         dtype: torch.dtype = torch.float32, return_trace: bool = False):
    """HF ``GenerationMixin.beam_search`` + ``BeamSearchScorer`` (third party, transformers 4.37.2 generation/utils.py and
    generation/beam_search.py; call sites llava_llama.py:212 with ``num_beams`` from eval_spatial.py:234 / eval_region_cls.py:320) for one
    prompt given as ``inputs_embeds`` (so the decoder prompt length is 0 and only NEW ids are returned).  Restated:
      * beam_scores start at [0, -1e9, ...]; every step: log_softmax of the fp32 logits of each beam + its beam score, the 2 x num_beams
        best (score, beam, token) of the flattened [num_beams x vocab] table, walked in rank order;
      * an EOS candidate of rank < num_beams closes a hypothesis: score = sum_logprobs / (generated_len ** length_penalty) with
        generated_len counting the EOS; a hypothesis list keeps the num_beams best; other candidates fill the next beams until
        num_beams are taken;
      * the prompt is done when the list is full and (early_stopping or worst kept score >= best running sum_logprobs /
        cur_len ** length_penalty), cur_len counting the token just chosen; otherwise at max_new_tokens the running beams are
        added as hypotheses (no EOS) and the best hypothesis is returned, with the EOS re-appended when it is shorter than the longest
        allowed output (finalize).
    Ties in the top-k are broken towards the lower flat index (beam-major); torch.topk leaves them unspecified."""
    eos = []
    if eos_token_id is not None:
        eos = list(eos_token_id) if isinstance(eos_token_id, (list, tuple, set)) else [int(eos_token_id)]
    k, V = num_beams, cfg.vocab
    emb = w_llm["model.embed_tokens.weight"]
    logits, cache = llama_forward(cfg, w_llm, inputs_embeds, None, dtype)
    last = [logits[-1]] * k
    caches = [cache] * k  # llama_forward never mutates a cache it is given (it returns a new list of concatenated tensors)
    seqs: List[List[int]] = [[] for _ in range(k)]
    beam_scores = torch.full((k,), -1e9)
    beam_scores[0] = 0.0
    hyps: List = []  # (score, tokens)
    worst = 1e9
    done = False
    trace = []
    for step in range(max_new_tokens):
        table = torch.stack([F.log_softmax(x.float(), dim=-1) for x in last]) + beam_scores[:, None]
        flat = table.view(-1)
        # top 2k, ties towards the lower flat index (stable sort of the negated scores)
        order = torch.sort(-flat, stable=True).indices[: max(2, 1 + len(eos)) * k]
        cur_len = step + 1  # tokens generated once this step's choice is appended (decoder prompt length is 0)
        nxt = []
        for rank, fi in enumerate(order.tolist()):
            b, t, sc = fi // V, fi % V, float(flat[fi])
            if t in eos:
                if rank >= k:
                    continue
                score = sc / (cur_len ** length_penalty)
                if len(hyps) < k or score > worst:
                    hyps.append((score, list(seqs[b])))
                    if len(hyps) > k:
                        hyps.remove(min(hyps, key=lambda h: h[0]))
                    worst = min(h[0] for h in hyps)
            else:
                nxt.append((sc, b, t))
            if len(nxt) == k:
                break
        assert len(nxt) == k
        if return_trace:
            trace.append([(round(sc, 6), b, t) for sc, b, t in nxt])
        best_running = float(flat[order[0]])
        if len(hyps) >= k and (early_stopping or worst >= best_running / (cur_len ** length_penalty)):
            done = True
        seqs = [seqs[b] + [t] for _, b, t in nxt]
        beam_scores = torch.tensor([sc for sc, _, _ in nxt])
        if done or step == max_new_tokens - 1:
            break
        parents = [b for _, b, _ in nxt]
        new_last, new_caches = [], []
        for i, (_, b, t) in enumerate(nxt):
            lg, c = llama_forward(cfg, w_llm, emb[t][None].to(dtype), caches[b], dtype)
            new_last.append(lg[-1])
            new_caches.append(c)
        last, caches = new_last, new_caches
    if not done:  # finalize: the running beams become hypotheses, scored over their generated length
        for i in range(k):
            score = float(beam_scores[i]) / (len(seqs[i]) ** length_penalty)
            if len(hyps) < k or score > worst:
                hyps.append((score, list(seqs[i])))
                if len(hyps) > k:
                    hyps.remove(min(hyps, key=lambda h: h[0]))
                worst = min(h[0] for h in hyps)
    best = max(hyps, key=lambda h: h[0])
    out = list(best[1])
    if len(out) < max_new_tokens and eos:  # beam_search.py finalize: the EOS is written back when the hypothesis is shorter than the output
        out.append(eos[0])
    ids = torch.tensor(out, dtype=torch.long)
    if return_trace:
        return ids, trace, best[0]
    return ids


# ----------------------------------------------------------------------------------------------
# whole request
# ----------------------------------------------------------------------------------------------

def generate(cfg: OracleConfig, weights, input_ids: torch.Tensor, images: torch.Tensor,
             depths: Optional[torch.Tensor], masks, max_new_tokens: int, eos_token_id=None,
             dtype: torch.dtype = torch.float32, return_all: bool = False):
    """``LlavaLlamaModel.generate`` (llava_llama.py:194-213) for batch 1 (the reference's only
    shipped mode, eval_region_cls.py:269)."""
    assert input_ids.dim() == 2 and input_ids.shape[0] == 1
    enc = encode_multimodal(cfg, weights, images, depths, masks, dtype)
    emb_table = weights["llm"]["model.embed_tokens.weight"].to(dtype)
    embeds = splice_embeddings(cfg, emb_table, input_ids, enc["image_features"], enc["mask_embeds"],
                               enc["depth_embeds"], None, depths_given=depths is not None)[0]
    ids, logits = greedy_generate(cfg, weights["llm"], embeds, max_new_tokens, eos_token_id, dtype, return_logits=True)
    if return_all:
        enc["inputs_embeds"] = embeds
        enc["logits"] = logits
        return ids, enc
    return ids


# ----------------------------------------------------------------------------------------------
# depth map preparation (eval_spatial.py:92-106)
# ----------------------------------------------------------------------------------------------

def depth_to_u8x3(depth: torch.Tensor, out_h: int, out_w: int) -> torch.Tensor:
    """``get_depth_map`` post-processing (llava/eval/eval_spatial.py:99-105): bilinear resize
    (align_corners=False) of the [1,h,w] fp32 depth to (H,W); min-max normalise * 255; uint8
    (truncation, numpy ``astype``); replicate to 3 channels -> [H, W, 3] u8."""
    d = F.interpolate(depth[None].float(), (out_h, out_w), mode="bilinear", align_corners=False)[0, 0]
    d = (d - d.min()) / (d.max() - d.min()) * 255.0
    d8 = d.to(torch.uint8)
    return d8[..., None].expand(out_h, out_w, 3).contiguous()


# ----------------------------------------------------------------------------------------------
# synthetic request
# ----------------------------------------------------------------------------------------------

def synth_request(cfg: OracleConfig, n_regions: int, t_text: int, seed: int = 1234, kind: str = "mask"):
    g = torch.Generator().manual_seed(seed)
    R = cfg.image_size
    images = torch.rand(1, 3, R, R, generator=g) * 2 - 1
    d1 = torch.rand(1, 1, R, R, generator=g) * 2 - 1
    depths = d1.expand(1, 3, R, R).contiguous()
    masks = torch.zeros(n_regions, R, R)
    for m in range(n_regions):
        side_h = int(torch.randint(R // 8, R // 2 + 1, (1,), generator=g))
        side_w = int(torch.randint(R // 8, R // 2 + 1, (1,), generator=g))
        y0 = int(torch.randint(0, R - side_h + 1, (1,), generator=g))
        x0 = int(torch.randint(0, R - side_w + 1, (1,), generator=g))
        box = torch.zeros(R, R)
        box[y0:y0 + side_h, x0:x0 + side_w] = 1
        if kind == "mask":
            lo = max(R // 16, 2)
            noise = torch.rand(1, 1, lo, lo, generator=g)
            blob = (F.interpolate(noise, (R, R), mode="bilinear", align_corners=False)[0, 0] > 0.45).float()
            box = box * blob
            if box.sum() == 0:
                box[y0, x0] = 1
        masks[m] = box
    hi = min(30000, cfg.vocab - 8)
    lo_id = min(1000, hi // 4)
    ids = [1] + torch.randint(lo_id, hi, (t_text - 1,), generator=g).tolist()
    ids[8] = IMAGE_TOKEN_INDEX
    p = 10
    for m in range(n_regions):
        ids[p] = cfg.mask_token_id
        p += 1
        if cfg.enable_depth:
            ids[p] = cfg.depth_token_id
            p += 1
        p += 2
    assert p <= t_text, "t_text too short for the requested regions"
    input_ids = torch.tensor([ids], dtype=torch.long)
    return input_ids, images, depths, [masks]
