"""Vision tower: SigLIP / CLIP forward on the sm_100a kernels.

Mirrors ``VisionTower.forward`` + ``feature_select`` (llava/model/multimodal_encoder/
vision_encoder.py:26-34,115-132), ``SiglipVisionTower`` (siglip_encoder.py:7-17) and ``CLIPVisionTower``
(clip_encoder.py:8-13): returns ``hidden_states[select_layer]`` - all rows for "cls_patch" (SigLIP has no
class token), the rows after the class token for "patch" (CLIP; vision_encoder.py:28-29).
Only the layers that output depends on are executed (select_layer=-2 -> L-1 layers; the reference
runs all L layers, the post-layernorm and the pooling head and throws them away).
"""
from __future__ import annotations

from types import SimpleNamespace

import torch

from . import ops
from .config import LlavaConfig
from .weights import VisionW, patch_ldk


def tower_layers_needed(num_hidden_layers: int, select_layer: int) -> int:
    # hidden_states = (embeddings, layer_1, ..., layer_L)
    idx = select_layer if select_layer >= 0 else num_hidden_layers + 1 + select_layer
    if not 0 <= idx <= num_hidden_layers:
        raise ValueError(f"select_layer {select_layer} out of range for {num_hidden_layers} layers")
    return idx


class VisionTower:
    def __init__(self, cfg: LlavaConfig, w: VisionW, image_processor=None):
        self.cfg = cfg
        self.vc = cfg.vision
        self.w = w
        self.select_layer = cfg.mm_vision_select_layer
        self.select_feature = cfg.mm_vision_select_feature
        if self.select_feature not in ("cls_patch", "patch"):
            raise ValueError(f"Unexpected select feature: {self.select_feature}")  # vision_encoder.py:33
        acts = {"gelu_pytorch_tanh": ops.EPI_BIAS_GELU_TANH, "quick_gelu": ops.EPI_BIAS_QUICK_GELU, "gelu": ops.EPI_BIAS_GELU_ERF}
        if self.vc.hidden_act not in acts:
            raise NotImplementedError(f"vision hidden_act {self.vc.hidden_act!r} (have {sorted(acts)})")
        self._fc1_epilogue = acts[self.vc.hidden_act]
        if self.vc.is_clip and (w.cls_emb is None or w.pre_ln_w is None):
            raise ValueError("a CLIP tower needs class_embedding and pre_layrnorm weights")
        self.n_layers = tower_layers_needed(self.vc.num_hidden_layers, self.select_layer)
        if len(w.layers) < self.n_layers:
            raise ValueError(f"need {self.n_layers} vision layers, weights hold {len(w.layers)}")
        self.is_loaded = True
        self.image_processor = image_processor
        self._layer_array = ops.make_siglip_layer_array(w.layers[: self.n_layers])  # one C call runs the whole stack
        # builder.py:190-192 records the special token ids here
        self.config = SimpleNamespace(llm_mask_token_id=cfg.llm_mask_token_id, llm_depth_token_id=cfg.llm_depth_token_id,
                                      hidden_size=self.vc.hidden_size, image_size=self.vc.image_size,
                                      patch_size=self.vc.patch_size)

    @property
    def device(self):
        return self.w.patch_w.device

    @property
    def dtype(self):
        return self.w.patch_w.dtype

    @property
    def hidden_size(self):
        return self.vc.hidden_size

    @property
    def num_patches(self):
        return self.vc.grid ** 2

    @ops.in_own_dtype
    def forward(self, images: torch.Tensor) -> torch.Tensor:
        """images [N, 3, R, R] (fp32 or the model dtype, on the device) -> [N, T, D] in the model dtype."""
        if isinstance(images, (list, tuple)):  # vision_encoder.py:116-125
            return [self.forward(im.unsqueeze(0) if im.dim() == 3 else im) for im in images]
        vc, w = self.vc, self.w
        if images.dim() != 4 or images.shape[-1] != vc.image_size or images.shape[-2] != vc.image_size:
            raise ValueError(f"expected images [N, 3, {vc.image_size}, {vc.image_size}], got {tuple(images.shape)}")
        images = images.to(device=self.device)
        if images.dtype not in (torch.float32, self.dtype):
            images = images.float()
        images = images.contiguous()
        N, T, D, nh, hd = images.shape[0], vc.grid ** 2, vc.hidden_size, vc.num_attention_heads, vc.head_dim
        a = ops.patchify(images, vc.patch_size, patch_ldk(vc.patch_size))
        if vc.is_clip:
            # CLIPVisionEmbeddings: bias-free convolution, class token prepended, position embedding added; then pre_layrnorm
            x = ops.clip_embed(ops.gemm(a, w.patch_w), w.cls_emb, w.pos_emb, N, T)
            x = ops.layernorm(x, w.pre_ln_w, w.pre_ln_b, vc.layer_norm_eps, out=x)
            T += 1
        else:
            x = ops.gemm(a, w.patch_w, bias=w.patch_b, residual=w.pos_emb, epilogue=ops.EPI_BIAS_RESIDUAL, res_row_mod=T)
        ops.siglip_layers(x, self._layer_array, self.n_layers, N, T, D, nh, vc.intermediate_size, vc.layer_norm_eps, self._fc1_epilogue)
        x = x.view(N, T, D)
        if self.select_feature == "patch":  # vision_encoder.py:28-29: drop row 0 (a plain strided device copy)
            x = x[:, 1:].contiguous()
        return x

    __call__ = forward


SiglipVisionTower = VisionTower
