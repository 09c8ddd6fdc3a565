"""Dimensions of the models on the generate() path and the reference's
``LlavaConfig`` field names (llava/model/configuration_llava.py:7-59) so that a reference
``config.json`` loads unchanged."""
from __future__ import annotations

import json
import os
from dataclasses import asdict, dataclass, field
from typing import Any, Dict, Optional


@dataclass
class VisionConfig:
    """SigLIP or CLIP vision transformer (HF ``SiglipVisionConfig`` / ``CLIPVisionConfig`` field names).  ``model_type``
    "clip_vision_model" adds the class token, the pre-layernorm and the bias-free patch convolution of HF CLIPVisionModel
    (clip_encoder.py:8-13)."""
    image_size: int = 448
    patch_size: int = 14
    hidden_size: int = 1152
    num_hidden_layers: int = 27
    num_attention_heads: int = 16
    intermediate_size: int = 4304
    layer_norm_eps: float = 1e-6
    hidden_act: str = "gelu_pytorch_tanh"
    model_type: str = "siglip_vision_model"

    @property
    def grid(self) -> int:
        return self.image_size // self.patch_size

    @property
    def is_clip(self) -> bool:
        return "clip" in self.model_type and "siglip" not in self.model_type

    @property
    def tokens(self) -> int:
        """Rows per image inside the encoder (CLIP: patches + the class token)."""
        return self.grid * self.grid + (1 if self.is_clip else 0)

    @property
    def head_dim(self) -> int:
        return self.hidden_size // self.num_attention_heads


@dataclass
class LlamaDims:
    """Llama decoder (HF ``LlamaConfig`` field names)."""
    hidden_size: int = 4096
    num_hidden_layers: int = 32
    num_attention_heads: int = 32
    num_key_value_heads: int = 8
    head_dim: int = 128
    intermediate_size: int = 14336
    vocab_size: int = 128259
    rope_theta: float = 500000.0
    rope_scaling_factor: float = 1.0  # "linear" RoPE scaling (positions / factor; language_model/builder.py:31-38, modeling_llama.py:133-141)
    rms_norm_eps: float = 1e-5
    max_position_embeddings: int = 8192
    bos_token_id: Optional[int] = None
    eos_token_id: Any = None
    pad_token_id: Optional[int] = None
    tokenizer_model_max_length: Optional[int] = None
    tokenizer_padding_side: str = "right"


@dataclass
class LlavaConfig:
    """Top-level VLM config; same field names as the reference ``LlavaConfig``."""
    model_type: str = "llava_llama"
    architectures: tuple = ("LlavaLlamaModel",)
    llm_cfg: Any = None
    vision_tower_cfg: Any = None
    mm_projector_cfg: Any = None
    region_extractor_cfg: Any = None
    resume_path: Optional[str] = None
    hidden_size: Optional[int] = None
    mm_hidden_size: Optional[int] = None
    image_aspect_ratio: Optional[str] = "resize"
    num_video_frames: Optional[int] = None
    mm_vision_select_layer: int = -2
    mm_vision_select_feature: str = "cls_patch"
    mm_use_im_start_end: bool = False
    mm_use_im_patch_token: bool = False
    mm_projector_lr: Optional[float] = None
    vision_resolution: Optional[int] = None
    interpolate_mode: Optional[str] = None
    s2: Optional[bool] = None
    s2_scales: Optional[str] = None
    s2_max_split_size: Optional[int] = None
    enable_region: bool = True
    enable_depth: bool = True
    model_dtype: str = "torch.float16"  # llava_arch.py:74; overwritten with the dtype the weights were loaded / cast to
    # ours (not in the reference): resolved sub-configs and special-token ids
    vision: VisionConfig = field(default_factory=VisionConfig)
    llama: LlamaDims = field(default_factory=LlamaDims)
    mm_projector_type: str = "mlp_downsample"
    region_extractor_type: str = "regiongpt"
    llm_mask_token_id: int = -1
    llm_depth_token_id: int = -1

    def __post_init__(self):
        if self.hidden_size is None:
            self.hidden_size = self.llama.hidden_size
        if self.mm_hidden_size is None:
            self.mm_hidden_size = self.vision.hidden_size

    def to_dict(self) -> Dict[str, Any]:
        d = asdict(self)
        d["architectures"] = list(self.architectures)
        return d

    @classmethod
    def from_json_file(cls, path: str) -> "LlavaConfig":
        with open(path) as f:
            raw = json.load(f)
        known = {k: v for k, v in raw.items() if k in cls.__dataclass_fields__ and k not in ("vision", "llama")}
        if "architectures" in known:
            known["architectures"] = tuple(known["architectures"])
        cfg = cls(**known)
        if isinstance(raw.get("vision"), dict):
            cfg.vision = VisionConfig(**raw["vision"])
        if isinstance(raw.get("llama"), dict):
            cfg.llama = LlamaDims(**raw["llama"])
        cfg.hidden_size = raw.get("hidden_size") or cfg.llama.hidden_size
        cfg.mm_hidden_size = raw.get("mm_hidden_size") or cfg.vision.hidden_size
        return cfg


LlavaLlamaConfig = LlavaConfig  # llava/model/language_model/llava_llama.py:43


# ---- the BASELINE.json configurations -------------------------------------
def siglip_so400m(image_size: int = 448) -> VisionConfig:
    return VisionConfig(image_size=image_size)


def llama3_8b() -> LlamaDims:
    return LlamaDims(vocab_size=128259, eos_token_id=None)


def llama2_7b() -> LlamaDims:
    return LlamaDims(num_key_value_heads=32, intermediate_size=11008, vocab_size=32002, rope_theta=10000.0,
                     max_position_embeddings=4096)


def sheared_llama_2p7b() -> LlamaDims:
    return LlamaDims(hidden_size=2560, num_attention_heads=20, num_key_value_heads=20, intermediate_size=6912,
                     vocab_size=32002, rope_theta=10000.0, max_position_embeddings=4096)


def baseline_config(name: str) -> LlavaConfig:
    """c1..c5 of BASELINE.json."""
    if name in ("c2", "c3", "c5", "llama3_8b"):
        ll, px = llama3_8b(), 448
    elif name in ("c4", "llama2_7b"):
        ll, px = llama2_7b(), 448
    elif name in ("c1", "sheared_3b"):
        ll, px = sheared_llama_2p7b(), 336
    else:
        raise ValueError(f"unknown baseline config {name!r}")
    cfg = LlavaConfig(vision=siglip_so400m(px), llama=ll)
    cfg.llm_mask_token_id = ll.vocab_size - 2
    cfg.llm_depth_token_id = ll.vocab_size - 1
    return cfg
