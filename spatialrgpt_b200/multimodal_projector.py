"""mm_projector ``mlp_downsample`` (llava/model/multimodal_projector/base_projector.py:32-52,73-80):
DownSampleBlock + LayerNorm(4C) fused in one gather kernel, then two tcgen05 GEMMs (GELU-erf fused)."""
from __future__ import annotations

import torch

from . import ops
from .config import LlavaConfig
from .weights import ProjectorW


class MultimodalProjector:
    def __init__(self, cfg: LlavaConfig, w: ProjectorW):
        if cfg.mm_projector_type != "mlp_downsample":
            raise ValueError(f"Unknown projector type: {cfg.mm_projector_type}")  # base_projector.py:91
        self.cfg = cfg
        self.w = w

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """[N, side*side, C] -> [N, ceil(side/2)^2, H]."""
        w = self.w
        x = ops.downsample_layernorm(x.contiguous(), w.ln_w, w.ln_b, 1e-5)
        N, T4, C4 = x.shape
        h = ops.gemm(x.view(N * T4, C4), w.fc1_w, bias=w.fc1_b, epilogue=ops.EPI_BIAS_GELU_ERF)
        o = ops.gemm(h, w.fc2_w, bias=w.fc2_b, epilogue=ops.EPI_BIAS)
        return o.view(N, T4, -1)

    __call__ = forward
