"""mm_projector (llava/model/multimodal_projector/base_projector.py:55-94).

``mlp_downsample`` (the SpatialRGPT / VILA-1.5 type, :73-80): DownSampleBlock + LayerNorm(4C) fused in one gather kernel, then two
tcgen05 GEMMs (GELU-erf fused into the first epilogue).  The remaining reference types are the same GEMM kernel with other
epilogues: ``linear`` (:71-72), ``mlpNx_gelu`` (:81-88: Linear, then N-1 x (GELU, Linear)) and ``identity`` (:69-70)."""
from __future__ import annotations

import re

import torch

from . import ops
from .config import LlavaConfig
from .weights import ProjectorW


class MultimodalProjector:
    def __init__(self, cfg: LlavaConfig, w: ProjectorW):
        t = cfg.mm_projector_type
        if not (t in ("mlp_downsample", "linear", "identity") or re.match(r"^mlp(\d+)x_gelu$", t or "")):
            raise ValueError(f"Unknown projector type: {t}")  # base_projector.py:91
        self.cfg = cfg
        self.w = w
        self.downsamples = t == "mlp_downsample"

    def tokens_out(self, side: int) -> int:
        """Rows one image's side x side feature grid becomes."""
        return ((side + 1) // 2) ** 2 if self.downsamples else side * side

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """[N, side*side, C] -> [N, tokens_out(side), H]."""
        w = self.w
        if self.downsamples:
            x = ops.downsample_layernorm(x.contiguous(), w.ln_w, w.ln_b, 1e-5)
            N, T4, C4 = x.shape
            h = ops.gemm(x.view(N * T4, C4), w.fc1_w, bias=w.fc1_b, epilogue=ops.EPI_BIAS_GELU_ERF)
            o = ops.gemm(h, w.fc2_w, bias=w.fc2_b, epilogue=ops.EPI_BIAS)
            return o.view(N, T4, -1)
        if not w.linears:  # identity
            return x
        N, T, C = x.shape
        h = x.contiguous().view(N * T, C)
        for i, (lw, lb) in enumerate(w.linears):  # GELU (erf, nn.GELU()) follows every linear but the last
            h = ops.gemm(h, lw, bias=lb, epilogue=ops.EPI_BIAS if i == len(w.linears) - 1 else ops.EPI_BIAS_GELU_ERF)
        return h.view(N, T, -1)

    __call__ = forward
