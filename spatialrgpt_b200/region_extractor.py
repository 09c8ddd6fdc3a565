"""Region extractor on the sm_100a kernels (reference: llava/model/region_extractor/
base_extractor.py:112-173, type "regiongpt").

feature_refinement: ConvT(k2,s2) -> LayerNorm2d -> GELU -> ConvT(k2,s2) -> GELU as two tcgen05
GEMMs + one row LayerNorm kernel.  The up-sampled feature map stays in the *nested* pixel order the
GEMMs emit (DESIGN.md "hres layout"); mask pooling and AdaptiveAvgPool2d(27) index it directly, so the
37.7 MB/image tensor is written once and read once per consumer with no pixel-shuffle pass.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch

from . import ops
from .config import LlavaConfig
from .weights import RegionW

ADA_POOL = 27  # base_extractor.py:123 ("hardcoded pooling size")


class MaskPooling:
    """base_extractor.py:27-84.  ``x`` [B, L, C]; ``order`` is the row order of x."""

    def forward(self, x: torch.Tensor, mask_list: Optional[Sequence[Optional[torch.Tensor]]], order: int = ops.ORDER_ROWMAJOR,
                return_list: bool = True) -> List[Optional[torch.Tensor]]:
        B, L, _ = x.shape
        side = int(round(L ** 0.5))
        if mask_list is None:
            mask_list = [None] * B
        out: List[Optional[torch.Tensor]] = [None] * B
        idx = [i for i in range(B) if mask_list[i] is not None]
        if not idx:
            return out
        pre = getattr(self, "_pre", None)
        if pre is not None and pre[0] is mask_list and pre[1] == B and (side, order) in pre[2]:
            # weights were computed on the side stream while the tower ran (precompute): order the pooling after them
            cur = torch.cuda.current_stream(x.device)
            cur.wait_event(pre[3])
            kind, wts = pre[2][(side, order)]
            if kind == "stack":
                wts.record_stream(cur)
                pooled = ops.mask_pool(x, wts)
                for j, i in enumerate(idx):
                    out[i] = pooled[j]
            else:
                for i in idx:
                    wts[i].record_stream(cur)
                    out[i] = ops.mask_pool(x[i:i + 1], wts[i])[0]
        else:
            kind, wts = self._weights(mask_list, idx, B, side, order, x.device)
            if kind == "stack":
                pooled = ops.mask_pool(x, wts)
                for j, i in enumerate(idx):
                    out[i] = pooled[j]
            else:
                for i in idx:
                    out[i] = ops.mask_pool(x[i:i + 1], wts[i])[0]
        if not return_list:
            return torch.cat([o for o in out if o is not None])
        return out

    def _weights(self, mask_list, idx, B: int, side: int, order: int, device):
        """Normalised bf16 pooling weights: one [B, M, L] tensor when every image has the same number of regions, else per image."""
        same = all(mask_list[i].shape == mask_list[idx[0]].shape and mask_list[i].dtype == mask_list[idx[0]].dtype for i in idx)
        if same and len(idx) == B and B > 1:
            return "stack", ops.mask_weights(torch.stack([self._prep(mask_list[i], device) for i in idx], 0), side, order)
        return "list", {i: ops.mask_weights(self._prep(mask_list[i], device)[None], side, order) for i in idx}

    def precompute(self, mask_list, n_images: int, specs, device) -> None:
        """The pooling weights depend on the masks alone (base_extractor.py:53-66), not on any feature map: compute them for the
        given (side, order) geometries on a SIDE stream now - the caller launches the vision tower next, so the resampling /
        normalisation kernels run under it instead of between the refinement and the pooling.  ``forward`` picks them up."""
        self._pre = None
        if mask_list is None:
            return
        idx = [i for i in range(min(n_images, len(mask_list))) if mask_list[i] is not None]
        if not idx or len(mask_list) != n_images:
            return
        cur = torch.cuda.current_stream(device)
        if getattr(self, "_stream", None) is None:
            self._stream = torch.cuda.Stream(device=device)
        self._stream.wait_stream(cur)  # the masks (e.g. their host->device copies) are ordered before the side work
        with torch.cuda.stream(self._stream):
            table = {(side, order): self._weights(mask_list, idx, n_images, side, order, device) for side, order in specs}
            ev = torch.cuda.Event()
            ev.record(self._stream)
        self._pre = (mask_list, n_images, table, ev)

    @staticmethod
    def _prep(mask: torch.Tensor, device) -> torch.Tensor:
        m = mask.detach().to(device=device)
        if m.dtype not in (torch.float32, ops.ELEM()):
            m = m.float()  # base_extractor.py:55
        return m.contiguous()

    __call__ = forward


class RegionExtractor:
    def __init__(self, cfg: LlavaConfig, w: RegionW):
        if cfg.region_extractor_type != "regiongpt":
            raise NotImplementedError(f"{cfg.region_extractor_type} not implemented")  # base_extractor.py:161
        self.cfg = cfg
        self.w = w
        self.mask_pooling = MaskPooling()
        self.C = cfg.vision.hidden_size

    @property
    def dtype(self):
        return self.w.rgb_w.dtype

    @ops.in_own_dtype
    def feature_refinement_nested(self, tower_features: torch.Tensor):
        """[N, T, C] -> (hres in nested order [N, 16T, C], lres [N, 729, C] row-major)."""
        N, T, C = tower_features.shape
        P = int(round(T ** 0.5))
        w = self.w
        x = tower_features.reshape(N * T, C)
        y1 = ops.gemm(x, w.deconv1_w, bias=w.deconv1_b, epilogue=ops.EPI_BIAS)  # [N*T, 4C] == [N*4T, C]
        y1 = ops.layernorm(y1.view(N * T * 4, C), w.ln_w, w.ln_b, 1e-6, act=1)  # LayerNorm2d + GELU
        y2 = ops.gemm(y1, w.deconv2_w, bias=w.deconv2_b, epilogue=ops.EPI_BIAS_GELU_ERF)  # [N*4T, 4C] == [N*16T, C]
        hres = y2.view(N, 16 * T, C)
        lres = ops.adaptive_avgpool(hres, 4 * P, ADA_POOL, ops.ORDER_NESTED)
        return hres, lres

    @ops.in_own_dtype
    def feature_refinement(self, tower_features: torch.Tensor):
        """Reference signature/layout (base_extractor.py:137-147): hres flattened row-major (H W)."""
        hres, lres = self.feature_refinement_nested(tower_features)
        side = 4 * int(round(tower_features.shape[1] ** 0.5))
        return ops.reorder_rows(hres, side, ops.ORDER_NESTED, ops.ORDER_ROWMAJOR), lres

    def extract_region_features(self, features: torch.Tensor, masks, proj_w, proj_b, order: int):
        pooled = self.mask_pooling(features, masks, order=order, return_list=True)
        idx = [i for i, p in enumerate(pooled) if p is not None]
        out: List[Optional[torch.Tensor]] = [None] * len(pooled)
        if idx:  # one projector GEMM over the regions of the whole batch
            rows = [pooled[i].shape[0] for i in idx]
            proj = ops.gemm(torch.cat([pooled[i] for i in idx], 0), proj_w, bias=proj_b, epilogue=ops.EPI_BIAS)
            for i, part in zip(idx, torch.split(proj, rows, 0)):
                out[i] = part
        return out

    @ops.in_own_dtype
    def forward(self, image_features: torch.Tensor, depth_features: Optional[torch.Tensor], masks, hres_order: int = ops.ORDER_ROWMAJOR):
        """base_extractor.py:167-173.  ``image_features`` = hres (row-major by default, like the reference)."""
        w = self.w
        mask_embeds = self.extract_region_features(image_features, masks, w.rgb_w, w.rgb_b, hres_order)
        depth_embeds = None
        if depth_features is not None:
            depth_embeds = self.extract_region_features(depth_features, masks, w.depth_w, w.depth_b, ops.ORDER_ROWMAJOR)
        self.mask_pooling._pre = None  # precomputed weights (if any) belong to this request only
        return mask_embeds, depth_embeds

    __call__ = forward
