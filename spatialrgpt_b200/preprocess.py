"""GPU versions of the reference's host preprocessing (``llava/mm_utils.py:421-542``: ``process_image`` / ``process_images`` /
``process_regions``), SURVEY.md §8f.2 - the CPU stage in front of TTFT.

The arithmetic is third-party and pinned by the reference (``pyproject.toml:17`` transformers==4.37.2, whose SigLIP / CLIP image
processors are the "slow", Pillow-based ones):

  images  PIL ``Image.resize((R, R), BICUBIC)`` on uint8  ->  ``image * rescale_factor`` (float64) -> float32  ->
          ``(x - mean) / std`` (float32)  ->  channels first                      (image_processing_siglip.py; mm_utils.py:431-470)
  masks   ``cv2.resize(m, (R, R), INTER_NEAREST)`` (mm_utils.py:521-523); the processor call after it is a same-size resize = copy,
          without rescale / normalisation (479-482)  ->  float [M, R, R]

Pillow's resampler (libImaging/Resample.c) is integer arithmetic over per-output-pixel coefficient windows.  ``resample_coeffs``
rebuilds those tables with the same double arithmetic (the bicubic kernel a = -0.5, support 2 x max(scale, 1), window bounds,
normalisation, 22-bit fixed point with round-half-away); the kernels in csrc/preprocess.cu then apply them, horizontal pass first,
8-bit intermediate, exactly like ``ImagingResample``.  Results are bit-identical to Pillow / OpenCV (tests/test_preprocess_cpu.py
checks the tables through a numpy emulation against Pillow itself; tests/test_gpu_ops.py the kernels).

Note: the container's transformers 5.5 implements the same processors on torchvision; its bicubic differs from Pillow's by one
8-bit step on ~1 % of the pixels.  This module follows the PINNED behaviour (Pillow).
"""
from __future__ import annotations

import math
from functools import lru_cache
from typing import List, Sequence, Tuple

import numpy as np
import torch

PRECISION_BITS = 32 - 8 - 2


def _bicubic(x: np.ndarray, a: float = -0.5) -> np.ndarray:
    x = np.abs(x)
    near = ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    far = (((x - 5) * x + 8) * x - 4) * a
    return np.where(x < 1.0, near, np.where(x < 2.0, far, 0.0))


@lru_cache(maxsize=64)
def resample_coeffs(in_size: int, out_size: int) -> Tuple[np.ndarray, np.ndarray, int]:
    """Pillow ``precompute_coeffs`` + ``normalize_coeffs_8bpc`` for the bicubic filter over the whole axis.
    Returns (kk int32 [out_size, ksize], bounds int32 [out_size, 2] = (xmin, count), ksize)."""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    ss = 1.0 / filterscale
    xx = np.arange(out_size, dtype=np.float64)
    center = (xx + 0.5) * scale
    xmin = np.maximum((center - support + 0.5).astype(np.int64), 0)          # C (int) cast: truncation (arguments are > -1)
    xmax = np.minimum((center + support + 0.5).astype(np.int64), in_size)
    cnt = xmax - xmin
    k = np.arange(ksize, dtype=np.float64)[None, :]
    w = _bicubic((k + xmin[:, None] - center[:, None] + 0.5) * ss)
    w = np.where(k < cnt[:, None], w, 0.0)
    ww = np.cumsum(w, axis=1)[:, -1]                                          # sequential double sum, like the C loop
    w = np.where(ww[:, None] != 0.0, w / np.where(ww == 0.0, 1.0, ww)[:, None], w)
    fixed = np.where(w < 0, -0.5 + w * (1 << PRECISION_BITS), 0.5 + w * (1 << PRECISION_BITS)).astype(np.int64).astype(np.int32)  # (int): truncation
    fixed = np.where(k < cnt[:, None], fixed, 0).astype(np.int32)
    bounds = np.stack([xmin, cnt], axis=1).astype(np.int32)
    return np.ascontiguousarray(fixed), np.ascontiguousarray(bounds), ksize


def resample_reference_numpy(img: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """The two Pillow passes in numpy integer arithmetic over ``resample_coeffs`` (the CPU statement of what the kernels do;
    test infrastructure for the tables).  img uint8 [H, W, C]."""
    def one_axis(a: np.ndarray, axis: int, out_size: int) -> np.ndarray:
        kk, bounds, ksize = resample_coeffs(a.shape[axis], out_size)
        a = np.moveaxis(a, axis, 0).astype(np.int64)
        out = np.empty((out_size,) + a.shape[1:], dtype=np.uint8)
        for i in range(out_size):
            lo, n = int(bounds[i, 0]), int(bounds[i, 1])
            acc = (1 << (PRECISION_BITS - 1)) + np.tensordot(kk[i, :n].astype(np.int64), a[lo:lo + n], axes=(0, 0))
            out[i] = np.clip(acc >> PRECISION_BITS, 0, 255)
        return np.moveaxis(out, 0, axis)
    H, W = img.shape[:2]
    if out_w != W:
        img = one_axis(img, 1, out_w)
    if out_h != H:
        img = one_axis(img, 0, out_h)
    return img


@lru_cache(maxsize=64)
def nearest_indices(in_size: int, out_size: int) -> np.ndarray:
    """cv2.resize INTER_NEAREST source indices: min(floor(x * (1 / (out / in))), in - 1) in double arithmetic (resizeNN)."""
    ifx = 1.0 / (out_size / in_size)
    return np.minimum(np.floor(np.arange(out_size, dtype=np.float64) * ifx).astype(np.int64), in_size - 1).astype(np.int32)


def _dev_table(arr: np.ndarray, device) -> torch.Tensor:
    return torch.from_numpy(arr).to(device, non_blocking=True)


def _processor_params(image_processor):
    size = getattr(image_processor, "crop_size", None) or image_processor.size
    if "height" in size:
        R_h, R_w = int(size["height"]), int(size["width"])
    else:
        R_h = R_w = int(size["shortest_edge"])
    resample = getattr(image_processor, "resample", 3)
    if int(resample) != 3:
        raise NotImplementedError(f"GPU preprocessing implements the BICUBIC resample of the SigLIP / CLIP processors, got resample={resample}")
    return R_h, R_w


def resize_bicubic_u8(img: torch.Tensor, out_h: int, out_w: int) -> torch.Tensor:
    """uint8 [H, W, C] on the GPU -> uint8 [out_h, out_w, C], bit-identical to PIL ``Image.resize((out_w, out_h), BICUBIC)``."""
    from . import ops
    H, W, C = img.shape
    cur = img.contiguous()
    if out_w != W:
        kk, bounds, ksize = resample_coeffs(W, out_w)
        cur = ops.resample_u8(cur, 1, out_w, _dev_table(kk, img.device), _dev_table(bounds, img.device), ksize)
    if out_h != H:
        kk, bounds, ksize = resample_coeffs(H, out_h)
        cur = ops.resample_u8(cur, 0, out_h, _dev_table(kk, img.device), _dev_table(bounds, img.device), ksize)
    return cur


def process_images_gpu(images: Sequence, image_processor, model_cfg, device) -> torch.Tensor:
    """``process_images`` (mm_utils.py:535-542) with the resize / rescale / normalise on the GPU.  ``images``: PIL images or uint8
    arrays [H, W, 3].  Returns float32 [N, 3, R, R] on ``device``, bit-identical to the pinned CPU path."""
    from . import ops
    from .mm_utils import _expand2square
    R_h, R_w = _processor_params(image_processor)
    mode = getattr(model_cfg, "image_aspect_ratio", None)
    mean = [float(v) for v in image_processor.image_mean]
    std = [float(v) for v in image_processor.image_std]
    out = []
    for im in images:
        if not isinstance(im, np.ndarray):
            im = im.convert("RGB")
            if mode == "pad":
                im = _expand2square(im, tuple(int(x * 255) for x in image_processor.image_mean))
            im = np.asarray(im)
        elif mode == "pad":
            raise NotImplementedError("pad mode needs PIL inputs")
        # "resize" mode resizes to (R, R) with PIL's default filter (BICUBIC) first (mm_utils.py:441); the processor's own resize to
        # the same size is then a copy.  Without it the processor resizes directly: the same single bicubic pass either way.
        d = torch.from_numpy(np.ascontiguousarray(im)).to(device, non_blocking=True)
        u8 = resize_bicubic_u8(d, R_h, R_w)
        out.append(ops.u8_to_normalized_chw(u8, float(image_processor.rescale_factor), mean, std, bool(getattr(image_processor, "do_normalize", True))))
    return torch.stack(out, 0)


def process_regions_gpu(masks: Sequence[np.ndarray], image_processor, data_args, device) -> torch.Tensor:
    """``process_regions`` (mm_utils.py:477-532) in "resize" mode on the GPU: uint8 masks [H, W] -> float32 [M, R, R]."""
    from . import ops
    mode = getattr(data_args, "image_aspect_ratio", None)
    if mode != "resize":
        raise NotImplementedError("GPU region preprocessing covers image_aspect_ratio='resize' (all SpatialRGPT scripts); use process_regions otherwise")
    R_h, R_w = _processor_params(image_processor)
    out = []
    for m in masks:
        m = np.ascontiguousarray(np.asarray(m, dtype=np.uint8))
        H, W = m.shape
        d = torch.from_numpy(m).to(device, non_blocking=True)
        out.append(ops.resize_nearest_u8(d, R_h, R_w, _dev_table(nearest_indices(H, R_h), device), _dev_table(nearest_indices(W, R_w), device)))
    return torch.stack(out, 0)
