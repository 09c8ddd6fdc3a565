"""Host-side pre/post-processing API of the reference (llava/mm_utils.py): image / region
preparation, ``<image>`` tokenisation and the keyword stopping criterion.

These run on the CPU in front of the GPU path exactly like the reference's (PIL / cv2 / the HF image
processor), with the same names, argument meaning and output conventions:
  process_images  (mm_utils.py:535-542)   list of PIL images -> float tensor [N, 3, R, R]
  process_regions (mm_utils.py:477-532)   list of uint8 masks [H, W] -> float tensor [M, R, R]
  tokenizer_image_token (545-570)          prompt with "<image>" -> ids with IMAGE_TOKEN_INDEX (-200)
  KeywordsStoppingCriteria (586-617)
"""
from __future__ import annotations

import copy
from typing import List, Optional, Sequence

import numpy as np
import torch

from .constants import IMAGE_TOKEN_INDEX


def _target_size(image_processor) -> dict:
    """CLIP processors carry ``crop_size``, SigLIP ones ``size`` (mm_utils.py:488-495)."""
    if hasattr(image_processor, "crop_size") and image_processor.crop_size is not None:
        return image_processor.crop_size
    assert hasattr(image_processor, "size")
    return image_processor.size


def expand2square(pil_img, background_color):
    """mm_utils.py:249-276: pad to a square with ``background_color`` (its first component for mode "L"), image centred."""
    from PIL import Image

    w, h = pil_img.size
    if pil_img.mode == "L":
        background_color = background_color[0]
    if w == h:
        return pil_img
    side = max(w, h)
    result = Image.new(pil_img.mode, (side, side), background_color)
    result.paste(pil_img, ((side - w) // 2, (side - h) // 2))
    return result


_expand2square = expand2square


def load_image_from_base64(image):
    """mm_utils.py:245-246."""
    import base64
    from io import BytesIO

    from PIL import Image
    return Image.open(BytesIO(base64.b64decode(image)))


def is_gemma_tokenizer(tokenizer) -> bool:
    """mm_utils.py:573-574."""
    return "gemma" in tokenizer.__class__.__name__.lower()


def process_image(image_file, data_args, image_folder=None, return_info: bool = False):
    """mm_utils.py:421-474.  ``image_aspect_ratio``: "resize" (all srgpt scripts), "pad" or the processor default."""
    import os

    from PIL import Image

    processor = data_args.image_processor
    if isinstance(image_file, str):
        path = os.path.join(image_folder, image_file) if image_folder is not None else image_file
        image = Image.open(path)
    else:
        image = image_file
    image = image.convert("RGB")
    ori_w, ori_h = image.size
    mode = getattr(data_args, "image_aspect_ratio", None)
    if mode == "resize":
        size = _target_size(processor)
        image = image.resize((size["height"], size["width"]))
    elif mode == "pad":
        image = _expand2square(image, tuple(int(x * 255) for x in processor.image_mean))
    pixel = processor.preprocess(image, return_tensors="pt")["pixel_values"][0]
    if return_info:
        return pixel, {"width": ori_w, "height": ori_h}
    return pixel


def process_images(images, image_processor, model_cfg, device=None):
    """mm_utils.py:535-542: stacks when every image has the same shape, else returns a list.  With ``device`` (a CUDA device) the
    resize / rescale / normalise run in the sm_100a kernels of preprocess.py (bit-identical to the pinned Pillow-based processor) and
    the result stays on the GPU - the CPU stage in front of TTFT disappears."""
    model_cfg.image_processor = image_processor
    if device is not None and torch.device(device).type == "cuda":
        from .preprocess import process_images_gpu
        return process_images_gpu(images, image_processor, model_cfg, torch.device(device))
    out = [process_image(im, model_cfg, None) for im in images]
    if all(x.shape == out[0].shape for x in out):
        out = torch.stack(out, dim=0)
    return out


def _mask_processor(image_processor):
    """Same processor with normalisation / rescaling / RGB conversion off (mm_utils.py:479-482): masks come
    out as resampled floats (bicubic for SigLIP), not strictly binary."""
    mp = copy.deepcopy(image_processor)
    mp.do_normalize = False
    mp.do_convert_rgb = False
    mp.rescale_factor = 1.0
    return mp


def _pad_to_square(a: np.ndarray) -> np.ndarray:
    h, w = a.shape
    side = max(h, w)
    out = np.zeros((side, side), dtype=np.uint8)
    out[(side - h) // 2:(side - h) // 2 + h, (side - w) // 2:(side - w) // 2 + w] = a
    return out


def process_regions(masks: Sequence[np.ndarray], image_processor, data_args, device=None) -> torch.Tensor:
    """mm_utils.py:477-532: uint8 region masks [H, W] -> float tensor [M, R, R] (on ``device`` through the nearest-neighbour kernel
    when a CUDA device is given and the aspect mode is "resize")."""
    if device is not None and torch.device(device).type == "cuda" and getattr(data_args, "image_aspect_ratio", None) == "resize":
        from .preprocess import process_regions_gpu
        return process_regions_gpu(masks, image_processor, data_args, torch.device(device))
    import cv2

    mp = _mask_processor(image_processor)
    mode = getattr(data_args, "image_aspect_ratio", None)
    prepared = []
    for m in masks:
        m = np.asarray(m)
        if mode == "resize":
            size = _target_size(data_args.image_processor)
            m = cv2.resize(m, (size["width"], size["height"]), interpolation=cv2.INTER_NEAREST)
        elif mode == "pad":
            m = _pad_to_square(m)
        prepared.append(m)
    rows = [mp.preprocess(m[None, ...], return_tensors="pt")["pixel_values"][0] for m in prepared]
    return torch.vstack(rows).float()


def process_masks(sources, data_args, image_info=None) -> torch.Tensor:
    """mm_utils.py:279-375 (the dataset-side twin of process_regions, data/dataset.py:1760): regions given as COCO run-length masks
    ("rle"), polygons ("segmentation") or boxes ("bbox") in ``sources[0]`` -> float tensor [M, R, R].  Like the reference, ONE of
    the modalities present is drawn with ``random.choice`` (the global generator), the uint8 masks are resized with
    cv2.INTER_NEAREST ("resize") or zero-padded to a square ("pad") and then go through the mask processor (no normalisation, no
    rescale).  Run-length masks are decoded by this package's own COCO RLE decoder (eval_spatial.rle_decode); polygons need
    pycocotools' rasteriser and raise without it."""
    import random

    import cv2

    mp = _mask_processor(data_args.image_processor)
    modality = random.choice([m for m in ("rle", "segmentation", "bbox") if m in sources[0].keys()])
    mode = getattr(data_args, "image_aspect_ratio", None)
    size = _target_size(data_args.image_processor) if mode == "resize" else None

    def fit(m: np.ndarray) -> np.ndarray:
        if mode == "resize":
            m = cv2.resize(m, (size["width"], size["height"]), interpolation=cv2.INTER_NEAREST)
        if mode == "pad":
            m = _pad_to_square(m)
        return m

    masks = []
    if modality == "rle":
        from .eval_spatial import rle_decode
        masks = [fit(rle_decode(r).astype(np.uint8)) for r in sources[0]["rle"]]
    elif modality == "segmentation":
        try:
            from pycocotools import mask as cocomask
        except ImportError as e:
            raise NotImplementedError("polygon regions are rasterised by pycocotools (mm_utils.py:335-346), which is not installed") from e
        info = sources[0]["image_info"]
        for poly in sources[0]["segmentation"]:
            m = np.sum(cocomask.decode(cocomask.frPyObjects(poly, info["height"], info["width"])), axis=2)
            masks.append(fit(m.astype(np.uint8)))
    else:
        info = image_info if image_info is not None else sources[0]["image_info"]
        masks = [fit(m) for m in boxes_to_masks(sources[0]["bbox"], info["height"], info["width"])]
    rows = [mp.preprocess(m[None, ...], return_tensors="pt")["pixel_values"][0] for m in masks]
    return torch.vstack(rows).float()


def process_depth(depth_file, data_args, depth_folder=None) -> torch.Tensor:
    """mm_utils.py:377-418: a (pre-normalised) depth image through the image path - plain ``Image.resize`` to the tower size for
    "resize" (PIL's default filter, like process_image), mean-colour square padding for "pad", then the HF processor."""
    import os

    from PIL import Image

    processor = data_args.image_processor
    if isinstance(depth_file, str):
        depth = Image.open(os.path.join(depth_folder, depth_file) if depth_folder is not None else depth_file)
    else:
        depth = depth_file
    mode = getattr(data_args, "image_aspect_ratio", None)
    if mode == "resize":
        size = _target_size(processor)
        depth = depth.resize((size["height"], size["width"]))
    if mode == "pad":
        depth = _expand2square(depth, tuple(int(x * 255) for x in processor.image_mean))
    return processor.preprocess(depth, return_tensors="pt")["pixel_values"][0]


def boxes_to_masks(bboxes, image_h: int, image_w: int) -> List[np.ndarray]:
    """Box regions as the reference builds them (mm_utils.py:349-364, eval_spatial.py:158-161): clamp,
    then fill [y1:y2, x1:x2] with ones."""
    out = []
    for bbox in bboxes:
        x1, y1, x2, y2 = map(int, bbox)
        x1, x2 = max(0, min(x1, image_w)), max(0, min(x2, image_w))
        y1, y2 = max(0, min(y1, image_h)), max(0, min(y2, image_h))
        m = np.zeros((image_h, image_w), dtype=np.uint8)
        m[y1:y2, x1:x2] = 1
        out.append(m)
    return out


def tokenizer_image_token(prompt: str, tokenizer, image_token_index: int = IMAGE_TOKEN_INDEX, return_tensors: Optional[str] = None,
                          lstrip: bool = False):
    """mm_utils.py:545-570: tokenise the text around every "<image>" and join the pieces with
    ``image_token_index``, keeping a single BOS."""
    chunks = [tokenizer(c).input_ids for c in prompt.split("<image>")]
    ids: List[int] = []
    offset = 0
    if lstrip:
        offset = 1
    elif chunks and chunks[0] and chunks[0][0] == tokenizer.bos_token_id:
        offset = 1
        ids.append(chunks[0][0])
    sep = [image_token_index] * (offset + 1)
    pieces = []
    for c in chunks:
        pieces.extend((c, sep))
    pieces = pieces[:-1]
    for k, piece in enumerate(pieces):
        ids.extend(piece if (k == 0 and lstrip) else piece[offset:])
    if return_tensors is not None:
        if return_tensors == "pt":
            return torch.tensor(ids, dtype=torch.long)
        raise ValueError(f"Unsupported tensor type: {return_tensors}")
    return ids


def get_model_name_from_path(model_path: str) -> str:
    parts = model_path.strip("/").split("/")
    return parts[-2] + "_" + parts[-1] if parts[-1].startswith("checkpoint-") else parts[-1]


class KeywordsStoppingCriteria:
    """mm_utils.py:586-617.  ``output_ids`` is what the generation loop has produced so far ([B, n])."""

    def __init__(self, keywords, tokenizer, input_ids):
        self.keywords = keywords
        self.keyword_ids = []
        self.max_keyword_len = 0
        for kw in keywords:
            ids = tokenizer(kw).input_ids
            if len(ids) > 1 and ids[0] == tokenizer.bos_token_id:
                ids = ids[1:]
            self.max_keyword_len = max(self.max_keyword_len, len(ids))
            self.keyword_ids.append(torch.tensor(ids))
        self.tokenizer = tokenizer
        self.start_len = input_ids.shape[1]

    def call_for_batch(self, output_ids: torch.Tensor, scores=None, **kwargs) -> bool:
        offset = min(output_ids.shape[1] - self.start_len, self.max_keyword_len)
        self.keyword_ids = [k.to(output_ids.device) for k in self.keyword_ids]
        for k in self.keyword_ids:
            if output_ids.shape[1] >= k.shape[0] and bool((output_ids[0, -k.shape[0]:] == k).all()):
                return True
        text = self.tokenizer.batch_decode(output_ids[:, -offset:], skip_special_tokens=True)[0]
        return any(kw in text for kw in self.keywords)

    def __call__(self, output_ids: torch.Tensor, scores=None, **kwargs) -> bool:
        return all(self.call_for_batch(output_ids[i].unsqueeze(0), scores) for i in range(output_ids.shape[0]))
