"""The coefficient / index tables behind the GPU preprocessing kernels (spatialrgpt_b200/preprocess.py) against the third-party code
the pinned reference calls: Pillow's bicubic resize (transformers 4.37.2 SiglipImageProcessor -> PIL) and OpenCV's INTER_NEAREST
(llava/mm_utils.py:521-523).  The numpy emulation applies exactly the integer arithmetic of csrc/preprocess.cu."""
import numpy as np
import pytest
from PIL import Image

from spatialrgpt_b200 import preprocess as P


@pytest.mark.parametrize("H,W,oh,ow", [(40, 70, 56, 56), (480, 640, 448, 448), (100, 100, 448, 448), (1000, 750, 336, 336), (448, 448, 448, 448),
                                       (37, 91, 384, 384), (1, 5, 8, 8), (2048, 32, 448, 448)])
def test_bicubic_tables_reproduce_pillow_bit_exactly(H, W, oh, ow):
    rng = np.random.RandomState(H * 31 + W)
    a = rng.randint(0, 256, (H, W, 3), dtype=np.uint8)
    ref = np.asarray(Image.fromarray(a).resize((ow, oh), Image.BICUBIC))
    assert np.array_equal(P.resample_reference_numpy(a, oh, ow), ref)
    kk, bounds, ksize = P.resample_coeffs(W, ow)
    assert kk.shape == (ow, ksize) and bounds.shape == (ow, 2) and int(bounds[:, 1].max()) <= ksize
    assert (bounds[:, 0] >= 0).all() and (bounds[:, 0] + bounds[:, 1] <= W).all()
    # the fixed-point windows sum to ~1.0 (2^22) like Pillow's normalised coefficients
    assert np.abs(kk.sum(1).astype(np.int64) - (1 << P.PRECISION_BITS)).max() <= ksize


@pytest.mark.parametrize("h,w,R", [(40, 70, 56), (480, 640, 448), (37, 91, 384), (448, 448, 448), (1000, 3, 336)])
def test_nearest_indices_reproduce_opencv(h, w, R):
    import cv2
    rng = np.random.RandomState(h + w)
    m = rng.randint(0, 2, (h, w), dtype=np.uint8)
    ref = cv2.resize(m, (R, R), interpolation=cv2.INTER_NEAREST)
    ys, xs = P.nearest_indices(h, R), P.nearest_indices(w, R)
    assert np.array_equal(m[ys][:, xs], ref)


def test_pinned_processor_pipeline_in_numpy():
    """rescale in float64 -> float32, normalise in float32, channels first: the 4.37.2 slow-processor arithmetic the GPU path follows;
    the installed transformers (torchvision backend) may differ from it by one 8-bit resize step."""
    from transformers import SiglipImageProcessor
    rng = np.random.RandomState(0)
    a = rng.randint(0, 256, (40, 70, 3), dtype=np.uint8)
    proc = SiglipImageProcessor(size={"height": 56, "width": 56})
    pinned = P.resample_reference_numpy(a, 56, 56).astype(np.float64) * proc.rescale_factor
    pinned = ((pinned.astype(np.float32) - np.float32(0.5)) / np.float32(0.5)).transpose(2, 0, 1)
    got = proc.preprocess(Image.fromarray(a), return_tensors="pt")["pixel_values"][0].numpy()
    assert np.abs(got - pinned).max() <= 2.0 / 255 + 1e-6  # <= one 8-bit step (x2 from the 1/0.5 normalisation)
