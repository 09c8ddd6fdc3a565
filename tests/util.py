"""Shared helpers for the GPU parity tests."""
import numpy as np
import torch


def rms(x: torch.Tensor) -> float:
    return float(x.float().pow(2).mean().sqrt())


def err_stats(out: torch.Tensor, ref: torch.Tensor):
    out, ref = out.float().cpu(), ref.float().cpu()
    assert out.shape == ref.shape, (tuple(out.shape), tuple(ref.shape))
    d = (out - ref).abs()
    return float(d.max()), float(d.pow(2).mean().sqrt()), rms(ref)


def assert_close(out: torch.Tensor, ref: torch.Tensor, rel_rms: float, rel_max: float, what: str = ""):
    """Error measured relative to the RMS of the reference (robust for tensors with zeros).

    rel_rms bounds rms(out-ref)/rms(ref); rel_max bounds max|out-ref|/rms(ref)."""
    assert torch.isfinite(out.float()).all(), f"{what}: non-finite values in output"
    mx, er, rr = err_stats(out, ref)
    rr = max(rr, 1e-12)
    assert er / rr <= rel_rms and mx / rr <= rel_max, (
        f"{what}: rms err {er / rr:.3e} (limit {rel_rms:.1e}), max err {mx / rr:.3e} (limit {rel_max:.1e}), ref rms {rr:.3e}")


# tolerances (stated once, used everywhere):
#   one bf16 rounding of an fp32-accumulated result            -> rms 2^-9 ~ 2e-3, max 2^-8 ~ 4e-3 of the value
BF16_1ROUND = dict(rel_rms=4e-3, rel_max=3e-2)
#   a few chained bf16 ops (norm/activation/residual epilogues) -> 1e-2 rms
BF16_CHAIN = dict(rel_rms=1.5e-2, rel_max=2.5e-1)  # max over up to ~2e7 elements: a few-sigma tail of bf16 roundings
#   a whole network stage (tens of layers) vs the fp32 oracle    -> 5e-2 rms
BF16_STAGE = dict(rel_rms=5e-2, rel_max=6e-1)


def load_npz(path):
    z = np.load(path)
    return {k: torch.from_numpy(z[k]) for k in z.files}
