"""Op-level parity of every C-ABI kernel against plain fp32 torch / the CPU oracle.  Needs a B200."""
import math

import numpy as np
import os

import pytest
import torch
import torch.nn.functional as F

from oracle import srgpt_oracle as O
from tests.util import BF16_1ROUND, BF16_CHAIN, assert_close, load_npz

pytestmark = pytest.mark.gpu

DEV = "cuda"
BF = torch.bfloat16


@pytest.fixture(scope="module")
def ops():
    from spatialrgpt_b200 import _lib, ops as _ops
    _lib.load()
    sm, maj, mnr = _lib.device_info()
    assert maj == 10, f"these tests need an sm_100 device, got sm_{maj}{mnr}"
    return _ops


def rnd(*shape, seed=0, scale=1.0, dtype=BF):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(dtype)


# ------------------------------------------------------------------------------------------ GEMM
GEMM_SHAPES = [
    (128, 128, 64), (128, 128, 128), (256, 256, 512), (259, 384, 512), (1, 128, 64), (3, 8, 8), (77, 200, 72),
    (300, 4304, 1152), (300, 1152, 4304), (1024, 1152, 592), (259, 6144, 4096), (130, 272, 4096),
    (4096, 4608, 320), (2100, 9000, 136),  # >= 2 waves of 128x256 tiles -> the BN=256 configuration (N tail: 9000 % 256 = 40)
]


@pytest.mark.parametrize("M,N,K", GEMM_SHAPES)
def test_gemm_plain(ops, M, N, K):
    a, w = rnd(M, K, seed=1), rnd(N, K, seed=2, scale=K ** -0.5)
    ref = a.float() @ w.float().t()
    out = ops.gemm(a.to(DEV), w.to(DEV))
    assert_close(out, ref, **BF16_1ROUND, what=f"gemm {M}x{N}x{K}")
    out32 = ops.gemm(a.to(DEV), w.to(DEV), out_fp32=True)
    assert_close(out32, ref, rel_rms=1e-4, rel_max=1e-3, what=f"gemm fp32-out {M}x{N}x{K}")


@pytest.mark.parametrize("M,N,K", [(259, 384, 512), (200, 296, 144), (1024, 1152, 1152), (77, 304, 72), (4100, 4624, 192)])
def test_gemm_epilogues(ops, M, N, K):
    a, w = rnd(M, K, seed=1), rnd(N, K, seed=2, scale=K ** -0.5)
    bias, res = rnd(N, seed=3), rnd(M, N, seed=4)
    acc = a.float() @ w.float().t()
    ad, wd, bd, rd = a.to(DEV), w.to(DEV), bias.to(DEV), res.to(DEV)
    assert_close(ops.gemm(ad, wd, bias=bd, epilogue=ops.EPI_BIAS), acc + bias.float(), **BF16_1ROUND, what="bias")
    assert_close(ops.gemm(ad, wd, bias=bd, epilogue=ops.EPI_BIAS_GELU_TANH), F.gelu(acc + bias.float(), approximate="tanh"),
                 **BF16_CHAIN, what="gelu_tanh")
    assert_close(ops.gemm(ad, wd, bias=bd, epilogue=ops.EPI_BIAS_GELU_ERF), F.gelu(acc + bias.float()), **BF16_CHAIN, what="gelu_erf")
    assert_close(ops.gemm(ad, wd, bias=bd, residual=rd, epilogue=ops.EPI_BIAS_RESIDUAL), acc + bias.float() + res.float(),
                 **BF16_CHAIN, what="bias_residual")
    assert_close(ops.gemm(ad, wd, residual=rd, epilogue=ops.EPI_BIAS_RESIDUAL), acc + res.float(), **BF16_CHAIN, what="residual")
    # in-place residual (out aliases residual), as the decoder layers use it
    x = rd.clone()
    ops.gemm(ad, wd, residual=x, epilogue=ops.EPI_BIAS_RESIDUAL, out=x)
    assert_close(x, acc + res.float(), **BF16_CHAIN, what="residual in place")
    # broadcast residual rows (position embeddings): row % mod
    mod = 37
    pos = rnd(mod, N, seed=5)
    ref = acc + bias.float() + pos.float()[torch.arange(M) % mod]
    assert_close(ops.gemm(ad, wd, bias=bd, residual=pos.to(DEV), epilogue=ops.EPI_BIAS_RESIDUAL, res_row_mod=mod), ref,
                 **BF16_CHAIN, what="pos-emb residual")
    # SwiGLU over interleaved (gate, up) rows (output width N/2 must keep 16-byte rows)
    if (N // 2) % 8 == 0:
        g, u = acc[:, 0::2], acc[:, 1::2]
        assert_close(ops.gemm(ad, wd, epilogue=ops.EPI_SWIGLU), F.silu(g) * u, **BF16_CHAIN, what="swiglu")
    else:
        from spatialrgpt_b200 import SrgptError
        with pytest.raises(SrgptError):
            ops.gemm(ad, wd, epilogue=ops.EPI_SWIGLU)


def test_gemm_cluster_multicast_path():
    """The opt-in 2-CTA cluster configuration (TMA-multicast weight tile, multicast tcgen05.commit) stays correct.
    The choice is read once per process, so it runs in a child process with SRGPT_GEMM_CL=2."""
    import subprocess
    import sys
    code = (
        "import torch\n"
        "from spatialrgpt_b200 import ops\n"
        "g = torch.Generator().manual_seed(0)\n"
        "for (M, N, K) in [(259, 640, 512), (1024, 1152, 1152), (4096, 4608, 320), (300, 4304, 1152), (32, 4096, 4096), (100, 1536, 4096), (32, 6144, 4096), (32, 28672, 4096)]:\n"
        "    a = torch.randn(M, K, generator=g).bfloat16(); w = (torch.randn(N, K, generator=g) * K ** -0.5).bfloat16()\n"
        "    out = ops.gemm(a.cuda(), w.cuda(), out_fp32=True).cpu()\n"
        "    ref = a.float() @ w.float().t()\n"
        "    err = (out - ref).abs().max().item() / ref.pow(2).mean().sqrt().item()\n"
        "    assert err < 1e-3, (M, N, K, err)\n"
        "print('CLUSTER_OK')\n")
    # 2-CTA shared weight tile; 4-CTA shared activation tile (M <= 384); grouped tile rasterisation (3 m-units per group)
    # ... and the CTA-pair kernel (tcgen05 cta_group::2, one 256 x 256 tile per 2-CTA cluster) forced on for every shape
    # ... the tall stream-K kernel forced on for every M <= 384 whatever the weight size / off everywhere, and the direct
    # (non-TMA-store) epilogue
    for knob, val in (("SRGPT_GEMM_CL", "2"), ("SRGPT_GEMM_TALL", "1"), ("SRGPT_GEMM_GM", "3"), ("SRGPT_GEMM_PAIR", "1"),
                      ("SRGPT_GEMM_EW", "16"), ("SRGPT_GEMM_TSK", "1"), ("SRGPT_GEMM_TSK", "-1"), ("SRGPT_GEMM_DIRECT_EPI", "1"),
                      ("SRGPT_GEMM_TSK_WHOLE", "1")):  # whole narrow weight tiles at M <= 128 instead of the stream-K split
        env = dict(os.environ, **{knob: val})
        r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300,
                           cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        assert r.returncode == 0 and "CLUSTER_OK" in r.stdout, knob + ": " + r.stdout + r.stderr


@pytest.mark.parametrize("M,N,K,epi", [(259, 6144, 4096, "none"), (259, 4096, 14336, "residual"), (259, 28672, 4096, "swiglu"),
                                       (300, 4608, 1152, "gelu_erf"), (100, 20008, 1024, "bias"), (384, 1000, 4096, "none"),
                                       (130, 520, 8200, "residual"), (3, 128259, 4096, "none"),
                                       # one-tile-high problems of a batched decode step (stream-K split; whole narrow tiles under SRGPT_GEMM_TSK_WHOLE=1)
                                       (32, 4096, 4096, "residual"), (32, 6144, 4096, "none"), (32, 28672, 4096, "swiglu"),
                                       (32, 4096, 14336, "residual"), (17, 1000, 4096, "bias"), (64, 4096, 4096, "gelu_erf"),
                                       (128, 4104, 2048, "residual"), (5, 1003, 512, "none")])
def test_gemm_tall_stream_k(ops, M, N, K, epi):
    """Short prompts (M <= 384): all rows in one CTA, (n-tile, k-block) units balanced over the SMs, split tiles combined through
    fp32 partials (gemm_tall_sk_kernel).  Shapes above 4 MB of weights take this path by default; every epilogue; ragged M / N / K;
    repeated launches (epoch flags) and bit-identical results across launches (deterministic partial order)."""
    a, w = rnd(M, K, seed=21), rnd(N, K, seed=22, scale=K ** -0.5)
    acc = a.float() @ w.float().t()
    ad, wd = a.to(DEV), w.to(DEV)
    bias, res = rnd(N, seed=23), rnd(M, N, seed=24)
    if epi == "none":
        ldc = (N + 7) // 8 * 8
        run = lambda: ops.gemm(ad, wd, out=torch.empty(M, ldc, dtype=BF, device=DEV)[:, :N])  # noqa: E731
        ref, tol = acc, BF16_1ROUND
    elif epi == "bias":
        run = lambda: ops.gemm(ad, wd, bias=bias.to(DEV), epilogue=ops.EPI_BIAS)  # noqa: E731
        ref, tol = acc + bias.float(), BF16_1ROUND
    elif epi == "gelu_erf":
        run = lambda: ops.gemm(ad, wd, bias=bias.to(DEV), epilogue=ops.EPI_BIAS_GELU_ERF)  # noqa: E731
        ref, tol = F.gelu(acc + bias.float()), BF16_CHAIN
    elif epi == "residual":
        def run():
            x = res.to(DEV)
            return ops.gemm(ad, wd, residual=x, epilogue=ops.EPI_BIAS_RESIDUAL, out=x)  # in place, as the decoder layers use it
        ref, tol = acc + res.float(), BF16_CHAIN
    else:
        run = lambda: ops.gemm(ad, wd, epilogue=ops.EPI_SWIGLU)  # noqa: E731
        ref, tol = F.silu(acc[:, 0::2]) * acc[:, 1::2], BF16_CHAIN
    out1 = run()
    assert_close(out1, ref, **tol, what=f"tall stream-K {epi}")
    for _ in range(3):
        assert torch.equal(run(), out1)


def test_gemm_grouped_rasterisation_large_activation(ops):
    """An activation larger than the L2 budget (> 80 MB) switches the tile order to L2-sized row groups; every output tile is
    still produced exactly once (checked on sampled rows against fp32)."""
    M, N, K = 24000, 520, 2048  # A = 98 MB
    g = torch.Generator().manual_seed(5)
    a = torch.randn(M, K, generator=g).to(BF)
    w = (torch.randn(N, K, generator=g) * K ** -0.5).to(BF)
    out = ops.gemm(a.to(DEV), w.to(DEV), out=torch.full((M, N), float("nan"), dtype=BF, device=DEV))
    rows = torch.cat([torch.arange(0, 300), torch.randint(0, M, (600,), generator=g), torch.arange(M - 300, M)])
    ref = a[rows].float() @ w.float().t()
    assert_close(out[rows.to(DEV)], ref, **BF16_1ROUND, what="grouped rasterisation")
    assert bool(torch.isfinite(out.float()).all())  # no tile left unwritten (the buffer started as NaN)


def test_gemm_strided_views(ops):
    """A and the output may be column slices of wider buffers (fused qkv)."""
    M, K, N = 200, 144, 136
    big = rnd(M, 3 * K, seed=7).to(DEV)
    w = rnd(N, K, seed=8, scale=K ** -0.5)
    a = big[:, K:2 * K]
    outbuf = torch.zeros(M, 2 * N, dtype=BF, device=DEV)
    ops.gemm(a, w.to(DEV), out=outbuf[:, N:])
    assert_close(outbuf[:, N:], a.float().cpu() @ w.float().t(), **BF16_1ROUND, what="strided")
    assert float(outbuf[:, :N].abs().max()) == 0.0


def test_gemm_rejects_bad_args(ops):
    from spatialrgpt_b200 import SrgptError
    a, w = rnd(16, 12).to(DEV), rnd(8, 12).to(DEV)  # K=12 -> 24-byte rows, not TMA-able
    with pytest.raises(SrgptError):
        ops.gemm(a, w)
    with pytest.raises(SrgptError):
        ops.gemm(rnd(16, 16), rnd(8, 16))  # CPU tensors: no fallback


# ------------------------------------------------------------------------------------------ row ops
@pytest.mark.parametrize("rows,cols", [(7, 144), (300, 1152), (50, 4608), (3, 8)])
def test_layernorm(ops, rows, cols):
    x, w, b = rnd(rows, cols, seed=1, scale=2.0), 1 + 0.1 * rnd(cols, seed=2).float(), 0.1 * rnd(cols, seed=3).float()
    w, b = w.to(BF), b.to(BF)
    ref = F.layer_norm(x.float(), (cols,), w.float(), b.float(), 1e-6)
    assert_close(ops.layernorm(x.to(DEV), w.to(DEV), b.to(DEV), 1e-6), ref, **BF16_1ROUND, what="layernorm")
    assert_close(ops.layernorm(x.to(DEV), w.to(DEV), b.to(DEV), 1e-6, act=1), F.gelu(ref), **BF16_CHAIN, what="layernorm+gelu")


@pytest.mark.parametrize("rows,cols", [(5, 256), (259, 4096), (2, 2560)])
def test_rmsnorm(ops, rows, cols):
    x, w = rnd(rows, cols, seed=1, scale=3.0), (1 + 0.1 * rnd(cols, seed=2).float()).to(BF)
    ref = O.rms_norm(x.float(), w.float(), 1e-5)
    assert_close(ops.rmsnorm(x.to(DEV), w.to(DEV), 1e-5), ref, **BF16_CHAIN, what="rmsnorm")
    # bit-exact against the rounding-faithful bf16 restatement
    ref16 = O.rms_norm(x, w, 1e-5)
    got = ops.rmsnorm(x.to(DEV), w.to(DEV), 1e-5).cpu()
    mism = (got.float() != ref16.float()).float().mean().item()
    assert mism < 2e-3, f"{mism:.2e} of elements differ from the bf16 restatement"  # rsqrt ulp differences only


@pytest.mark.parametrize("side,C", [(27, 144), (27, 1152), (4, 8), (5, 16)])
def test_downsample_layernorm(ops, side, C):
    x = rnd(2, side * side, C, seed=4)
    w, b = (1 + 0.1 * rnd(4 * C, seed=5).float()).to(BF), (0.1 * rnd(4 * C, seed=6).float()).to(BF)
    ds = O.downsample_block(x.float())
    ref = F.layer_norm(ds, (4 * C,), w.float(), b.float(), 1e-5)
    out = ops.downsample_layernorm(x.to(DEV), w.to(DEV), b.to(DEV), 1e-5)
    assert out.shape == ref.shape
    assert_close(out, ref, **BF16_1ROUND, what="downsample+LN")


@pytest.mark.parametrize("R,dtype", [(56, torch.float32), (112, BF), (448, torch.float32)])
def test_patchify_matches_conv(ops, R, dtype):
    n, ps, D = 2, 14, 32
    img = rnd(n, 3, R, R, seed=9, dtype=dtype)
    wconv = rnd(D, 3, ps, ps, seed=10, scale=0.05)
    ldk = 592
    A = ops.patchify(img.to(DEV), ps, ldk).cpu()
    assert A.shape == (n * (R // ps) ** 2, ldk)
    assert float(A[:, 588:].abs().max()) == 0.0
    ref = F.conv2d(img.to(BF).float(), wconv.float(), stride=ps).flatten(2).transpose(1, 2).reshape(-1, D)
    got = A[:, :588].float() @ wconv.float().reshape(D, -1).t()
    assert_close(got, ref, rel_rms=1e-5, rel_max=1e-4, what="patchify")


def test_splice_rows(ops):
    H = 64
    srcs = [rnd(10 + i, H, seed=20 + i).to(DEV) for i in range(4)]
    g = torch.Generator().manual_seed(1)
    sid = torch.randint(0, 4, (50,), generator=g).to(torch.int32)
    srow = torch.tensor([int(torch.randint(0, 10 + int(s), (1,), generator=g)) for s in sid], dtype=torch.int32)
    out = ops.splice_rows(*srcs, sid.to(DEV), srow.to(DEV)).cpu()
    ref = torch.stack([srcs[int(s)][int(r)].cpu() for s, r in zip(sid, srow)])
    assert torch.equal(out, ref)


def test_argmax_first_index_on_ties(ops):
    x = torch.zeros(3, 1000)
    x[0, 17] = x[0, 900] = 5.0
    x[1, 999] = 1.0
    x[2] = -1.0
    assert ops.argmax_f32(x.to(DEV)).tolist() == [17, 999, 0]


# ------------------------------------------------------------------------------------------ region kernels
@pytest.mark.parametrize("tag", ["rgb448", "depth448", "rgb384", "odd336"])
def test_mask_pooling_golden(ops, golden_dir, tag):
    """Reference MaskPooling known answers (tests/golden/op_kats.npz), row-major features."""
    g = load_npz(os.path.join(golden_dir, "op_kats.npz"))
    x, masks = g[f"{tag}_x"].to(BF), g[f"{tag}_masks"].float()
    side = int(round(x.shape[1] ** 0.5))
    w = ops.mask_weights(masks[None].to(DEV), side, ops.ORDER_ROWMAJOR)
    # the resampled + normalised weights are bit-exact vs torch (same fp32 taps, same bf16 roundings)
    m = F.interpolate(masks[None], scale_factor=(x.shape[1] / masks.shape[-1] ** 2) ** 0.5, mode="bilinear")[0].to(BF)
    wref = (m.flatten(1) / (m.sum(dim=(-1, -2)) + 1e-8).unsqueeze(-1))
    frac = (w[0].cpu().float() != wref.float()).float().mean().item()
    assert frac < 1e-3, f"{frac:.2e} of the mask weights differ from torch"
    out = ops.mask_pool(x.to(DEV), w)[0]
    assert_close(out, g[f"{tag}_out_f32"], **BF16_CHAIN, what=f"mask_pool {tag} vs fp32 reference")
    assert_close(out, g[f"{tag}_out_bf16"], **BF16_CHAIN, what=f"mask_pool {tag} vs bf16 reference")


def _nested_perm(side):
    P = side // 4
    idx = torch.empty(side * side, dtype=torch.long)
    for y in range(side):
        for x in range(side):
            idx[y * side + x] = (((y >> 2) * P + (x >> 2)) << 4) | ((((y >> 1) & 1) * 2 + ((x >> 1) & 1)) << 2) | ((y & 1) * 2 + (x & 1))
    return idx  # idx[rowmajor] = nested row


@pytest.mark.parametrize("side,C,M", [(16, 144, 3), (128, 1152, 8), (128, 1152, 11), (32, 256, 16)])
def test_mask_pool_nested_order_and_multi_pass(ops, side, C, M):
    R = side * 7 // 2
    x = rnd(2, side * side, C, seed=3)
    masks = (torch.rand(2, M, R, R, generator=torch.Generator().manual_seed(4)) > 0.7).float()
    masks[:, 0] = 1.0  # all-ones mask -> plain mean
    ref = torch.stack(O.mask_pooling(x.float(), [masks[0], masks[1]]))
    perm = _nested_perm(side)
    xn = torch.empty_like(x)
    xn[:, perm] = x
    w = ops.mask_weights(masks.to(DEV), side, ops.ORDER_NESTED)
    out = ops.mask_pool(xn.to(DEV), w)
    assert_close(out, ref, **BF16_CHAIN, what="nested mask_pool")
    assert_close(out[:, 0], x.float().mean(1), **BF16_CHAIN, what="all-ones mask == mean")
    # reorder kernel round trip
    back = ops.reorder_rows(xn.to(DEV), side, ops.ORDER_NESTED, ops.ORDER_ROWMAJOR).cpu()
    assert torch.equal(back, x)
    # bf16 masks are accepted as well
    w2 = ops.mask_weights(masks.to(BF).to(DEV), side, ops.ORDER_NESTED)
    assert torch.equal(w2, w)


@pytest.mark.parametrize("side,C", [(16, 144), (128, 64), (96, 32), (108, 16)])
def test_adaptive_avgpool(ops, side, C):
    x = rnd(2, side * side, C, seed=5)
    ref = F.adaptive_avg_pool2d(x.float().view(2, side, side, C).permute(0, 3, 1, 2), 27).flatten(2).transpose(1, 2)
    out = ops.adaptive_avgpool(x.to(DEV), side, 27, ops.ORDER_ROWMAJOR)
    assert_close(out, ref, **BF16_1ROUND, what="adaptive_avgpool row-major")
    if side % 4 == 0:
        xn = torch.empty_like(x)
        xn[:, _nested_perm(side)] = x
        out2 = ops.adaptive_avgpool(xn.to(DEV), side, 27, ops.ORDER_NESTED)
        assert torch.equal(out2, out)


@pytest.mark.parametrize("h,w,H,W", [(37, 53, 448, 448), (518, 518, 480, 640), (8, 8, 8, 8)])
def test_depth_to_u8x3(ops, h, w, H, W):
    d = torch.rand(1, h, w, generator=torch.Generator().manual_seed(6)) * 10 + 1
    ref = O.depth_to_u8x3(d, H, W)
    out = ops.depth_to_u8x3(d.to(DEV), H, W).cpu()
    assert out.shape == ref.shape and out.dtype == torch.uint8
    assert torch.equal(out, ref), "byte work is bit-exact: the kernel mirrors ATen's fp32 operation order (region.cu bilinear_tap_aten)"
    # the reference interpolates on the GPU (eval_spatial.py:96-100) and normalises in numpy: the same bytes
    dev_interp = torch.nn.functional.interpolate(d.to(DEV)[None], (H, W), mode="bilinear", align_corners=False)[0, 0].cpu().numpy()
    n = (dev_interp - dev_interp.min()) / (dev_interp.max() - dev_interp.min()) * 255.0
    assert torch.equal(out[..., 0], torch.from_numpy(n.astype("uint8")))
    assert int(out.min()) == 0 and int(out.max()) == 255


# ------------------------------------------------------------------------------------------ attention
def _sdpa_ref(q, k, v, nh, nkv, hd, scale, causal):
    B, S = q.shape[0], q.shape[1]
    qh = q.float().view(B, S, nh, hd).transpose(1, 2)
    kh = k.float().view(B, S, nkv, hd).transpose(1, 2).repeat_interleave(nh // nkv, 1)
    vh = v.float().view(B, S, nkv, hd).transpose(1, 2).repeat_interleave(nh // nkv, 1)
    att = qh @ kh.transpose(-1, -2) * scale
    if causal:
        att = att.masked_fill(~torch.ones(S, S, dtype=torch.bool).tril(), float("-inf"))
    return (att.softmax(-1) @ vh).transpose(1, 2).reshape(B, S, nh * hd)


@pytest.mark.parametrize("B,S,nh,nkv,hd,causal", [
    (2, 16, 2, 2, 72, False), (2, 64, 2, 2, 72, False), (1, 1024, 4, 4, 72, False), (3, 100, 2, 2, 72, False),
    (1, 259, 8, 2, 128, True), (2, 65, 4, 1, 128, True), (1, 1, 2, 2, 128, True), (1, 300, 4, 4, 128, True),
    (1, 130, 2, 2, 128, False), (1, 97, 2, 1, 64, True),
])
def test_attention_prefill(ops, B, S, nh, nkv, hd, causal):
    qkv = rnd(B * S, (nh + 2 * nkv) * hd, seed=11)
    qd, kd = nh * hd, nkv * hd
    q, k, v = qkv[:, :qd], qkv[:, qd:qd + kd], qkv[:, qd + kd:]
    scale = hd ** -0.5
    ref = _sdpa_ref(q.reshape(B, S, -1), k.reshape(B, S, -1), v.reshape(B, S, -1), nh, nkv, hd, scale, causal).reshape(B * S, -1)
    d = qkv.to(DEV)
    out = ops.attention_prefill(d[:, :qd], d[:, qd:qd + kd], d[:, qd + kd:], B, S, nh, nkv, hd, scale, causal)
    assert_close(out, ref, rel_rms=1e-2, rel_max=8e-2, what=f"attention B{B} S{S} hd{hd} causal={causal}")


@pytest.mark.parametrize("lens,nh,nkv,hd,causal", [
    ([259, 1, 64, 130], 8, 2, 128, True), ([5, 300], 4, 4, 128, True), ([100, 7, 1024], 2, 2, 72, False), ([64], 2, 1, 64, True),
])
def test_attention_prefill_varlen(ops, lens, nh, nkv, hd, causal):
    """Packed variable-length sequences (modeling_llama.py:540-562): every sequence equals its own dense attention."""
    S = sum(lens)
    qkv = rnd(S, (nh + 2 * nkv) * hd, seed=21)
    qd, kd = nh * hd, nkv * hd
    scale = hd ** -0.5
    cu = torch.tensor([0] + torch.tensor(lens).cumsum(0).tolist(), dtype=torch.int32)
    d = qkv.to(DEV)
    out = ops.attention_prefill_varlen(d[:, :qd], d[:, qd:qd + kd], d[:, qd + kd:], cu.to(DEV), max(lens), nh, nkv, hd, scale, causal)
    o = 0
    for n in lens:
        blk = qkv[o:o + n]
        ref = _sdpa_ref(blk[None, :, :qd], blk[None, :, qd:qd + kd], blk[None, :, qd + kd:], nh, nkv, hd, scale, causal)[0]
        assert_close(out[o:o + n], ref, rel_rms=1e-2, rel_max=8e-2, what=f"varlen attention len {n}")
        # and the single-sequence entry point on the same rows: bit-identical when both run the same kernel; the dense head-72 path
        # may take the ping-pong kernel (different summation order) -> one bf16 rounding of difference at most
        one = ops.attention_prefill(d[o:o + n, :qd], d[o:o + n, qd:qd + kd], d[o:o + n, qd + kd:], 1, n, nh, nkv, hd, scale, causal)
        if hd == 72 and not causal:
            assert_close(one, out[o:o + n], rel_rms=4e-3, rel_max=5e-2, what="dense vs varlen entry")
        else:
            assert torch.equal(one, out[o:o + n])
        o += n


def test_rope_kv_append_varlen_equals_per_sequence(ops):
    hd, page, nh, nkv = 128, 16, 4, 2
    lens = [37, 1, 259, 16]
    cos, sin = _rope_tables(hd, 500000.0, 512)
    qkv = rnd(sum(lens), (nh + 2 * nkv) * hd, seed=22).to(DEV)
    cap = 20
    n_pages = len(lens) * cap
    perm = torch.randperm(n_pages, generator=torch.Generator().manual_seed(5)).to(torch.int32).view(len(lens), cap).to(DEV)
    starts = torch.tensor([0, 3, 0, 32], dtype=torch.int32, device=DEV)  # non-zero start positions too
    cu = torch.tensor([0] + torch.tensor(lens).cumsum(0).tolist(), dtype=torch.int32, device=DEV)
    pages_a = torch.zeros(n_pages, 2, page, nkv, hd, dtype=BF, device=DEV)
    pages_b = torch.zeros_like(pages_a)
    a = qkv.clone()
    ops.rope_kv_append_varlen(a, nh, nkv, hd, cos, sin, starts, pages_a, perm, page, cu)
    b = qkv.clone()
    o = 0
    for i, n in enumerate(lens):
        blk = b[o:o + n]
        ops.rope_kv_append(blk, nh, nkv, hd, cos, sin, starts[i:i + 1], pages_b, perm[i], page)
        o += n
    assert torch.equal(a, b) and torch.equal(pages_a, pages_b)
    assert float(pages_a.float().abs().sum()) > 0


def test_argmax_bf16_first_index_on_ties(ops):
    x = torch.zeros(3, 1000, dtype=BF)
    x[0, 17] = x[0, 900] = 5.0
    x[1, 999] = 1.0
    x[2] = -1.0
    wide = torch.zeros(3, 1024, dtype=BF)
    wide[:, :1000] = x
    wide[:, 1000:] = 9.0  # beyond `cols`: must be ignored (strided rows)
    assert ops.argmax_bf16(wide.to(DEV)[:, :1000]).tolist() == [17, 999, 0]
    # vocabulary-wide rows take the segmented kernel (one CTA per 4096 columns, 64-bit atomicMax of (value, ~index)): same answers,
    # ties across segments -> lowest index, negative values, -inf, a NaN that must not win, an all-NaN row -> 0
    V = 128259
    g = torch.Generator().manual_seed(3)
    y = (torch.randn(5, V + 5, generator=g) * 3).to(BF)
    y[1, 70000] = y[1, 120] = 50.0
    y[2] = -torch.rand(V + 5, generator=g).to(BF) - 1
    y[3, 5] = float("nan")
    y[4] = float("nan")
    y[:, V:] = 99.0
    ref = y[:, :V].float().nan_to_num(nan=float("-inf")).argmax(-1).tolist()
    ref[4] = 0
    got = ops.argmax_bf16(y.to(DEV)[:, :V]).tolist()
    assert got == ref and got[1] == 120


def _rope_tables(hd, theta, max_pos):
    from spatialrgpt_b200.config import LlamaDims
    from spatialrgpt_b200.llama_decoder import build_rope_tables
    return build_rope_tables(LlamaDims(head_dim=hd, rope_theta=theta), max_pos, DEV)


@pytest.mark.parametrize("S,nh,nkv,theta", [(37, 4, 2, 10000.0), (259, 8, 2, 500000.0), (16, 2, 2, 10000.0), (700, 4, 1, 500000.0), (513, 2, 2, 10000.0)])
def test_rope_kv_append_and_decode_attention(ops, S, nh, nkv, theta):
    hd, page = 128, 16
    cfg = O.OracleConfig(head_dim=hd, rope_theta=theta)
    qkv = rnd(S, (nh + 2 * nkv) * hd, seed=12)
    cos, sin = _rope_tables(hd, theta, 1024)
    cr, sr = O.rope_cos_sin(cfg, torch.arange(1024), BF)
    assert torch.equal(cos.cpu(), cr[:, : hd // 2]) and torch.equal(sin.cpu(), sr[:, : hd // 2])
    n_pages = (S + 1 + page - 1) // page + 2
    pages = torch.zeros(n_pages, 2, page, nkv, hd, dtype=BF, device=DEV)
    perm = torch.randperm(n_pages, generator=torch.Generator().manual_seed(3)).to(torch.int32)  # scattered physical pages
    pt = perm.to(DEV)
    d = qkv.to(DEV).clone()
    ops.rope_kv_append(d, nh, nkv, hd, cos, sin, torch.zeros(1, dtype=torch.int32, device=DEV), pages, pt, page)
    # reference: bf16 rope exactly as modeling_llama.py:186-191
    q = qkv[:, : nh * hd].view(S, nh, hd).transpose(0, 1)
    k = qkv[:, nh * hd:(nh + nkv) * hd].view(S, nkv, hd).transpose(0, 1)
    v = qkv[:, (nh + nkv) * hd:].view(S, nkv, hd)
    c, s = cr[:S], sr[:S]
    qr = (q * c[None]) + (O.rotate_half(q) * s[None])
    kr = (k * c[None]) + (O.rotate_half(k) * s[None])
    got = d.cpu()
    assert torch.equal(got[:, : nh * hd].view(S, nh, hd).transpose(0, 1), qr)
    assert torch.equal(got[:, nh * hd:(nh + nkv) * hd].view(S, nkv, hd).transpose(0, 1), kr)
    pc = pages.cpu()
    for pos in (0, S // 2, S - 1):
        pg, sl = int(perm[pos // page]), pos % page
        assert torch.equal(pc[pg, 0, sl], kr[:, pos]) and torch.equal(pc[pg, 1, sl], v[pos])
    # decode attention for a new query at position S-1 over the S cached tokens
    qn = rnd(nh * hd, seed=13)
    out = torch.empty(nh * hd, dtype=BF, device=DEV)
    ops.attention_decode(qn.to(DEV), out, pages, pt, page, torch.tensor([S - 1], dtype=torch.int32, device=DEV), nh, nkv, hd, hd ** -0.5)
    kk = kr.float().repeat_interleave(nh // nkv, 0)
    vv = v.float().transpose(0, 1).repeat_interleave(nh // nkv, 0)
    att = torch.einsum("hd,hsd->hs", qn.float().view(nh, hd), kk) * hd ** -0.5
    ref = torch.einsum("hs,hsd->hd", att.softmax(-1), vv).reshape(-1)
    assert_close(out, ref, rel_rms=6e-3, rel_max=5e-2, what="decode attention")


# ------------------------------------------------------------------------------------------ GEMV (decode)
@pytest.mark.parametrize("N,K", [(512, 256), (4096, 4096), (6144, 4096), (4096, 14336), (1000, 1032), (2, 8)])
def test_gemv_plain_and_residual(ops, N, K):
    x, w, r = rnd(K, seed=1), rnd(N, K, seed=2, scale=K ** -0.5), rnd(N, seed=3)
    ref = w.float() @ x.float()
    y = torch.empty(N, dtype=BF, device=DEV)
    ops.gemv(x.to(DEV), w.to(DEV), y)
    assert_close(y, ref, **BF16_1ROUND, what="gemv")
    h = r.to(DEV).clone()
    ops.gemv(x.to(DEV), w.to(DEV), h, residual=h)
    assert_close(h, ref + r.float(), **BF16_CHAIN, what="gemv + residual in place")


@pytest.mark.parametrize("I,K", [(384, 256), (14336, 4096), (100, 64)])
def test_gemv_rmsnorm_swiglu(ops, I, K):
    x, nw = rnd(K, seed=1, scale=2.0), (1 + 0.1 * rnd(K, seed=4).float()).to(BF)
    gate, up = rnd(I, K, seed=2, scale=K ** -0.5), rnd(I, K, seed=3, scale=K ** -0.5)
    from spatialrgpt_b200.weights import interleave_rows
    w = interleave_rows(gate, up)
    xn = O.rms_norm(x.float(), nw.float(), 1e-5)
    ref = F.silu(gate.float() @ xn) * (up.float() @ xn)
    y = torch.empty(I, dtype=BF, device=DEV)
    ops.gemv(x.to(DEV), w.to(DEV), y, norm_weight=nw.to(DEV), eps=1e-5, mode=ops.GEMV_SWIGLU)
    assert_close(y, ref, **BF16_CHAIN, what="gemv rmsnorm+swiglu")


@pytest.mark.parametrize("nh,nkv,K,pos", [(2, 1, 256, 5), (32, 8, 4096, 300), (4, 4, 512, 0)])
def test_gemv_qkv_rope_cache(ops, nh, nkv, K, pos):
    hd, page = 128, 16
    N = (nh + 2 * nkv) * hd
    x, nw, w = rnd(K, seed=1, scale=2.0), (1 + 0.1 * rnd(K, seed=4).float()).to(BF), rnd(N, K, seed=2, scale=K ** -0.5)
    cos, sin = _rope_tables(hd, 10000.0, 512)
    n_pages = pos // page + 2
    pages = torch.zeros(n_pages, 2, page, nkv, hd, dtype=BF, device=DEV)
    pt = torch.arange(n_pages - 1, -1, -1, dtype=torch.int32, device=DEV)
    q = torch.empty(nh * hd, dtype=BF, device=DEV)
    ops.gemv(x.to(DEV), w.to(DEV), q, norm_weight=nw.to(DEV), eps=1e-5, mode=ops.GEMV_QKV_ROPE, n_heads=nh, n_kv_heads=nkv,
             head_dim=hd, cos_tab=cos, sin_tab=sin, pos=torch.tensor([pos], dtype=torch.int32, device=DEV), kv_pages=pages,
             page_table=pt, page_size=page)
    xn = O.rms_norm(x.float(), nw.float(), 1e-5)
    full = (w.float() @ xn)
    c, s = cos[pos].cpu().float().repeat(2), sin[pos].cpu().float().repeat(2)
    qr = full[: nh * hd].view(nh, hd)
    kr = full[nh * hd:(nh + nkv) * hd].view(nkv, hd)
    vr = full[(nh + nkv) * hd:].view(nkv, hd)
    qref = qr * c + O.rotate_half(qr) * s
    kref = kr * c + O.rotate_half(kr) * s
    assert_close(q.view(nh, hd), qref, **BF16_CHAIN, what="q rope")
    pg = int(pt[pos // page])
    assert_close(pages[pg, 0, pos % page], kref, **BF16_CHAIN, what="k cache")
    assert_close(pages[pg, 1, pos % page], vr, **BF16_CHAIN, what="v cache")


@pytest.mark.parametrize("V,K", [(512, 256), (32003, 512), (128259, 4096)])
def test_lm_head_argmax(ops, V, K):
    x, nw, w = rnd(K, seed=1), (1 + 0.1 * rnd(K, seed=4).float()).to(BF), rnd(V, K, seed=2, scale=4 * K ** -0.5)
    emb = rnd(V, K, seed=5)
    xn = O.rms_norm(x.float(), nw.float(), 1e-5)
    ref = (w.float() @ xn)
    ws = ops.lm_head_workspace(V, DEV)
    out_ids = torch.full((8,), -1, dtype=torch.int64, device=DEV)
    step = torch.tensor([3], dtype=torch.int32, device=DEV)
    pos = torch.tensor([41], dtype=torch.int32, device=DEV)
    nxt = torch.zeros(K, dtype=BF, device=DEV)
    logits = torch.empty(V, dtype=torch.float32, device=DEV)
    ops.lm_head_argmax(x.to(DEV), w.to(DEV), nw.to(DEV), 1e-5, ws, out_ids, step, pos, embed_table=emb.to(DEV), next_x=nxt, logits_out=logits)
    assert_close(logits, ref, **BF16_1ROUND, what="logits")
    tok = int(out_ids[3])
    assert tok == int(torch.argmax(logits.cpu())), "argmax must agree with its own logits (first index on ties)"
    assert float(ref[tok]) >= float(ref.max()) - 0.02 * float(ref.std())
    assert int(step) == 4 and int(pos) == 42
    assert torch.equal(nxt.cpu(), emb[tok])
    assert out_ids.tolist()[:3] == [-1, -1, -1]


# ------------------------------------------------------------------------------------------ sampling
def _hf_nucleus(logits, temperature, top_k, top_p):
    """Probabilities after HF's TemperatureLogitsWarper -> TopKLogitsWarper -> TopPLogitsWarper (transformers 4.37.2 semantics)."""
    scores = logits.double() / temperature
    if top_k and top_k < scores.numel():
        kth = scores.topk(top_k).values[-1]
        scores = scores.masked_fill(scores < kth, float("-inf"))
    if top_p < 1.0:
        sorted_logits, sorted_idx = torch.sort(scores, descending=False)
        cum = sorted_logits.softmax(-1).cumsum(-1)
        remove = cum <= (1 - top_p)
        remove[-1:] = False  # min_tokens_to_keep = 1
        scores = scores.masked_fill(torch.zeros_like(remove).scatter(0, sorted_idx, remove), float("-inf"))
    return scores.softmax(-1)


@pytest.mark.parametrize("V,temperature,top_k,top_p", [(1000, 0.7, 50, 0.8), (4099, 1.0, 0, 0.9), (333, 1.3, 20, 1.0), (128259, 0.2, 50, 0.7)])
def test_sample_top_p_matches_hf_distribution(ops, V, temperature, top_k, top_p):
    g = torch.Generator().manual_seed(V)
    logits = (torch.randn(V, generator=g) * 2.5)
    ref = _hf_nucleus(logits, temperature, top_k, top_p)
    support = ref > 0
    n = 3000
    d_logits = logits.to(DEV)
    params = torch.tensor([temperature, top_p, float(top_k)], device=DEV)
    out = torch.zeros(n, dtype=torch.int64, device=DEV)
    step = torch.zeros(1, dtype=torch.int32, device=DEV)
    for i in range(n):
        step.fill_(i + 1)
        ops.sample_top_p(d_logits, params, 1234, step, -1, out)
    ids = out.cpu()
    assert bool(support[ids].all()), "a sampled token lies outside the HF nucleus"
    freq = torch.bincount(ids, minlength=V).double() / n
    # every token's frequency within 5 binomial standard deviations (+ one count) of its HF probability
    sd = (ref * (1 - ref) / n).sqrt()
    assert bool(((freq - ref).abs() <= 5 * sd + 1.5 / n).all()), float(((freq - ref).abs() - 5 * sd).max())
    # same (seed, step) -> same token; top_k = 1 -> the arg max
    step.fill_(7)
    a = torch.zeros(8, dtype=torch.int64, device=DEV); b = torch.zeros(8, dtype=torch.int64, device=DEV)
    ops.sample_top_p(d_logits, params, 99, step, -1, a); ops.sample_top_p(d_logits, params, 99, step, -1, b)
    assert int(a[6]) == int(b[6])
    ops.sample_top_p(d_logits, torch.tensor([temperature, top_p, 1.0], device=DEV), 5, step, -1, a)
    assert int(a[6]) == int(logits.argmax())


# ------------------------------------------------------------------------------------------ tensor-parallel pieces (one GPU)
@pytest.mark.parametrize("world", [2, 4])
def test_tp_shards_reproduce_the_full_decode_ops(ops, world):
    """The per-rank kernels of tensor_parallel.py, all ranks emulated on ONE GPU: the column-parallel QKV GEMV (+RoPE, KV append at the
    rank's kv-head offset) writes exactly the q / K / V the full kernel writes; attention over the rank's heads equals the slice of
    the full attention; row-parallel partial sums add up (fp32) to the full o_proj product; vocabulary-parallel (value, index)
    candidates pick the full arg max."""
    from spatialrgpt_b200.config import LlamaDims
    from spatialrgpt_b200.llama_decoder import build_rope_tables
    nh, nkv, hd, H, V = 8, 4, 128, 512, 1003
    dims = LlamaDims(hidden_size=H, num_attention_heads=nh, num_key_value_heads=nkv, head_dim=hd, intermediate_size=1024, vocab_size=V)
    cos, sin = build_rope_tables(dims, 256, DEV)
    x = rnd(H, seed=1).to(DEV)
    nw = (1 + 0.1 * rnd(H, seed=2)).to(DEV)
    wqkv = rnd((nh + 2 * nkv) * hd, H, seed=3, scale=H ** -0.5).to(DEV)
    n_pages, page = 8, 16
    pt = torch.arange(n_pages, dtype=torch.int32, device=DEV)
    pos = torch.tensor([37], dtype=torch.int32, device=DEV)
    base_pages = rnd(n_pages, 2, page, nkv, hd, seed=4).to(DEV)
    full_pages, tp_pages = base_pages.clone(), base_pages.clone()
    q_full = torch.zeros(nh * hd, dtype=BF, device=DEV)
    ops.gemv(x, wqkv, q_full, norm_weight=nw, eps=1e-5, mode=ops.GEMV_QKV_ROPE, n_heads=nh, n_kv_heads=nkv, head_dim=hd, cos_tab=cos, sin_tab=sin,
             pos=pos, kv_pages=full_pages, page_table=pt, page_size=page)
    attn_full = torch.zeros(nh * hd, dtype=BF, device=DEV)
    ops.attention_decode(q_full, attn_full, full_pages, pt, page, pos, nh, nkv, hd, hd ** -0.5)
    wo = rnd(H, nh * hd, seed=5, scale=(nh * hd) ** -0.5).to(DEV)
    o_ref = attn_full.float() @ wo.float().t()
    nhl, nkvl = nh // world, nkv // world
    partial_sum = torch.zeros(H, dtype=torch.float32, device=DEV)
    for r in range(world):
        w_local = torch.cat([wqkv[r * nhl * hd:(r + 1) * nhl * hd], wqkv[(nh + r * nkvl) * hd:(nh + (r + 1) * nkvl) * hd],
                             wqkv[(nh + nkv + r * nkvl) * hd:(nh + nkv + (r + 1) * nkvl) * hd]], 0).contiguous()
        q_loc = torch.zeros(nhl * hd, dtype=BF, device=DEV)
        ops.gemv_tp_qkv(x, w_local, q_loc, nw, 1e-5, nhl, nkvl, hd, cos, sin, pos, tp_pages, pt, page, nkv, r * nkvl)
        assert torch.equal(q_loc, q_full[r * nhl * hd:(r + 1) * nhl * hd])
        a_loc = torch.zeros(nhl * hd, dtype=BF, device=DEV)
        ops.attention_decode_tp(q_loc, a_loc, tp_pages, pt, page, pos, nhl, nh // nkv, nkv, r * nkvl, hd, hd ** -0.5)
        assert torch.equal(a_loc, attn_full[r * nhl * hd:(r + 1) * nhl * hd])
        part = torch.zeros(H, dtype=torch.float32, device=DEV)
        ops.gemv_tp_partial(a_loc, wo[:, r * nhl * hd:(r + 1) * nhl * hd].contiguous(), part)
        partial_sum += part
    assert torch.equal(tp_pages, full_pages), "the ranks together append exactly the K/V rows of the full kernel"
    assert_close(partial_sum, o_ref, rel_rms=1e-4, rel_max=1e-3, what="row-parallel partial sums")
    h = rnd(H, seed=6).to(DEV)
    h2 = h.clone()
    ops.tp_residual_add(h2, partial_sum)
    assert torch.equal(h2, (partial_sum.to(BF).float() + h.float()).to(BF))
    # vocabulary-parallel arg max (ties -> lowest index, like the full kernel)
    wl = rnd(V, H, seed=7, scale=H ** -0.5).to(DEV)
    wl[700] = wl[123]  # an exact tie across two ranks' blocks
    full_ids = torch.zeros(4, dtype=torch.int64, device=DEV)
    st, ps = torch.zeros(1, dtype=torch.int32, device=DEV), torch.zeros(1, dtype=torch.int32, device=DEV)
    ops.lm_head_argmax(x, wl, nw, 1e-5, ops.lm_head_workspace(V, DEV), full_ids, st, ps)
    per = (V + world - 1) // world
    best_all = torch.zeros(2 * world, dtype=torch.int32, device=DEV)
    for r in range(world):
        v0, v1 = min(V, r * per), min(V, (r + 1) * per)
        ops.lm_head_local_best(x, wl[v0:v1], nw, 1e-5, ops.lm_head_workspace(v1 - v0, DEV), v0, best_all[2 * r:2 * r + 2])
    tp_ids = torch.zeros(4, dtype=torch.int64, device=DEV)
    st2, ps2 = torch.zeros(1, dtype=torch.int32, device=DEV), torch.zeros(1, dtype=torch.int32, device=DEV)
    nxt = torch.zeros(H, dtype=BF, device=DEV)
    emb = rnd(V, H, seed=8).to(DEV)
    ops.tp_pick_token(best_all, world, emb, nxt, tp_ids, st2, ps2)
    assert int(tp_ids[0]) == int(full_ids[0]) and int(st2) == 1 and int(ps2) == 1
    assert torch.equal(nxt, emb[int(full_ids[0])])


@pytest.mark.parametrize("side,C,M,R", [(27, 72, 5, 378), (27, 1152, 17, 384), (24, 200, 1, 336)])
def test_mask_pool_odd_sides_and_ragged_channels(ops, side, C, M, R):
    """L = side^2 not a multiple of 8 (the 27 x 27 tower grid of a 384-px SigLIP: padded weight rows, TMA view of L columns), channel
    counts that are not a multiple of the 128-channel CTA tile (out-of-bounds channels are zero-filled and never stored), more than
    16 masks (two passes over the features)."""
    x = rnd(2, side * side, C, seed=13)
    masks = (torch.rand(2, M, R, R, generator=torch.Generator().manual_seed(14)) > 0.6).float()
    ref = torch.stack(O.mask_pooling(x.float(), [masks[0], masks[1]]))
    w = ops.mask_weights(masks.to(DEV), side, ops.ORDER_ROWMAJOR)
    assert w.shape == (2, M, side * side)
    out = ops.mask_pool(x.to(DEV), w)
    assert_close(out, ref, **BF16_CHAIN, what="odd-side mask_pool")
    assert torch.equal(ops.mask_pool(x.to(DEV), w), out)


@pytest.mark.parametrize("nh,nkv", [(8, 2), (4, 4), (8, 1), (6, 2)])
def test_attention_decode_batched_equals_per_sequence(ops, nh, nkv):
    """Batched decode attention (one CTA per kv head and sequence, the GQA group served from one pass over K / V; group sizes without
    a specialisation fall back to one CTA per query head) against the single-sequence kernel on every sequence: different lengths,
    scattered pages, q read as a column slice of a fused qkv buffer."""
    hd, page, B = 128, 16, 5
    lens = [1, 17, 300, 64, 259]
    n_pages = sum((n + page - 1) // page for n in lens) + 3
    pages = rnd(n_pages, 2, page, nkv, hd, seed=31).to(DEV)
    perm = torch.randperm(n_pages, generator=torch.Generator().manual_seed(32)).tolist()
    cap = max((n + page - 1) // page for n in lens) + 1
    pts = torch.zeros(B, cap, dtype=torch.int32)
    o = 0
    for b, n in enumerate(lens):
        k = (n + page - 1) // page
        pts[b, :k] = torch.tensor(perm[o:o + k], dtype=torch.int32)
        o += k
    pts = pts.to(DEV)
    qkv = rnd(B, (nh + 2 * nkv) * hd, seed=33).to(DEV)
    pos = torch.tensor([n - 1 for n in lens], dtype=torch.int32, device=DEV)
    out = torch.zeros(B, nh * hd, dtype=BF, device=DEV)
    ops.attention_decode_batched(qkv[:, :nh * hd], out, pages, pts, page, pos, nh, nkv, hd, hd ** -0.5)
    for b in range(B):
        one = torch.zeros(nh * hd, dtype=BF, device=DEV)
        ops.attention_decode(qkv[b, :nh * hd].contiguous(), one, pages, pts[b].contiguous(), page, pos[b:b + 1].contiguous(), nh, nkv, hd, hd ** -0.5)
        assert_close(out[b], one, rel_rms=2e-3, rel_max=2e-2, what=f"batched decode attention seq {b}")


# ------------------------------------------------------------------------------------------ preprocessing on the GPU
@pytest.mark.parametrize("H,W,oh,ow", [(40, 70, 56, 56), (480, 640, 448, 448), (100, 100, 448, 448), (1000, 750, 336, 336), (448, 448, 448, 448), (37, 91, 384, 384)])
def test_gpu_bicubic_resize_is_pillow_exact(ops, H, W, oh, ow):
    """f2: Pillow's BICUBIC resize (what the pinned SiglipImageProcessor of transformers 4.37.2 calls) reproduced bit for bit."""
    from PIL import Image

    from spatialrgpt_b200.preprocess import resize_bicubic_u8
    a = np.random.RandomState(H + W).randint(0, 256, (H, W, 3), dtype=np.uint8)
    ref = np.asarray(Image.fromarray(a).resize((ow, oh), Image.BICUBIC))
    got = resize_bicubic_u8(torch.from_numpy(a).to(DEV), oh, ow).cpu().numpy()
    assert np.array_equal(got, ref)


def test_gpu_process_images_and_regions_match_the_pinned_cpu_path():
    import cv2
    from PIL import Image
    from transformers import SiglipImageProcessor
    from types import SimpleNamespace

    from spatialrgpt_b200 import mm_utils as M
    proc = SiglipImageProcessor(size={"height": 448, "width": 448})
    cfg = SimpleNamespace(image_aspect_ratio="resize", image_processor=proc)
    rng = np.random.RandomState(5)
    imgs = [Image.fromarray(rng.randint(0, 256, (h, w, 3), dtype=np.uint8)) for (h, w) in ((480, 640), (333, 500))]
    got = M.process_images(imgs, proc, cfg, device=DEV)
    assert got.shape == (2, 3, 448, 448) and got.dtype == torch.float32 and got.is_cuda
    for i, im in enumerate(imgs):
        u8 = np.asarray(im.resize((448, 448), Image.BICUBIC))
        pinned = ((u8.astype(np.float64) * proc.rescale_factor).astype(np.float32) - np.float32(0.5)) / np.float32(0.5)
        assert np.array_equal(got[i].cpu().numpy(), pinned.transpose(2, 0, 1))
    # the installed transformers (torchvision backend) is within one 8-bit step of the pinned Pillow arithmetic
    cpu = M.process_images(imgs, proc, cfg)
    assert float((got.cpu() - cpu).abs().max()) <= 2.0 / 255 + 1e-6
    # pad mode: expand2square with the mean colour, then the same resize
    cfg_pad = SimpleNamespace(image_aspect_ratio="pad", image_processor=proc)
    gp = M.process_images(imgs[:1], proc, cfg_pad, device=DEV)
    sq = M._expand2square(imgs[0].convert("RGB"), tuple(int(x * 255) for x in proc.image_mean))
    u8 = np.asarray(sq.resize((448, 448), Image.BICUBIC))
    pinned = ((u8.astype(np.float64) * proc.rescale_factor).astype(np.float32) - np.float32(0.5)) / np.float32(0.5)
    assert np.array_equal(gp[0].cpu().numpy(), pinned.transpose(2, 0, 1))
    # regions: cv2 INTER_NEAREST
    masks = [(rng.rand(480, 640) > 0.5).astype(np.uint8), (rng.rand(97, 211) > 0.7).astype(np.uint8)]
    gr = M.process_regions(masks, proc, cfg, device=DEV)
    assert gr.shape == (2, 448, 448) and gr.dtype == torch.float32
    for i, m in enumerate(masks):
        assert np.array_equal(gr[i].cpu().numpy(), cv2.resize(m, (448, 448), interpolation=cv2.INTER_NEAREST).astype(np.float32))
