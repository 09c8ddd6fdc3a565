"""Host-side weight re-layouts (spatialrgpt_b200/weights.py) checked against the torch ops of the reference they stand for -
the GEMM-shaped forms the CUDA kernels consume must be exact re-orderings, not approximations.  CPU only."""
import torch
import torch.nn.functional as F

from oracle import srgpt_oracle as O
from spatialrgpt_b200 import LlavaConfig, LlamaDims, VisionConfig
from spatialrgpt_b200.weights import _deconv_as_gemm, from_state_dicts, interleave_rows, patch_ldk


def _nested_row(y, x, P):
    """row of pixel (y, x) of the 4x up-sampled map in the nested 2x2 order the two deconv GEMMs emit."""
    return ((y >> 2) * P + (x >> 2)) * 16 + (((y >> 1) & 1) * 2 + ((x >> 1) & 1)) * 4 + ((y & 1) * 2 + (x & 1))


def test_deconv_as_gemm_equals_conv_transpose_in_nested_order():
    """ConvTranspose2d(k=2, s=2) x2 (base_extractor.py:92-97, without the LayerNorm / GELU in between) == two GEMMs whose
    output rows are re-read as [4x rows, C]; the resulting pixel order is the nested order mask pooling indexes."""
    torch.manual_seed(0)
    P, C = 3, 8
    x = torch.randn(1, C, P, P)
    w1, b1 = torch.randn(C, C, 2, 2), torch.randn(C)
    w2, b2 = torch.randn(C, C, 2, 2), torch.randn(C)
    ref = F.conv_transpose2d(F.conv_transpose2d(x, w1, b1, stride=2), w2, b2, stride=2)  # [1, C, 4P, 4P]
    rows = x[0].permute(1, 2, 0).reshape(P * P, C)                        # tokens in raster order, channels last
    y1 = (rows @ _deconv_as_gemm(w1).t() + b1.repeat(4)).reshape(P * P * 4, C)
    y2 = (y1 @ _deconv_as_gemm(w2).t() + b2.repeat(4)).reshape(P * P * 16, C)
    for yy in range(4 * P):
        for xx in range(4 * P):
            assert torch.allclose(y2[_nested_row(yy, xx, P)], ref[0, :, yy, xx], atol=1e-4), (yy, xx)


def test_interleaved_gate_up_is_swiglu_local():
    torch.manual_seed(1)
    H, I = 16, 24
    gate, up, x = torch.randn(I, H), torch.randn(I, H), torch.randn(5, H)
    y = x @ interleave_rows(gate, up).t()            # columns 2i = gate_i, 2i+1 = up_i (the SwiGLU epilogue pairs them)
    assert torch.allclose(F.silu(y[:, 0::2]) * y[:, 1::2], F.silu(x @ gate.t()) * (x @ up.t()), atol=1e-5)


def test_from_state_dicts_fuses_and_pads_exactly():
    kw = dict(image_size=28, patch_size=14, v_hidden=16, v_layers=2, v_heads=2, v_inter=24, hidden=32, layers=1, heads=4, kv_heads=2,
              head_dim=8, inter=40, vocab=48)
    oc = O.OracleConfig(**kw)
    sd = O.make_weights(oc, seed=2)
    cfg = LlavaConfig(vision=VisionConfig(image_size=28, patch_size=14, hidden_size=16, num_hidden_layers=2, num_attention_heads=2, intermediate_size=24),
                      llama=LlamaDims(hidden_size=32, num_hidden_layers=1, num_attention_heads=4, num_key_value_heads=2, head_dim=8,
                                      intermediate_size=40, vocab_size=48))
    w = from_state_dicts(cfg, sd, "cpu")
    v, l = sd["vision_tower"], sd["llm"]
    bf = lambda t: t.to(torch.bfloat16)
    # patch embedding: conv weight [D, 3, 14, 14] flattened and zero-padded to a 16-byte row stride
    assert w.vision.patch_w.shape == (16, patch_ldk(14)) and patch_ldk(14) % 8 == 0
    assert torch.equal(w.vision.patch_w[:, :588], bf(v["vision_model.embeddings.patch_embedding.weight"]).reshape(16, -1))
    assert float(w.vision.patch_w[:, 588:].abs().sum()) == 0.0
    # SigLIP q/k/v fused along the output dimension, in that order
    p = "vision_model.encoder.layers.0.self_attn."
    assert torch.equal(w.vision.layers[0].qkv_w, torch.cat([bf(v[p + n + ".weight"]) for n in ("q_proj", "k_proj", "v_proj")], 0))
    assert torch.equal(w.vision.layers[0].qkv_b, torch.cat([bf(v[p + n + ".bias"]) for n in ("q_proj", "k_proj", "v_proj")], 0))
    # Llama: [q; k; v] rows and interleaved gate/up
    q = "model.layers.0."
    assert w.llama.layers[0].qkv_w.shape == ((4 + 2 * 2) * 8, 32)
    assert torch.equal(w.llama.layers[0].qkv_w[:32], bf(l[q + "self_attn.q_proj.weight"]))
    assert torch.equal(w.llama.layers[0].gateup_w[0::2], bf(l[q + "mlp.gate_proj.weight"]))
    assert torch.equal(w.llama.layers[0].gateup_w[1::2], bf(l[q + "mlp.up_proj.weight"]))
    # deconv biases are repeated once per (di, dj) block of the GEMM output
    r = sd["region_extractor"]
    assert torch.equal(w.region.deconv1_b, bf(r["feature_refinement_module.0.bias"]).repeat(4))
    assert w.nbytes() > 0
