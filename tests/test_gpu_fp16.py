"""fp16 mode (the reference loader's default dtype, llava/model/builder.py:62; eval_region_cls.py:316-317 feeds fp16 inputs): the IEEE
half build of the kernels (libsrgpt_b200_f16.so, csrc/common.cuh) behind the same host API.

Bars: the golden fixtures produced by the reference's own modules in fp32 (tests/golden) with tolerances 4x TIGHTER than the bf16 ones
(fp16 carries 3 more mantissa bits; the rounding points are the same), ids exact; op-level checks against torch fp32 with the same
rounding points; the oracle in fp16 mode as a second opinion on the logits."""
import os

import pytest
import torch

from oracle import srgpt_oracle as O
from tests.golden.make_golden import CASES
from tests.util import assert_close, load_npz

pytestmark = pytest.mark.gpu
DEV = "cuda"
F16 = torch.float16
# one fp16 rounding of an fp32-accumulated result: rms 2^-12, max 2^-11 of the value; a network stage: bf16's bounds / 4
F16_OP = dict(rel_rms=1e-3, rel_max=2e-2)
F16_STAGE = dict(rel_rms=1.25e-2, rel_max=1.5e-1)
F16_LOGIT_SIGMA = 0.0075  # bf16: 0.06; measured: ~0.002 (the oracle in fp16 mode: 0.0021)


def build_model(case_kw, weight_seed, dtype=F16, max_seq_len=512):
    from spatialrgpt_b200 import LlavaConfig, LlamaDims, VisionConfig
    from spatialrgpt_b200.llava_llama import LlavaLlamaModel
    from spatialrgpt_b200.weights import from_state_dicts

    oc = O.OracleConfig(**case_kw)
    cfg = LlavaConfig(
        vision=VisionConfig(image_size=oc.image_size, patch_size=oc.patch_size, hidden_size=oc.v_hidden,
                            num_hidden_layers=oc.v_layers, num_attention_heads=oc.v_heads, intermediate_size=oc.v_inter,
                            layer_norm_eps=oc.v_eps),
        llama=LlamaDims(hidden_size=oc.hidden, num_hidden_layers=oc.layers, num_attention_heads=oc.heads,
                        num_key_value_heads=oc.kv_heads, head_dim=oc.head_dim, intermediate_size=oc.inter, vocab_size=oc.vocab,
                        rope_theta=oc.rope_theta, rms_norm_eps=oc.rms_eps),
        enable_region=oc.enable_region, enable_depth=oc.enable_depth, mm_vision_select_layer=oc.select_layer)
    cfg.llm_mask_token_id, cfg.llm_depth_token_id = oc.mask_token_id, oc.depth_token_id
    sd = O.make_weights(oc, seed=weight_seed)
    model = LlavaLlamaModel(cfg, from_state_dicts(cfg, sd, DEV, dtype=dtype), max_seq_len=max_seq_len)
    return oc, sd, model


def test_fp16_gemm_epilogues_against_torch():
    from spatialrgpt_b200 import ops
    g = torch.Generator(device=DEV).manual_seed(5)
    rn = lambda *s, k=1.0: (torch.randn(*s, generator=g, device=DEV) * k).to(F16)  # noqa: E731
    with ops.elem_dtype(F16):
        for (M, N, K) in [(200, 256, 320), (2048, 1152, 1152), (37, 512, 4096), (4096, 1024, 512)]:
            a, w, b, r = rn(M, K), rn(N, K, k=K ** -0.5), rn(N, k=0.1), rn(M, N)
            acc = a.float() @ w.float().t()
            rnd = lambda t: t.to(F16).float()  # noqa: E731
            assert_close(ops.gemm(a, w), acc, **F16_OP, what=f"gemm {M}x{N}x{K}")
            assert_close(ops.gemm(a, w, bias=b, epilogue=ops.EPI_BIAS), acc + b.float(), **F16_OP, what="bias")
            assert_close(ops.gemm(a, w, bias=b, residual=r, epilogue=ops.EPI_BIAS_RESIDUAL), rnd(acc + b.float()) + r.float(), **F16_OP,
                         what="bias+residual")
            assert_close(ops.gemm(a, w, bias=b, epilogue=ops.EPI_BIAS_GELU_ERF), torch.nn.functional.gelu(rnd(acc + b.float())), **F16_OP,
                         what="gelu_erf")
            got = ops.gemm(a, w, bias=b, epilogue=ops.EPI_BIAS_GELU_TANH)
            assert_close(got, torch.nn.functional.gelu(rnd(acc + b.float()), approximate="tanh"), rel_rms=2e-3, rel_max=2e-2, what="gelu_tanh")
            if N % 2 == 0:
                sw = ops.gemm(a, w, epilogue=ops.EPI_SWIGLU)
                gate, up = rnd(acc[:, 0::2]), rnd(acc[:, 1::2])
                assert_close(sw, rnd(torch.nn.functional.silu(gate)) * up, **F16_OP, what="swiglu")
        # the output really is fp16 and a bf16 tensor is rejected loudly while the fp16 build is selected
        assert ops.gemm(rn(8, 64), rn(16, 64)).dtype == F16
        with pytest.raises(Exception):
            ops.gemm(rn(8, 64).to(torch.bfloat16), rn(16, 64).to(torch.bfloat16))


def test_fp16_rowops_attention_and_decode_ops_against_torch():
    from spatialrgpt_b200 import ops
    g = torch.Generator(device=DEV).manual_seed(6)
    rn = lambda *s, k=1.0: (torch.randn(*s, generator=g, device=DEV) * k).to(F16)  # noqa: E731
    with ops.elem_dtype(F16):
        x, w, b = rn(300, 1152), rn(1152, k=0.2) + 1, rn(1152, k=0.1)
        ref = torch.nn.functional.layer_norm(x.float(), (1152,), w.float(), b.float(), 1e-6)
        assert_close(ops.layernorm(x, w, b, 1e-6), ref, **F16_OP, what="layernorm")
        xf = x.float()
        ref = (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + 1e-5)).to(F16).float() * w.float()
        assert_close(ops.rmsnorm(x, w, 1e-5), ref, **F16_OP, what="rmsnorm")
        # SigLIP attention (hd 72, non-causal, dense) and the causal GQA prefill (hd 128)
        for (B, S, nh, nkv, hd, causal) in [(3, 1024, 16, 16, 72, False), (2, 259, 32, 8, 128, True)]:
            qkv = rn(B * S, (nh + 2 * nkv) * hd, k=0.7)
            q, k, v = qkv[:, :nh * hd], qkv[:, nh * hd:(nh + nkv) * hd], qkv[:, (nh + nkv) * hd:]
            out = ops.attention_prefill(q, k, v, B, S, nh, nkv, hd, hd ** -0.5, causal=causal)
            qf = q.float().view(B, S, nh, hd).transpose(1, 2)
            kf = k.float().view(B, S, nkv, hd).transpose(1, 2).repeat_interleave(nh // nkv, 1)
            vf = v.float().view(B, S, nkv, hd).transpose(1, 2).repeat_interleave(nh // nkv, 1)
            att = qf @ kf.transpose(-1, -2) * hd ** -0.5
            if causal:
                att = att + torch.full((S, S), float("-inf"), device=DEV).triu(1)
            p = torch.softmax(att, -1).to(F16).float()
            ref = (p @ vf).transpose(1, 2).reshape(B * S, nh * hd)
            assert_close(out, ref, rel_rms=2e-3, rel_max=3e-2, what=f"attention hd{hd}")
        # decode GEMV with fused RMSNorm and residual
        K, N = 4096, 1024
        xv, wm, nw, res = rn(K), rn(N, K, k=K ** -0.5), rn(K, k=0.1) + 1, rn(N)
        y = torch.empty(N, dtype=F16, device=DEV)
        ops.gemv(xv, wm, y, norm_weight=nw, eps=1e-5, residual=res)
        xn = (xv.float() * torch.rsqrt(xv.float().pow(2).mean() + 1e-5)).to(F16).float() * nw.float()
        ref = (xn.to(F16).float() @ wm.float().t()).to(F16).float() + res.float()
        assert_close(y, ref, **F16_OP, what="gemv")


@pytest.mark.parametrize("name", list(CASES))
def test_fp16_stages_and_tokens_match_reference_fixture(golden_dir, name):
    kw, n_regions, t_text, kind, n_new, depth_on = CASES[name]
    g = load_npz(os.path.join(golden_dir, name + ".npz"))
    oc, sd, model = build_model(kw, int(g["weight_seed"]))
    assert model.dtype == F16 and model.config.model_dtype == "torch.float16"
    input_ids, images, depths, masks = O.synth_request(oc, n_regions, t_text, seed=1234, kind=kind)
    if not depth_on:
        depths = None
    # the callers cast the inputs to the model dtype (eval_region_cls.py:316-317)
    imd = images.to(DEV, dtype=F16)
    dd = None if depths is None else depths.to(DEV, dtype=F16)
    md = [m.to(DEV, dtype=F16) for m in masks]

    tower = model.get_vision_tower()(imd)
    assert tower.dtype == F16
    assert_close(tower, g["tower_features"], **F16_STAGE, what="tower_features")
    hres, lres = model.get_region_extractor().feature_refinement(tower)
    assert_close(hres, g["hres"], **F16_STAGE, what="hres")
    assert_close(lres, g["lres"], **F16_STAGE, what="lres")
    dfeat = model.get_vision_tower()(dd) if dd is not None else None
    with __import__("spatialrgpt_b200").ops.elem_dtype(F16):  # module-level call outside generate(): select the build explicitly
        me, de = model.get_region_extractor()(hres, dfeat, md)
        feats = model.get_mm_projector()(lres)
    # A region that is empty AFTER the bilinear resize (tiny_nodepth's 1-pixel mask) is where fp16 differs from bf16 in the reference
    # itself: denorm = sum + 1e-8 (base_extractor.py:66) rounds to 0 in fp16 (1e-8 is below half the smallest subnormal), so the row is
    # 0/0 = NaN, while bf16 keeps 1e-8 and gives 0.  The kernels reproduce that (the oracle in fp16 mode shows the same rows).
    o_me = O.mask_pooling(g["hres"].to(F16), [m.to(F16) for m in masks])[0]
    nan_rows = torch.isnan(o_me.float()).any(-1)
    assert torch.equal(torch.isnan(me[0].float()).any(-1).cpu(), nan_rows), "NaN rows differ from the fp16 oracle"
    assert_close(me[0][~nan_rows.to(DEV)], g["mask_embeds"][~nan_rows], **F16_STAGE, what="mask_embeds")
    if depth_on:
        assert_close(de[0][~nan_rows.to(DEV)], g["depth_embeds"][~nan_rows], **F16_STAGE, what="depth_embeds")
    assert_close(feats, g["image_features"], **F16_STAGE, what="image_features")
    if bool(nan_rows.any()):
        assert name == "tiny_nodepth"
        return  # NaN embeddings poison the prompt (in the reference too): nothing meaningful to compare downstream

    ids, logits = model.generate(input_ids.to(DEV), images=imd, depths=dd, masks=md, do_sample=False, max_new_tokens=n_new,
                                 use_cache=True, output_logits=True)
    ref_ids = g["new_ids"].tolist()
    sigma = float(g["logits"].std())
    err = (logits[0].cpu() - g["logits"]).abs().max().item()
    print(f"fp16 {name}: logit err {err / sigma:.4f} sigma")
    assert err <= F16_LOGIT_SIGMA * sigma, f"logit error {err:.4f} > {F16_LOGIT_SIGMA} * sigma ({sigma:.3f})"
    assert ids[0].tolist() == ref_ids
    # CUDA-graph decode path, twice (page recycling)
    for _ in range(2):
        ids2 = model.generate(input_ids.to(DEV), images=imd, depths=dd, masks=md, do_sample=False, max_new_tokens=n_new)
        assert ids2[0].tolist() == ref_ids


def test_fp16_and_bf16_models_in_one_process_and_the_cast():
    """A bf16 and an fp16 model coexist (each public call selects its build); ``model.to(dtype=torch.bfloat16)`` after an fp16 load -
    the reference's eval flow (builder.py:62 then eval_spatial.py:221) - gives the ids of the bf16 fixture."""
    name = "tiny_masks_gqa"
    kw, n_regions, t_text, kind, n_new, depth_on = CASES[name]
    g = load_npz(os.path.join(os.path.dirname(__file__), "golden", name + ".npz"))
    oc, sd, m16 = build_model(kw, int(g["weight_seed"]), dtype=F16)
    _, _, mbf = build_model(kw, int(g["weight_seed"]), dtype=torch.bfloat16)
    input_ids, images, depths, masks = O.synth_request(oc, n_regions, t_text, seed=1234, kind=kind)
    ref_ids = g["new_ids"].tolist()

    def run(m):
        dt = m.dtype
        return m.generate(input_ids.to(DEV), images=images.to(DEV, dtype=dt), depths=depths.to(DEV, dtype=dt),
                          masks=[x.to(DEV, dtype=dt) for x in masks], do_sample=False, max_new_tokens=n_new)[0].tolist()

    for _ in range(2):  # interleaved
        assert run(m16) == ref_ids
        assert run(mbf) == ref_ids
    m16.to(dtype=torch.bfloat16)
    assert m16.dtype == torch.bfloat16 and m16.config.model_dtype == "torch.bfloat16"
    assert run(m16) == ref_ids
    with pytest.raises(NotImplementedError):
        m16.to(dtype=torch.float64)


def test_fp16_logits_agree_with_the_oracle_in_fp16_mode():
    """Second opinion: the oracle run with dtype=float16 (torch CPU half arithmetic) is as far from the fp32 truth as the CUDA path is."""
    name = "tiny_boxes"
    kw, n_regions, t_text, kind, n_new, depth_on = CASES[name]
    g = load_npz(os.path.join(os.path.dirname(__file__), "golden", name + ".npz"))
    oc, sd, model = build_model(kw, int(g["weight_seed"]))
    input_ids, images, depths, masks = O.synth_request(oc, n_regions, t_text, seed=1234, kind=kind)
    ids, logits = model.generate(input_ids.to(DEV), images=images.to(DEV, dtype=F16), depths=depths.to(DEV, dtype=F16),
                                 masks=[m.to(DEV, dtype=F16) for m in masks], do_sample=False, max_new_tokens=n_new, output_logits=True)
    o_ids, enc = O.generate(oc, sd, input_ids, images, depths, masks, n_new, dtype=F16, return_all=True)
    sigma = float(g["logits"].std())
    e_cuda = (logits[0].cpu() - g["logits"]).abs().max().item() / sigma
    e_orac = (enc["logits"].float() - g["logits"]).abs().max().item() / sigma
    print(f"fp16 error vs fp32 fixture: cuda {e_cuda:.4f} sigma, oracle-fp16 {e_orac:.4f} sigma")
    assert e_cuda <= max(1.5 * e_orac, 0.005)
    assert ids[0].tolist() == o_ids.tolist() == g["new_ids"].tolist()
