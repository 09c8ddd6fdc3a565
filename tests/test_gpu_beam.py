"""Beam search (num_beams > 1 through llava_llama.py:212 -> HF GenerationMixin.beam_search): the candidates kernel against torch, and the
whole device/host loop against ids produced by HF's own generate() (tests/golden/beam_kats.npz, made by ``make_golden.py beam``)."""
import os

import pytest
import torch

from oracle import srgpt_oracle as O
from tests.golden.make_golden import BEAM_CASES, CASES
from tests.util import load_npz

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("V,k,n_cand", [(1003, 3, 6), (128259, 4, 8), (4096, 5, 15)])
def test_beam_candidates_kernel_against_torch(V, k, n_cand):
    from spatialrgpt_b200 import ops
    g = torch.Generator().manual_seed(V)
    logits = (torch.randn(k, V, generator=g) * 3).to(torch.bfloat16)
    logits[1, 7] = float("nan")  # a NaN never becomes a candidate
    scores = torch.tensor([0.0, -1.5, -0.25, -7.0, -1e9][:k])
    ld = (V + 7) // 8 * 8
    d_logits = torch.zeros(k, ld, dtype=torch.bfloat16, device=DEV)[:, :V]
    d_logits.copy_(logits)
    cs = torch.empty(k, n_cand, dtype=torch.float32, device=DEV)
    ct = torch.empty(k, n_cand, dtype=torch.int32, device=DEV)
    ops.beam_candidates(d_logits, scores.to(DEV), cs, ct)
    x = torch.nan_to_num(logits.float(), nan=float("-inf"))
    table = torch.log_softmax(x, -1) + scores[:, None]
    for b in range(k):
        # reference order: score descending, token ascending on ties (bf16 logits tie often)
        order = sorted(range(V), key=lambda t: (-float(x[b, t]), t))[:n_cand]
        assert ct[b].tolist() == order, (b, ct[b].tolist(), order)
        assert torch.allclose(cs[b].cpu(), table[b, order], rtol=0, atol=2e-5 * max(1.0, float(table[b, order].abs().max())))


def _model(dtype):
    from tests.test_gpu_fp16 import build_model
    g = load_npz(os.path.join(os.path.dirname(__file__), "golden", "beam_kats.npz"))
    oc, sd, model = build_model(CASES["tiny_masks_gqa"][0], int(g["weight_seed"]), dtype=dtype)
    return g, oc, sd, model


@pytest.mark.parametrize("dtype,need", [(torch.float16, 9), (torch.bfloat16, 7)])
def test_beam_search_matches_hf_generate(dtype, need):
    """fp16 reproduces every fixture case; bf16 is allowed the near-tie flips the reference's own bf16 arithmetic shows (the oracle in
    bf16 mode flips case 3)."""
    g, oc, sd, model = _model(dtype)
    emb = g["inputs_embeds"].to(DEV)
    ok = []
    for i, (nb, eos, n_new, lp, es) in enumerate(BEAM_CASES):
        ref = g[f"case{i}"].tolist()
        if nb == 1:
            ids = model.llm.generate_from_embeds(emb, n_new, eos_token_ids=eos).tolist()
        else:
            ids = model.llm.generate_beam(emb, nb, n_new, eos_token_ids=eos, length_penalty=lp, early_stopping=es).tolist()
            eager = model.llm.generate_beam(emb, nb, n_new, eos_token_ids=eos, length_penalty=lp, early_stopping=es, use_graph=False).tolist()
            assert eager == ids, (i, eager, ids)  # CUDA-graph and eager steps agree
        ok.append(ids == ref[: len(ids)] and all(t == 0 for t in ref[len(ids):]))
    print(f"beam {dtype}: cases equal to HF generate: {ok}")
    assert sum(ok) >= need, ok


def test_generate_api_with_num_beams_and_stopping_criteria(golden_dir):
    """The caller's form (eval_spatial.py:229-237 with --num_beams 3): generate(input_ids, images=, depths=, masks=, num_beams=3) equals the
    oracle's beam search over the oracle's spliced prompt on a margin-safe case, and stops when every beam satisfies the criterion."""
    from tests.test_gpu_fp16 import build_model
    name = "tiny_masks_gqa"
    kw, n_regions, t_text, kind, n_new, depth_on = CASES[name]
    g = load_npz(os.path.join(golden_dir, name + ".npz"))
    oc, sd, model = build_model(kw, int(g["weight_seed"]), dtype=torch.float16)
    input_ids, images, depths, masks = O.synth_request(oc, n_regions, t_text, seed=1234, kind=kind)
    enc = O.encode_multimodal(oc, sd, images, depths, masks)
    embeds = O.splice_embeddings(oc, sd["llm"]["model.embed_tokens.weight"].float(), input_ids, enc["image_features"], enc["mask_embeds"],
                                 enc["depth_embeds"])[0]
    ref = O.beam_search_generate(oc, sd["llm"], embeds, 3, 8).tolist()
    args = dict(images=images.to(DEV, torch.float16), depths=depths.to(DEV, torch.float16), masks=[m.to(DEV, torch.float16) for m in masks])
    out = model.generate(input_ids.to(DEV), num_beams=3, do_sample=False, max_new_tokens=8, **args)[0].tolist()
    assert out == ref, (out, ref)

    class StopAfter3:
        def __call__(self, output_ids, scores=None, **kw):
            return output_ids.shape[1] >= 3

    short = model.generate(input_ids.to(DEV), num_beams=3, do_sample=False, max_new_tokens=8, stopping_criteria=[StopAfter3()], **args)[0].tolist()
    assert len(short) == 3
    with pytest.raises(NotImplementedError):
        model.generate(input_ids.to(DEV), num_beams=3, do_sample=True, temperature=0.7, max_new_tokens=4, **args)
