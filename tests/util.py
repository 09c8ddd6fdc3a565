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


def write_synthetic_checkpoint(root, oc, sd, n_words=200, generation_eos=None, projector_type="mlp_downsample"):
    """A checkpoint directory in the reference's four-directory layout (llava_arch.py:181-250): top-level config.json, llm/
    (sharded safetensors + config + a real tokenizer + optional generation_config.json), vision_tower/ (+ preprocessor_config.json),
    mm_projector/, region_extractor/.  The tokenizer is a whitespace WordLevel fast tokenizer with `n_words` words + <s>/</s>/<unk>."""
    import json
    import os

    from safetensors.torch import save_file
    from tokenizers import Tokenizer, models, pre_tokenizers, processors
    from transformers import PreTrainedTokenizerFast, SiglipImageProcessor

    top = {"architectures": ["LlavaLlamaModel"], "model_type": "llava_llama", "enable_region": True, "enable_depth": bool(oc.enable_depth),
           "mm_vision_select_layer": oc.select_layer, "mm_vision_select_feature": getattr(oc, "select_feature", "cls_patch"), "image_aspect_ratio": "resize",
           "mm_use_im_start_end": False, "mm_use_im_patch_token": False, "model_dtype": "torch.bfloat16"}
    subs = {
        "llm": {"architectures": ["LlamaForCausalLM"], "model_type": "llama", "hidden_size": oc.hidden, "num_hidden_layers": oc.layers,
                "num_attention_heads": oc.heads, "num_key_value_heads": oc.kv_heads, "head_dim": oc.head_dim, "intermediate_size": oc.inter,
                "vocab_size": oc.vocab, "rope_theta": oc.rope_theta, "rms_norm_eps": oc.rms_eps, "max_position_embeddings": 2048,
                "bos_token_id": 1, "eos_token_id": 2, "tokenizer_padding_side": "right"},
        "vision_tower": {"model_type": "siglip_vision_model", "image_size": oc.image_size, "patch_size": oc.patch_size, "hidden_size": oc.v_hidden,
                         "num_hidden_layers": oc.v_layers, "num_attention_heads": oc.v_heads, "intermediate_size": oc.v_inter,
                         "layer_norm_eps": oc.v_eps, "hidden_act": "gelu_pytorch_tanh"} if getattr(oc, "v_type", "siglip") != "clip" else
                        {"architectures": ["CLIPVisionModel"], "model_type": "clip_vision_model", "image_size": oc.image_size,
                         "patch_size": oc.patch_size, "hidden_size": oc.v_hidden, "num_hidden_layers": oc.v_layers,
                         "num_attention_heads": oc.v_heads, "intermediate_size": oc.v_inter, "layer_norm_eps": oc.v_eps, "hidden_act": oc.v_act},
        "mm_projector": {"mm_projector_type": projector_type},
        "region_extractor": {"region_extractor_type": "regiongpt"},
    }
    os.makedirs(root)
    json.dump(top, open(os.path.join(root, "config.json"), "w"))
    for name, cfg in subs.items():
        d = os.path.join(root, name)
        os.makedirs(d)
        json.dump(cfg, open(os.path.join(d, "config.json"), "w"))
        tensors = {k: v.contiguous() for k, v in sd[name].items()}
        if name == "llm":  # sharded, like a real 8B checkpoint
            keys = sorted(tensors)
            save_file({k: tensors[k] for k in keys[::2]}, os.path.join(d, "model-00001-of-00002.safetensors"))
            save_file({k: tensors[k] for k in keys[1::2]}, os.path.join(d, "model-00002-of-00002.safetensors"))
        else:
            save_file(tensors, os.path.join(d, "model.safetensors"))
    vocab = {"<unk>": 0, "<s>": 1, "</s>": 2}
    for i in range(n_words):
        vocab[f"w{i}"] = len(vocab)
    for w in ("USER:", "ASSISTANT:", "A", "chat", "between", "a", "curious", "user", "and", "an", "artificial", "intelligence", "assistant.", "how", "far",
              "is", "from", "?"):
        vocab.setdefault(w, len(vocab))
    tk = Tokenizer(models.WordLevel(vocab, unk_token="<unk>"))
    tk.pre_tokenizer = pre_tokenizers.WhitespaceSplit()
    tk.post_processor = processors.TemplateProcessing(single="<s> $A", special_tokens=[("<s>", 1)])
    fast = PreTrainedTokenizerFast(tokenizer_object=tk, bos_token="<s>", eos_token="</s>", unk_token="<unk>")
    fast.save_pretrained(os.path.join(root, "llm"))
    SiglipImageProcessor(size={"height": oc.image_size, "width": oc.image_size}).save_pretrained(os.path.join(root, "vision_tower"))
    if generation_eos is not None:
        json.dump({"eos_token_id": generation_eos, "do_sample": True, "temperature": 0.6, "top_p": 0.9}, open(os.path.join(root, "llm", "generation_config.json"), "w"))
    return len(vocab)
