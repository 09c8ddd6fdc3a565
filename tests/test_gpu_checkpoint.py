"""a18 / f1: a synthetic checkpoint in the reference's four-directory layout (llava_arch.py:181-250) goes through
``load_pretrained_model`` (builder.py:36-213: sub-directory resolution, safetensors shards, tokenizer + <mask>/<depth> registration,
token-table resize, image processor, generation_config stop ids) and then through the caller's flow of eval_spatial.py:196-237 -
conversation template -> tokenizer_image_token -> process_images / region masks -> ``generate()`` on the CUDA path - and must
reproduce the CPU oracle run on the very weights the loader read."""
import numpy as np
import pytest
import torch

from oracle import srgpt_oracle as O
from tests.golden.make_golden import CASES
from tests.util import write_synthetic_checkpoint

pytestmark = pytest.mark.gpu


def test_load_pretrained_model_then_generate(tmp_path):
    from PIL import Image

    from llava.constants import IMAGE_TOKEN_INDEX
    from llava.conversation import conv_templates
    from llava.mm_utils import KeywordsStoppingCriteria, process_images, process_regions, tokenizer_image_token
    from llava.model.builder import load_pretrained_model
    from spatialrgpt_b200 import builder

    oc = O.OracleConfig(**CASES["tiny_masks_gqa"][0])
    sd = O.make_weights(oc, seed=3)
    root = str(tmp_path / "SpatialRGPT-tiny")
    write_synthetic_checkpoint(root, oc, sd, generation_eos=[2])
    tokenizer, model, image_processor, context_len = load_pretrained_model(root, "SpatialRGPT-tiny", None)
    assert context_len == 2048 and model.config.llm_mask_token_id == tokenizer.convert_tokens_to_ids("<mask>")
    assert model.config.llama.vocab_size == len(tokenizer)
    assert model.dtype == torch.float16  # builder.py:62: the loader hands out fp16 ...

    # ---- the caller's flow (eval_spatial.py:196-237)
    rng = np.random.RandomState(3)
    image = Image.fromarray(rng.randint(0, 255, (90, 120, 3), dtype=np.uint8))
    depth = Image.fromarray(np.repeat(rng.randint(0, 255, (90, 120, 1), dtype=np.uint8), 3, axis=2))
    region = [np.zeros((90, 120), dtype=np.uint8) for _ in range(2)]
    region[0][10:50, 20:70] = 1
    region[1][40:85, 60:110] = 1
    model.config.image_processor = image_processor
    images32 = process_images([image], image_processor, model.config)
    depths32 = process_images([depth], image_processor, model.config)
    masks = process_regions(region, image_processor, model.config)
    conv = conv_templates["llava_v1"].copy()
    conv.append_message(conv.roles[0], "<image>\n how far is <mask> <depth> from <mask> <depth> ?")
    conv.append_message(conv.roles[1], None)
    input_ids = tokenizer_image_token(conv.get_prompt(), tokenizer, IMAGE_TOKEN_INDEX, return_tensors="pt").unsqueeze(0)
    assert int((input_ids == IMAGE_TOKEN_INDEX).sum()) == 1 and int((input_ids == model.config.llm_mask_token_id).sum()) == 2
    stopping = KeywordsStoppingCriteria(["</s>"], tokenizer, input_ids)
    n_new = 10

    def run():
        dev, dt = model.device, model.dtype
        return model.generate(input_ids.to(dev), images=images32.to(dev, dtype=dt), depths=depths32.to(dev, dtype=dt), masks=[masks.to(dev, dtype=dt)],
                              do_sample=False, temperature=0, max_new_tokens=n_new, use_cache=True, stopping_criteria=[stopping])

    got16 = run()[0].tolist()               # ... which is how eval_region_cls.py:316-317 runs it
    model.to(dtype=torch.bfloat16)          # ... and eval_spatial.py:221 casts to bf16 before generating
    assert model.dtype == torch.bfloat16
    out = run()
    got = out[0].tolist()
    images, depths = images32.to(torch.bfloat16), depths32.to(torch.bfloat16)

    # ---- the oracle on the weights the loader actually read (token tables resized to len(tokenizer))
    cfg2, sd2, _, _ = builder.read_checkpoint(root)
    oc2 = O.OracleConfig(**{**CASES["tiny_masks_gqa"][0], "vocab": len(tokenizer), "mask_token_id": cfg2.llm_mask_token_id,
                            "depth_token_id": cfg2.llm_depth_token_id})
    bf = lambda t: t.to(torch.bfloat16).float()  # noqa: E731  (the inputs the CUDA path saw)
    ref_ids, enc = O.generate(oc2, sd2, input_ids, bf(images.cpu()), bf(depths.cpu()), [bf(masks)], n_new, eos_token_id=[2], return_all=True)
    top2 = enc["logits"].topk(2, -1).values
    margin = top2[:, 0] - top2[:, 1]
    tol = 0.06 * float(enc["logits"].std())
    safe = int((margin > 2 * tol).long().cumprod(0).sum())
    assert safe >= 1
    assert got[:safe] == ref_ids.tolist()[:safe], (got, ref_ids.tolist(), margin.tolist())
    assert got16[:safe] == ref_ids.tolist()[:safe], (got16, ref_ids.tolist(), margin.tolist())
    text = tokenizer.batch_decode(out, skip_special_tokens=True)[0]
    assert isinstance(text, str)
