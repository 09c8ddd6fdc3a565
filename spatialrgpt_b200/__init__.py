"""spatialrgpt_b200 — B200-native (sm_100a) implementation of SpatialRGPT's multimodal generate()
hot path behind the reference's Python API.  See DESIGN.md / INTEGRATION.md."""
from .config import LlavaConfig, LlavaLlamaConfig, VisionConfig, LlamaDims, baseline_config  # noqa: F401
from .constants import *  # noqa: F401,F403
from ._lib import SrgptError  # noqa: F401


def __getattr__(name):
    # heavy modules are imported lazily so that `import spatialrgpt_b200` works on a CPU-only box
    if name in ("LlavaLlamaModel", "LlavaLlamaForCausalLM"):
        from .llava_llama import LlavaLlamaModel
        return LlavaLlamaModel
    if name == "load_pretrained_model":
        from .builder import load_pretrained_model
        return load_pretrained_model
    raise AttributeError(name)
