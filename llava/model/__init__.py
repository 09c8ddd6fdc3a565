from spatialrgpt_b200.config import LlavaConfig, LlavaLlamaConfig  # noqa: F401
from spatialrgpt_b200.llava_llama import LlavaLlamaForCausalLM, LlavaLlamaModel  # noqa: F401
