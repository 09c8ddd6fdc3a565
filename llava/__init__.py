"""Drop-in alias package: the reference's import paths (``llava.*``) resolved to the B200-native
implementation in ``spatialrgpt_b200`` for the generate() hot path.  Only the modules on that path exist."""
from spatialrgpt_b200.llava_llama import LlavaLlamaModel  # noqa: F401
