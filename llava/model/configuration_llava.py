from spatialrgpt_b200.config import LlavaConfig  # noqa: F401
