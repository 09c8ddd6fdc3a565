from .utils import disable_torch_init  # noqa: F401
