"""Import shim that lets the UNMODIFIED reference modules under /root/reference import in the
authoring container (transformers 5.5, no accelerate / deepspeed / s2wrapper / pycocotools).

TEST INFRASTRUCTURE ONLY — used by ``tests/golden/make_golden.py`` to generate the committed
golden fixtures and by ``tests/test_oracle_vs_reference.py`` (skipped when /root/reference is
absent, i.e. on the GPU box).  Recipe: SURVEY.md Appendix C.
"""
from __future__ import annotations

import os
import sys
import types

REFERENCE_ROOT = os.environ.get("SRGPT_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "llava"))


def _stub(name: str, **attrs) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


_INSTALLED = False


def install() -> None:
    """Idempotently install the stubs and put the reference on sys.path."""
    global _INSTALLED
    if _INSTALLED:
        return
    if not reference_available():
        raise RuntimeError(f"reference checkout not found at {REFERENCE_ROOT}")
    import torch
    import transformers
    import transformers.modeling_utils as mu

    # 1. moved/removed HF symbols (llava_arch.py:34)
    if not hasattr(mu, "no_init_weights"):
        from transformers import initialization as _init

        mu.no_init_weights = lambda _enable=True: _init.no_init_weights()
    if not hasattr(mu, "ContextManagers"):
        from transformers.utils import ContextManagers

        mu.ContextManagers = ContextManagers

    # 2. absent third-party packages
    if "accelerate" not in sys.modules:
        try:
            import accelerate  # noqa: F401
            import accelerate.hooks  # noqa: F401
        except Exception:
            acc = _stub("accelerate")
            acc.hooks = _stub("accelerate.hooks", add_hook_to_module=lambda module, hook, append=False: module)
    if "s2wrapper" not in sys.modules:
        def _s2_forward(*a, **k):
            raise NotImplementedError("s2wrapper is not installed")
        _stub("s2wrapper", forward=_s2_forward)
    if "pycocotools" not in sys.modules:
        pc = _stub("pycocotools")
        pc.mask = _stub("pycocotools.mask")
    if "deepspeed" not in sys.modules:
        ds = _stub("deepspeed")
        ds.comm = torch.distributed
        sys.modules["deepspeed.comm"] = torch.distributed
    if "cv2" not in sys.modules:
        try:
            import cv2  # noqa: F401
        except Exception:
            _stub("cv2", INTER_NEAREST=0)

    sys.path.insert(0, REFERENCE_ROOT)

    # 3. in-repo modules that import HF symbols which no longer exist
    class _Unavailable:
        def __init__(self, *a, **k):
            raise NotImplementedError("stubbed out by oracle/ref_shim.py")

    _stub("llava.model.multimodal_encoder.radio_encoder", RADIOVisionTower=_Unavailable)
    _stub("llava.model.multimodal_encoder.intern_encoder", InternVisionTower=_Unavailable,
          InternVisionTowerS2=_Unavailable)
    _stub("llava.model.language_model.llava_mistral", LlavaMistralConfig=_Unavailable,
          LlavaMistralForCausalLM=_Unavailable)
    _stub("llava.model.language_model.llava_mixtral", LlavaMixtralConfig=_Unavailable,
          LlavaMixtralForCausalLM=_Unavailable)
    _INSTALLED = True


def load_standalone(relpath: str, name: str):
    """Load one reference file by path without importing the ``llava`` package (works for
    base_extractor.py / base_projector.py, which only depend on torch/einops/transformers)."""
    import importlib.util

    spec = importlib.util.spec_from_file_location(name, os.path.join(REFERENCE_ROOT, relpath))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod
