"""CPU-side checks of the C-ABI boundary: the library builds/loads and exports exactly the symbols
include/srgpt_b200.h declares (no compute calls without a GPU)."""
import os
import re
import subprocess

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
HEADER = os.path.join(ROOT, "include", "srgpt_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(srgpt_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_expected_groups():
    syms = declared_symbols()
    for must in ("srgpt_gemm_bf16", "srgpt_mask_pool_bf16", "srgpt_attention_prefill_bf16", "srgpt_attention_decode_bf16",
                 "srgpt_gemv_bf16", "srgpt_lm_head_argmax_bf16", "srgpt_depth_to_u8x3"):
        assert must in syms


@pytest.mark.parametrize("elem", ["bf16", "f16"])
def test_library_loads_and_exports_every_declared_symbol(elem):
    """Both builds of the kernels (bfloat16 / IEEE half elements, csrc/common.cuh) export the same C-ABI."""
    from spatialrgpt_b200 import _lib
    lib = _lib.load(elem=elem)
    assert lib.srgpt_abi_version() == 1 and lib.srgpt_elem_type() == {"bf16": 0, "f16": 1}[elem]
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.lib_path(elem)], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r"\sT\s+(srgpt_[a-z0-9_]+)", out))
    declared = set(declared_symbols())
    assert declared <= exported, f"declared but not exported: {sorted(declared - exported)}"
    assert exported <= declared, f"exported but not declared in the header: {sorted(exported - declared)}"
    assert set(_lib.SIGNATURES) == declared, "ctypes signature table out of sync with the header"


def test_library_contains_blackwell_sass():
    """The shipped .so must carry sm_100a code with tcgen05 / TMA instructions (UTCHMMA, UTMALDG, LDTM)."""
    from spatialrgpt_b200 import _lib
    _lib.load()
    r = subprocess.run(["cuobjdump", "-sass", _lib.lib_path()], capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in r.stdout
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM"):
        assert mnemonic in r.stdout, f"{mnemonic} missing from SASS"


def test_host_argument_validation_needs_no_gpu():
    """Bad arguments are rejected on the host before any CUDA call."""
    from spatialrgpt_b200 import _lib
    lib = _lib.load()
    rc = lib.srgpt_gemm_bf16(None, 0, None, 0, None, 0, 1, 1, 8, None, None, 0, 0, 0, 0, None)
    assert rc == -1 and "invalid argument" in _lib.last_error()
    assert lib.srgpt_mask_pool_workspace(0, 1, 1, 8) == -1
    assert lib.srgpt_lm_head_workspace(128259) > 0


def test_element_type_switch_selects_the_build():
    import torch

    from spatialrgpt_b200 import _lib, ops
    assert ops.ELEM() == torch.bfloat16 and _lib.load().srgpt_elem_type() == 0
    with ops.elem_dtype(torch.float16):
        assert ops.ELEM() == torch.float16 and _lib.load().srgpt_elem_type() == 1
        with ops.elem_dtype(torch.bfloat16):
            assert _lib.load().srgpt_elem_type() == 0
        assert _lib.load().srgpt_elem_type() == 1
    assert ops.ELEM() == torch.bfloat16
    with pytest.raises(_lib.SrgptError):
        ops.elem_dtype(torch.float32)


def test_product_path_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "spatialrgpt_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle|oracle[./]", src, flags=re.M), f"{f} references oracle/"
