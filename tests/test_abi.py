"""The C-ABI shared library builds, loads without a GPU and exports every symbol include/pcdm_b200.h declares; the
product has no CPU path (compute calls on CPU tensors raise)."""
import ctypes

import pytest
import torch


def test_library_builds_loads_and_exports_header_symbols():
    from pcdms_b200 import build, lib
    path = build.build()
    assert path.exists()
    cdll = lib.load()
    syms = lib.declared_symbols()
    assert len(syms) >= 15 and "pcdm_conv3x3" in syms and "pcdm_attention" in syms and "pcdm_cfg_ddim_step" in syms
    for s in syms:
        assert hasattr(cdll, s), f"{s} declared in include/pcdm_b200.h but not exported"
    assert cdll.pcdm_abi_version() == 1
    assert isinstance(cdll.pcdm_last_error(), bytes)


def test_bad_arguments_return_error_codes_without_a_gpu():
    from pcdms_b200 import lib
    cdll = lib.load()
    rc = cdll.pcdm_gemm(None, 0, None, 0, 0, None, None, 0, None, None, 0, 1, None, 0, 8, 8, 8, 0, 0, 0, None)
    assert rc == lib.ERR_INVALID and b"null" in cdll.pcdm_last_error()
    rc = cdll.pcdm_layernorm(ctypes.c_void_p(16), 8, ctypes.c_void_p(16), 8, ctypes.c_void_p(16), ctypes.c_void_p(16),
                             ctypes.c_float(1e-5), 4, 12, 0, None)
    assert rc == lib.ERR_UNSUPPORTED


def test_ops_refuse_cpu_tensors():
    from pcdms_b200 import ops
    with pytest.raises(RuntimeError, match="no CPU path"):
        ops.gemm(torch.zeros(4, 64, dtype=torch.float16), torch.zeros(32, 64, dtype=torch.float16))
    with pytest.raises(TypeError):
        ops.gemm(torch.zeros(4, 64), torch.zeros(32, 64))


def test_product_does_not_import_oracle():
    import pathlib
    import re
    root = pathlib.Path(__file__).resolve().parent.parent / "pcdms_b200"
    for p in root.rglob("*.py"):
        assert not re.search(r"^\s*(from|import)\s+oracle\b", p.read_text(), flags=re.M), p
