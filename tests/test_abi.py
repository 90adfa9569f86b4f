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
    assert cdll.pcdm_abi_version() == lib.ABI_VERSION == 3
    assert isinstance(cdll.pcdm_last_error(), bytes)


def test_bad_arguments_return_error_codes_without_a_gpu():
    from pcdms_b200 import lib
    cdll = lib.load()
    rc = cdll.pcdm_gemm(None, 0, None, 0, 0, None, None, 0, None, None, 0, 1, None, 0, 8, 8, 8, 0, 0, 0, None, None)
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


def test_new_entry_points_validate_before_touching_the_device():
    """Argument checks of the entry points added for the stage-1 prior / CLIP encoder run before any CUDA call, so
    they can be exercised here: head widths, null pointers, shapes, dtypes."""
    from pcdms_b200 import lib
    cdll = lib.load()
    p, f, ll, i = ctypes.c_void_p(256), ctypes.c_float, ctypes.c_longlong, ctypes.c_int
    # attention: head_dim must be 64 or 128; strides multiples of 8
    rc = cdll.pcdm_attention_hd(p, ll(640), p, ll(640), p, ll(640), p, ll(640), i(1), i(8), i(6), i(6), i(80), f(0.1),
                                i(0), None)
    assert rc == lib.ERR_UNSUPPORTED and b"head_dim" in cdll.pcdm_last_error()
    rc = cdll.pcdm_attention_hd(None, ll(128), p, ll(128), p, ll(128), p, ll(128), i(1), i(1), i(6), i(6), i(128), f(0.1),
                                i(0), None)
    assert rc == lib.ERR_INVALID
    rc = cdll.pcdm_attention_hd(p, ll(130), p, ll(128), p, ll(128), p, ll(128), i(1), i(1), i(6), i(6), i(128), f(0.1),
                                i(0), None)
    assert rc == lib.ERR_UNSUPPORTED
    # LayerNorm + GEMM: K % 64, N % 32, null gamma
    rc = cdll.pcdm_ln_gemm(p, ll(72), p, p, f(1e-5), None, p, p, ll(64), None, None, ll(0), i(1), None, ll(0), i(4), i(64),
                           i(72), i(0), i(0), None, None)
    assert rc == lib.ERR_UNSUPPORTED
    rc = cdll.pcdm_ln_gemm(p, ll(64), None, p, f(1e-5), None, p, p, ll(64), None, None, ll(0), i(1), None, ll(0), i(4),
                           i(64), i(64), i(0), i(0), None, None)
    assert rc == lib.ERR_INVALID
    # UnCLIP step: null pointers, dtype codes, a noisy step without noise
    row = (ctypes.c_float * 8)(0.5, 0.5, 0.1, 10.0, 1.0, 0.0, 0.0, 0.0)
    assert cdll.pcdm_unclip_step(None, i(2), p, p, p, i(2), row, ll(8), None) == lib.ERR_INVALID
    assert cdll.pcdm_unclip_step(p, i(7), p, p, p, i(2), row, ll(8), None) == lib.ERR_INVALID
    assert cdll.pcdm_unclip_step(p, i(2), p, None, p, i(2), row, ll(8), None) == lib.ERR_INVALID
    assert b"noise" in cdll.pcdm_last_error()
    rc = cdll.pcdm_cfg_unclip_step(p, ll(8), p, p, i(0), ll(4), p, p, p, f(2.0), i(1), i(2), i(8), None, None, None)
    assert rc == lib.ERR_INVALID            # ld_xin < E
    # pcdm_ext is validated before any CUDA call: size, cta group, workspace alignment
    bad = lib.Ext(4, 0, None, 0)   # a caller compiled against a shorter struct
    rc = cdll.pcdm_gemm(p, ll(64), None, ll(0), i(0), p, p, ll(64), None, None, ll(0), i(1), None, ll(0), i(128), i(64),
                        i(64), i(0), i(0), i(0), ctypes.byref(bad), None)
    assert rc == lib.ERR_INVALID and b"pcdm_ext" in cdll.pcdm_last_error()
    bad = lib.Ext(ctypes.sizeof(lib.Ext), 3, None, 0)
    rc = cdll.pcdm_conv3x3(p, p, p, None, None, ll(0), None, i(1), i(8), i(16), i(64), i(64), i(1), i(0), i(0), i(0),
                           ctypes.byref(bad), None)
    assert rc == lib.ERR_INVALID
    bad = lib.Ext(ctypes.sizeof(lib.Ext), 0, 8, 1024)
    rc = cdll.pcdm_gemm(p, ll(64), None, ll(0), i(0), p, p, ll(64), None, None, ll(0), i(1), None, ll(0), i(128), i(64),
                        i(64), i(0), i(0), i(0), ctypes.byref(bad), None)
    assert rc == lib.ERR_INVALID and b"aligned" in cdll.pcdm_last_error()
    assert cdll.pcdm_gemm_workspace_bytes(i(512), i(1280)) >= 2 * 512 * 1280 * 4


def test_release_library_has_no_mutable_global_hooks():
    """The release .so exports none of the process-wide tuning / experiment setters (they live in the experiment
    build, include/pcdm_b200_experiment.h); `nm -D` is the check VERDICT r1 asked for."""
    import re
    import subprocess
    from pcdms_b200 import build, lib
    path = build.build()
    out = subprocess.run(["nm", "-D", "--defined-only", str(path)], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r"\bT (pcdm_[a-z0-9_]+)", out))
    assert "pcdm_gemm" in exported and "pcdm_conv3x3" in exported
    setters = sorted(s for s in exported if s.startswith("pcdm_set_"))
    assert setters == [], setters
    hdr = (lib.HEADER_PATH.parent / "pcdm_b200_experiment.h").read_text()
    assert "pcdm_set_gemm_debug" in hdr and "pcdm_set_gemm_debug" not in lib.HEADER_PATH.read_text()
