"""ctypes binding of libpcdm_b200.so (the C ABI declared in include/pcdm_b200.h).

There is deliberately no CPU or PyTorch fallback here: if the shared library is missing or a kernel call fails the
caller gets an exception.  `load()` only dlopens the library (works without a GPU, used by the CPU test-suite to
check that every symbol declared in the header is exported); compute calls need a CUDA device.
"""
from __future__ import annotations

import ctypes as C
import os
import re
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
LIB_PATH = PKG_DIR / "libpcdm_b200.so"
HEADER_PATH = PKG_DIR.parent / "include" / "pcdm_b200.h"

ERR_INVALID, ERR_CUDA, ERR_UNSUPPORTED = -1, -2, -3
DT_F16, DT_BF16 = 0, 1
FLAG_GEGLU, FLAG_OUT_F32, FLAG_SILU, FLAG_GELU, FLAG_PAD_BR, FLAG_W_STATIC = 1, 2, 4, 8, 16, 32
FLAG_NO_SKINNY = 64                           # pcdm_gemm
FLAG_GN_TWO_PASS, FLAG_GN_ONE_PASS = 64, 128  # pcdm_groupnorm
ABI_VERSION = 3


class Ext(C.Structure):
    """`pcdm_ext` of include/pcdm_b200.h: optional per-call extras of pcdm_gemm / pcdm_conv3x3 / pcdm_ln_gemm."""
    _fields_ = [("size", C.c_int), ("force_cta_group", C.c_int), ("workspace", C.c_void_p),
                ("workspace_bytes", C.c_longlong), ("row_stats", C.c_void_p), ("row_stats_cap", C.c_int),
                ("row_stats_parts", C.c_int), ("ln_stats", C.c_void_p), ("ln_parts", C.c_int),
                ("ln_eps", C.c_float), ("chan_stats", C.c_void_p)]


class PcdmError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"pcdm_b200 error {code}: {msg}")
        self.code = code


_lib = None
launch_count = 0  # kernels launched through the C ABI by this process (bench.py reports it as gpu_launches)


def count(n: int = 1) -> None:
    global launch_count
    launch_count += n


def declared_symbols() -> list[str]:
    """Every function name declared in include/pcdm_b200.h."""
    text = HEADER_PATH.read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pcdm_[a-z0-9_]+)\s*\(", text)))


def load(build_if_missing: bool = True) -> C.CDLL:
    """dlopen the library.  The in-tree build is brought up to date first (`build.build()` is a no-op when the digest
    of csrc/ + include/ matches the stamp), so a stale .so from an older ABI is never loaded silently; where no
    compiler is available (a box that only received the prebuilt .so) the stamp must match the sources."""
    global _lib
    if _lib is not None:
        return _lib
    path = LIB_PATH
    if os.environ.get("PCDM_B200_LIB"):   # A/B hook for tools/: another build of the same library (never a fallback)
        path = Path(os.environ["PCDM_B200_LIB"])
    else:
        from . import build as _build
        stamp = _build._stamp(LIB_PATH)
        fresh = LIB_PATH.exists() and stamp.exists() and stamp.read_text().strip() == _build.source_digest()
        if not fresh:
            if not build_if_missing:
                raise FileNotFoundError(f"{LIB_PATH} is missing or stale; run python -m pcdms_b200.build")
            _build.build()
    lib = C.CDLL(str(path))
    lib.pcdm_last_error.restype = C.c_char_p
    lib.pcdm_abi_version.restype = C.c_int
    if lib.pcdm_abi_version() != ABI_VERSION:
        raise RuntimeError(f"{path} has ABI version {lib.pcdm_abi_version()}, this package binds version {ABI_VERSION}")
    lib.pcdm_groupnorm_workspace_bytes.restype = C.c_longlong
    lib.pcdm_gemm_workspace_bytes.restype = C.c_longlong
    # experiment build only (libpcdm_b200_exp.so through $PCDM_B200_LIB): environment-driven tuning hooks for tools/
    for env, fn in (("PCDM_GEMM_CTA_GROUP", "pcdm_set_gemm_cta_group"), ("PCDM_SKINNY", "pcdm_set_skinny_gemm"),
                    ("PCDM_ATT_SMALL", "pcdm_set_attention_small"), ("PCDM_PDL", "pcdm_set_pdl")):
        if os.environ.get(env) is not None:
            if not hasattr(lib, fn):
                raise RuntimeError(f"${env} needs the experiment build (python -m pcdms_b200.build --experiment, "
                                   f"PCDM_B200_LIB=pcdms_b200/libpcdm_b200_exp.so)")
            getattr(lib, fn)(C.c_int(int(os.environ[env])))
    _lib = lib
    return lib


def check(rc: int, kernels: int = 1) -> None:
    global launch_count
    launch_count += kernels
    if rc != 0:
        raise PcdmError(rc, load().pcdm_last_error().decode(errors="replace"))


def ptr(t) -> C.c_void_p:
    """Device pointer of a torch tensor (or NULL for None)."""
    if t is None:
        return C.c_void_p(0)
    return C.c_void_p(t.data_ptr())
