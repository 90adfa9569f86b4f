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
    global _lib
    if _lib is not None:
        return _lib
    path = LIB_PATH
    if os.environ.get("PCDM_B200_LIB"):   # A/B hook for tools/: another build of the same library (never a fallback)
        path = Path(os.environ["PCDM_B200_LIB"])
    elif not LIB_PATH.exists():
        if not build_if_missing:
            raise FileNotFoundError(f"{LIB_PATH} not built; run python -m pcdms_b200.build")
        from . import build as _build

        _build.build()
    lib = C.CDLL(str(path))
    lib.pcdm_last_error.restype = C.c_char_p
    lib.pcdm_abi_version.restype = C.c_int
    lib.pcdm_groupnorm_workspace_bytes.restype = C.c_longlong
    mode = os.environ.get("PCDM_GEMM_CTA_GROUP")  # tuning hook: 1 = single-CTA tiles, 2 = CTA pairs, unset = auto
    if mode:
        lib.pcdm_set_gemm_cta_group(C.c_int(int(mode)))
    if os.environ.get("PCDM_SKINNY") is not None:      # A/B hooks: 0 = M <= 32 GEMMs / short attention on the tcgen05 kernels
        lib.pcdm_set_skinny_gemm(C.c_int(int(os.environ["PCDM_SKINNY"])))
    if os.environ.get("PCDM_ATT_SMALL") is not None:
        lib.pcdm_set_attention_small(C.c_int(int(os.environ["PCDM_ATT_SMALL"])))
    if os.environ.get("PCDM_PDL") is not None:   # tuning hook: 0 disables programmatic dependent launch
        lib.pcdm_set_pdl(C.c_int(int(os.environ["PCDM_PDL"])))
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
