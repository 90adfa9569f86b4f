"""Build libpcdm_b200.so (hand-written sm_100a kernels + C ABI) in-tree with nvcc.

The .so is git-ignored but travels with the repo snapshot to the GPU box.  nvcc cross-compiles sm_100a without a
GPU, so this also runs in the CPU-only build container (`__graft_entry__.build()`).
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
LIB_PATH = PKG_DIR / "libpcdm_b200.so"
STAMP = PKG_DIR / ".libpcdm_b200.stamp"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "--use_fast_math",
    "-Xcompiler", "-fPIC",
    "-shared",
    "-cudart", "static",
]


def _sources() -> list[Path]:
    return sorted(CSRC.glob("*.cu"))


def _digest() -> str:
    h = hashlib.sha256()
    for p in sorted(list(CSRC.glob("*")) + [PKG_DIR.parent / "include" / "pcdm_b200.h"]):
        if p.is_file():
            h.update(p.name.encode())
            h.update(p.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile every .cu under csrc/ into one shared library. No-op when sources are unchanged."""
    digest = _digest()
    if not force and LIB_PATH.exists() and STAMP.exists() and STAMP.read_text().strip() == digest:
        return LIB_PATH
    cmd = [nvcc_path(), *NVCC_FLAGS, "-I", str(PKG_DIR.parent / "include"), "-o", str(LIB_PATH)]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += [str(s) for s in _sources()]
    # one nvcc invocation per source in parallel would be faster; a single call keeps the recipe obvious
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        sys.stderr.write(proc.stdout + proc.stderr)
        raise RuntimeError(f"nvcc failed building {LIB_PATH.name} (exit {proc.returncode})")
    if verbose:
        sys.stderr.write(proc.stderr)
    STAMP.write_text(digest)
    return LIB_PATH


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(path)
