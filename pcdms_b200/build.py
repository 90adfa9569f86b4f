"""Build libpcdm_b200.so (hand-written sm_100a kernels + C ABI) in-tree with nvcc.

The .so is git-ignored but travels with the repo snapshot to the GPU box.  nvcc cross-compiles sm_100a without a
GPU, so this also runs in the CPU-only build container (`__graft_entry__.build()`).

Two variants of the same sources:
  * libpcdm_b200.so      — the release library: no mutable process-wide state, no experiment hooks;
  * libpcdm_b200_exp.so  — `-DPCDM_EXPERIMENT`: adds the pcdm_set_* tuning / experiment hooks of
                           include/pcdm_b200_experiment.h.  Built on demand (`--experiment`), used by tools/ only
                           (through $PCDM_B200_LIB); nothing in pcdms_b200/ needs it.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
LIB_PATH = PKG_DIR / "libpcdm_b200.so"
EXP_LIB_PATH = PKG_DIR / "libpcdm_b200_exp.so"
OBJ_DIR = PKG_DIR / "build"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "--use_fast_math",
    "-Xcompiler", "-fPIC",
]
LINK_FLAGS = ["-shared", "-cudart", "static"]


def _sources() -> list[Path]:
    return sorted(CSRC.glob("*.cu"))


def _digest(extra: str = "") -> str:
    h = hashlib.sha256()
    for p in sorted(list(CSRC.glob("*")) + sorted((PKG_DIR.parent / "include").glob("*.h"))):
        if p.is_file():
            h.update(p.name.encode())
            h.update(p.read_bytes())
    h.update(" ".join(NVCC_FLAGS + LINK_FLAGS).encode())
    h.update(extra.encode())
    return h.hexdigest()


def source_digest() -> str:
    """Digest of everything the release library is built from (lib.load() compares it with the stamp)."""
    return _digest()


def _stamp(lib: Path) -> Path:
    return lib.with_name("." + lib.stem + ".stamp")


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def build(force: bool = False, verbose: bool = False, experiment: bool = False) -> Path:
    """Compile every .cu under csrc/ (one nvcc per source, in parallel) and link one shared library.  No-op when the
    sources are unchanged."""
    lib = EXP_LIB_PATH if experiment else LIB_PATH
    defs = ["-DPCDM_EXPERIMENT"] if experiment else []
    digest = _digest("exp" if experiment else "")
    stamp = _stamp(lib)
    if not force and lib.exists() and stamp.exists() and stamp.read_text().strip() == digest:
        return lib
    odir = OBJ_DIR / ("exp" if experiment else "rel")
    odir.mkdir(parents=True, exist_ok=True)
    nvcc = nvcc_path()
    inc = ["-I", str(PKG_DIR.parent / "include")]

    def compile_one(src: Path):
        obj = odir / (src.stem + ".o")
        cmd = [nvcc, *NVCC_FLAGS, *defs, *inc, "-c", "-o", str(obj), str(src)]
        if verbose:
            cmd += ["-Xptxas", "-v"]
        return obj, subprocess.run(cmd, capture_output=True, text=True)

    with ThreadPoolExecutor(max_workers=min(8, len(_sources()))) as ex:
        results = list(ex.map(compile_one, _sources()))
    for obj, proc in results:
        if proc.returncode != 0:
            sys.stderr.write(proc.stdout + proc.stderr)
            raise RuntimeError(f"nvcc failed compiling {obj.stem}.cu (exit {proc.returncode})")
        if verbose:
            sys.stderr.write(proc.stderr)
    proc = subprocess.run([nvcc, *NVCC_FLAGS, *LINK_FLAGS, "-o", str(lib), *[str(o) for o, _ in results]],
                          capture_output=True, text=True)
    if proc.returncode != 0:
        sys.stderr.write(proc.stdout + proc.stderr)
        raise RuntimeError(f"nvcc failed linking {lib.name} (exit {proc.returncode})")
    stamp.write_text(digest)
    return lib


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="-v" in sys.argv, experiment="--experiment" in sys.argv)
    print(path)
