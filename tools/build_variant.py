"""Build another variant of libpcdm_b200.so with extra nvcc defines, for same-box A/B through $PCDM_B200_LIB:
    python tools/build_variant.py .ab/libold.so -DPCDM_GELU_AS
(.ab/ is git-ignored but travels to the GPU box)."""
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from pcdms_b200 import build as B

out = Path(sys.argv[1]).resolve()
defs = sys.argv[2:]
odir = out.parent / (out.stem + "_obj")
odir.mkdir(parents=True, exist_ok=True)
nvcc = B.nvcc_path()
inc = ["-I", str(B.PKG_DIR.parent / "include")]


def one(src):
    obj = odir / (src.stem + ".o")
    r = subprocess.run([nvcc, *B.NVCC_FLAGS, *defs, *inc, "-c", "-o", str(obj), str(src)], capture_output=True, text=True)
    if r.returncode:
        sys.exit(r.stdout + r.stderr)
    return obj


with ThreadPoolExecutor(8) as ex:
    objs = list(ex.map(one, B._sources()))
r = subprocess.run([nvcc, *B.NVCC_FLAGS, *B.LINK_FLAGS, "-o", str(out), *map(str, objs)], capture_output=True, text=True)
if r.returncode:
    sys.exit(r.stdout + r.stderr)
print(out)
