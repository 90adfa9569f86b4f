"""The one-SFU-op GELU of the GEMM epilogues (`gelu2_phi<DEG>`, pcdms_b200/csrc/common.cuh) replaces torch's exact-erf
GELU (diffusers GEGLU: `hidden * F.gelu(gate)`).  This CPU test reads the polynomial coefficients
out of the CUDA source, runs the same fp32 Horner chain in numpy and checks the absolute error against the exact form —
so a typo in the kernel's constants fails here, without a GPU.  The GPU parity tests (`test_gemm_geglu*`) check the kernel
itself."""
import re
from pathlib import Path

import numpy as np
import pytest

scipy_special = pytest.importorskip("scipy.special")

SRC = Path(__file__).resolve().parent.parent / "pcdms_b200" / "csrc" / "common.cuh"


def _coefficients():
    text = SRC.read_text()
    body = text[text.index("float2 gelu2_phi(float2 x)"):text.index("#undef PCDM_C2")]
    lo, hi = body.split("} else {")
    num = r"PCDM_C2\((-?[0-9.]+e[+-][0-9]+)f\)"
    return {5: [float(v) for v in re.findall(num, lo)], 8: [float(v) for v in re.findall(num, hi)]}


def _gelu_kernel_math(x32, coef_high_to_low):
    t = np.abs(x32)
    p = np.full_like(t, np.float32(coef_high_to_low[0]))
    with np.errstate(over="ignore", under="ignore", invalid="ignore"):
        for c in coef_high_to_low[1:]:
            p = (p * t + np.float32(c)).astype(np.float32)
        e = np.exp2(p.astype(np.float64)).astype(np.float32)
        return (np.maximum(x32, np.float32(0)).astype(np.float64) - t.astype(np.float64) * e).astype(np.float32)


@pytest.mark.parametrize("deg,bound", [(5, 8e-7), (8, 4e-7)])
def test_gelu_polynomial_in_the_cuda_source_matches_exact_gelu(deg, bound):
    coef = _coefficients()[deg]
    assert len(coef) == deg + 1 and coef[0] < 0          # negative leading coefficient: the correction vanishes for large |x|
    x = np.concatenate([np.linspace(-40, 40, 400001), [-1e3, 1e3, -1e6, 1e6, 0.0]]).astype(np.float32)
    got = _gelu_kernel_math(x, coef).astype(np.float64)
    xd = x.astype(np.float64)
    want = 0.5 * xd * (1.0 + scipy_special.erf(xd / np.sqrt(2.0)))
    assert not np.isnan(got).any()
    err = np.abs(got - want)
    assert err.max() < bound, (err.max(), x[err.argmax()])
    # relative to the result wherever it is not negligible: far below a 16-bit ulp (fp16 4.9e-4 for DEG 8, bf16 3.9e-3 for 5)
    m = np.abs(want) > 1e-3
    assert (err[m] / np.abs(want[m])).max() < (1e-3 if deg == 5 else 5e-5)
