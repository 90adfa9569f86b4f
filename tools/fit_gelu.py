"""Coefficients of `gelu2_phi<DEG>` (pcdms_b200/csrc/common.cuh): polynomial fit of P(t) = log2 Phi(-t), t = |x|, used as
    gelu(x) = max(x, 0) - t * 2^P(t)
Weighted least squares with iterative re-weighting towards the minimax solution; the weight t * Phi(-t) makes the
fitted quantity the ABSOLUTE error of the GELU value.  The check evaluates the fp32 Horner chain the kernel runs and
compares with the exact erf form in fp64 over [-40, 40] plus a few huge arguments (no NaN, correct saturation).
usage: python tools/fit_gelu.py            (CPU only)"""
import numpy as np
from scipy.special import erf, log_ndtr, ndtr


def gelu(x):
    return 0.5 * x * (1 + erf(x / np.sqrt(2)))


def eval32(c, xs):
    t = np.abs(xs).astype(np.float32)
    r = np.full_like(t, np.float32(c[-1]))
    with np.errstate(over="ignore", under="ignore"):
        for k in range(len(c) - 2, -1, -1):
            r = (r * t + np.float32(c[k])).astype(np.float32)
        e = np.exp2(r.astype(np.float64)).astype(np.float32)
        return (np.maximum(xs, np.float32(0)).astype(np.float64) - t.astype(np.float64) * e).astype(np.float32).astype(np.float64)


def fit(deg, T):
    t = np.linspace(0, T, 40001)
    y = log_ndtr(-t) / np.log(2)
    wgt = t * ndtr(-t) + 1e-9
    w, best = np.ones_like(t), None
    V = np.vander(t / T, deg + 1, increasing=True)
    for _ in range(300):
        coef, *_ = np.linalg.lstsq(V * (w * wgt)[:, None], y * w * wgt, rcond=None)
        err = np.abs((V @ coef - y) * wgt)
        if best is None or err.max() < best[0]:
            best = (err.max(), coef.copy())
        w = w * (1 + 2 * err / err.max())
        w /= w.mean()
    return best[1] / T ** np.arange(deg + 1)


if __name__ == "__main__":
    xs = np.concatenate([np.linspace(-40, 40, 2000001), np.array([-1e3, 1e3, -1e5, 1e5, -1e10, 1e10, 0.0])]).astype(np.float32)
    ref = gelu(xs.astype(np.float64))
    for deg, T in ((5, 5.5), (8, 6.5)):
        c = fit(deg, T)
        out = eval32(c, xs)
        ae = np.abs(out - ref)
        m = np.abs(ref) > 1e-4
        print(f"DEG {deg}: max |err| {ae.max():.3e} at x = {xs[ae.argmax()]:.3f}; max relative error where |gelu| > 1e-4: "
              f"{(ae[m] / np.abs(ref[m])).max():.3e}; NaNs {int(np.isnan(out).sum())}; leading coefficient {c[-1]:.3e}")
        print("   low -> high:", ", ".join("%.9ef" % np.float32(v) for v in c))
