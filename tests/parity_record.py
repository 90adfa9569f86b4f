"""TEST INFRASTRUCTURE: one place where every whole-path parity comparison (CUDA path vs CPU oracle / reference golden)
is measured, asserted and RECORDED.

BASELINE's north_star states the tolerance `rtol=1e-3 / atol=1e-4` (fp16, against the fp32 CPU path).  Each kernel
meets it on identical inputs (tests/test_kernels_gpu.py).  For whole-UNet / whole-pipeline comparisons the share of
elements outside that tolerance is a measured quantity, not an argument: `check()` computes it and appends one JSON
record per comparison to `gpurun_out/r2_parity.jsonl` (override: $PCDM_PARITY_LOG) so the numbers survive `pytest -q`;
`tools/collect_parity.py` turns the log into the committed `profiles/r2_parity.json`.

The yardstick for that share: no 16-bit execution of this network meets the tolerance element-wise, the reference's own
included.  `tests/golden/half_envelope.json` (tools/make_half_envelope.py, CPU) holds the error of the REFERENCE's own
fp16 / bf16 execution (the oracle classes — bit-equal to the reference's — under `.half()` / `.bfloat16()`, PyTorch
CPU) against its fp32 path for the same inputs; wherever a comparison's `config` has such an entry, `check()` asserts
that the CUDA path is AT LEAST AS CLOSE to the fp32 CPU path as the reference's own 16-bit run (mean error, share outside
the tolerance; max error within 1.25x) and records both sides.
"""
from __future__ import annotations

import json
import os
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
RTOL, ATOL = 1e-3, 1e-4   # north_star


_ENVELOPE = None


def reference_half_envelope(config: str, dtype: str):
    """The reference's own 16-bit error for this comparison (None when the fixture has no such case)."""
    global _ENVELOPE
    if _ENVELOPE is None:
        f = ROOT / "tests" / "golden" / "half_envelope.json"
        _ENVELOPE = json.loads(f.read_text())["cases"] if f.exists() else {}
    return _ENVELOPE.get(config, {}).get(dtype)


def _log_path() -> Path:
    p = os.environ.get("PCDM_PARITY_LOG")
    return Path(p) if p else ROOT / "gpurun_out" / "r2_parity.jsonl"


def measure(got: torch.Tensor, want: torch.Tensor) -> dict:
    got, want = got.detach().float().cpu(), want.detach().float().cpu()
    err = (got - want).abs()
    scale = want.abs().max().item()
    strict = err > ATOL + RTOL * want.abs()
    return {
        "elements": int(want.numel()),
        "max_abs_ref": scale,
        "max_abs_err": err.max().item(),
        "mean_abs_err": err.mean().item(),
        "max_err_over_max_ref": err.max().item() / scale if scale else float("nan"),
        "mean_err_over_max_ref": err.mean().item() / scale if scale else float("nan"),
        "pct_outside_rtol1e-3_atol1e-4": 100.0 * strict.float().mean().item(),
        "nan": bool(torch.isnan(got).any()),
    }


def check(got, want, max_frac, mean_frac, label, *, config, dtype, against="oracle", extra=None):
    """Assert the envelope (max / mean |err| as fractions of max|ref|) and record the strict-tolerance share."""
    m = measure(got, want)
    rec = {"label": label, "config": config, "dtype": str(dtype).replace("torch.", ""), "against": against,
           "asserted_max_frac": max_frac, "asserted_mean_frac": mean_frac, **m, "when": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime())}
    if extra:
        rec.update(extra)
    env = reference_half_envelope(config, rec["dtype"]) if against == "oracle" else None
    if env:
        rec["reference_own_16bit_run"] = {k: env[k] for k in ("max_abs_err", "mean_abs_err", "max_err_over_max_ref",
                                                              "mean_err_over_max_ref", "pct_outside_rtol1e-3_atol1e-4")}
        rec["closer_to_fp32_than_reference_16bit"] = bool(
            m["mean_abs_err"] <= env["mean_abs_err"] and m["max_abs_err"] <= 1.25 * env["max_abs_err"] and
            m["pct_outside_rtol1e-3_atol1e-4"] <= env["pct_outside_rtol1e-3_atol1e-4"])
    try:
        p = _log_path()
        p.parent.mkdir(parents=True, exist_ok=True)
        with open(p, "a") as f:
            f.write(json.dumps(rec) + "\n")
    except OSError:
        pass
    print(f"[parity] {label}: max|err| {m['max_abs_err']:.3e} ({m['max_err_over_max_ref']:.2e} of max|ref|), "
          f"mean|err| {m['mean_abs_err']:.3e}, outside rtol 1e-3/atol 1e-4: {m['pct_outside_rtol1e-3_atol1e-4']:.2f}%")
    assert not m["nan"], label
    assert m["max_abs_err"] <= max_frac * m["max_abs_ref"], (label, m)
    assert m["mean_abs_err"] <= mean_frac * m["max_abs_ref"], (label, m)
    if env:
        print(f"[parity] {label}: the reference's own {rec['dtype']} run: max|err| {env['max_abs_err']:.3e}, mean|err| "
              f"{env['mean_abs_err']:.3e}, outside: {env['pct_outside_rtol1e-3_atol1e-4']:.2f}%")
        assert rec["closer_to_fp32_than_reference_16bit"], (label, m, env)
    return m
