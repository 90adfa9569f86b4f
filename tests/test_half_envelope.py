"""CPU: the yardstick of the whole-path parity record — how far the REFERENCE's own fp16 / bf16 execution is from its
fp32 CPU path (tests/golden/half_envelope.json, tools/make_half_envelope.py) — exists for every BASELINE comparison the
GPU tests hold the CUDA path to, is reproducible here, and shows that BASELINE's rtol 1e-3 / atol 1e-4 is not met
element-wise by ANY 16-bit execution of this network (which is why the GPU tests assert 'at least as close to the fp32
path as the reference's own 16-bit run' and record the share outside the tolerance instead of asserting it is zero)."""
import copy
import json
from pathlib import Path

import torch

FIXTURE = Path(__file__).parent / "golden" / "half_envelope.json"
CASES = ["tiny UNet B2 16x32 s_kv9", "cfg1 shapes, tiny weights: n2 16x32 10 DDIM steps",
         "cfg1 UNet eval: 868.9M, B2 32x64 s_kv258", "cfg1: 868.9M UNet, n1 (B2) 32x64 s_kv258, 10 DDIM steps, guidance 2",
         "cfg2 UNet eval: 868.9M, B16 32x64 s_kv258"]


def test_fixture_covers_the_baseline_comparisons():
    doc = json.loads(FIXTURE.read_text())
    for case in CASES:
        for dt in ("float16", "bfloat16"):
            m = doc["cases"][case][dt]
            assert not m["nan"] and m["elements"] > 0
            # no 16-bit run of the reference meets the element-wise tolerance: more than half of the elements are outside
            assert m["pct_outside_rtol1e-3_atol1e-4"] > 50.0, (case, dt, m)
            assert m["max_err_over_max_ref"] < (5e-3 if dt == "float16" else 4e-2)


def test_parity_record_looks_the_envelope_up_by_config_and_dtype():
    from tests.parity_record import reference_half_envelope
    assert reference_half_envelope(CASES[2], "float16")["elements"] == 2 * 4 * 32 * 64
    assert reference_half_envelope(CASES[2], "float32") is None
    assert reference_half_envelope("no such comparison", "float16") is None


@torch.no_grad()
def test_tiny_case_is_reproducible_here():
    """The tiny-UNet entry recomputed on this machine (PyTorch CPU half kernels may differ in summation order between
    machines, hence a band instead of equality)."""
    from oracle.factory import make_unet, make_unet_inputs
    from oracle.unet import UNetConfig
    from tests.parity_record import measure
    cfg = UNetConfig.tiny()
    o = make_unet(cfg, seed=0)
    i = make_unet_inputs(cfg, batch=2, h=16, w=32, s_kv=9)
    kw = {k: i[k] for k in ("class_labels", "my_pose_cond")}
    ref = o(i["sample"], 981, i["encoder_hidden_states"], **kw)[0]
    want = json.loads(FIXTURE.read_text())["cases"][CASES[0]]
    for name, dt in (("float16", torch.float16), ("bfloat16", torch.bfloat16)):
        oh = copy.deepcopy(o).to(dt)
        got = oh(i["sample"].to(dt), 981, i["encoder_hidden_states"].to(dt), **{k: v.to(dt) for k, v in kw.items()})[0]
        m = measure(got, ref)
        assert 0.6 * want[name]["mean_abs_err"] < m["mean_abs_err"] < 1.6 * want[name]["mean_abs_err"], (name, m, want[name])
