"""TEST INFRASTRUCTURE (CPU only): how far the REFERENCE's OWN half-precision execution is from its fp32 CPU path.

BASELINE's north_star states rtol 1e-3 / atol 1e-4 for an fp16 run against the fp32 CPU path.  No 16-bit execution of
this network meets that element-wise — including the reference's own: the oracle classes (bit-equal to the reference's
UNet / pipeline, tests/test_oracle_reference_shim.py) converted with `.half()` / `.bfloat16()` exactly as the
reference drivers do (`torch_dtype=torch.float16`, stage2_batchtest_inpaint_model.py:93,123) and run by PyTorch on the
CPU, compared with the same classes in fp32 on identical inputs.  The numbers this script writes
(tests/golden/half_envelope.json) are the yardstick the GPU parity tests hold the CUDA path to: its error against the
fp32 path must not exceed the reference's own 16-bit error (tests/test_unet_pipeline_gpu.py, tests/parity_record.py).

    python tools/make_half_envelope.py            # ~3 min of CPU
"""
from __future__ import annotations

import copy
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

from oracle.factory import make_inputs, make_unet, make_unet_inputs  # noqa: E402
from oracle.pipeline import denoise_loop, prepare_conditioning  # noqa: E402
from oracle.schedulers import OracleDDIMScheduler  # noqa: E402
from oracle.unet import UNetConfig  # noqa: E402
from tests.parity_record import measure  # noqa: E402

OUT = ROOT / "tests" / "golden" / "half_envelope.json"
DTYPES = {"float16": torch.float16, "bfloat16": torch.bfloat16}


@torch.no_grad()
def unet_case(o, i, t):
    kw = {k: i[k] for k in ("class_labels", "my_pose_cond") if k in i}
    ref = o(i["sample"], t, i["encoder_hidden_states"], **kw)[0]
    out = {}
    for name, dt in DTYPES.items():
        oh = copy.deepcopy(o).to(dt)
        got = oh(i["sample"].to(dt), t, i["encoder_hidden_states"].to(dt), **{k: v.to(dt) for k, v in kw.items()})[0]
        out[name] = measure(got, ref)
        del oh
    return out


@torch.no_grad()
def pipeline_case(o, pin, n, steps):
    def run(model, dt):
        cond = prepare_conditioning(s_img_proj_f=pin["s_img_proj_f"], pred_t_img_embed=pin["pred_t_img_embed"],
                                    st_pose_f=pin["st_pose_f"], masked_latents=pin["masked_latents"],
                                    height=pin["height"], width=pin["width"], num_images_per_prompt=n,
                                    guidance_scale=2.0, dtype=dt)
        return denoise_loop(model, OracleDDIMScheduler(), latents=pin["latents"], cond=cond, num_inference_steps=steps,
                            guidance_scale=2.0, dtype=dt)
    ref = run(o, torch.float32)
    out = {}
    for name, dt in DTYPES.items():
        oh = copy.deepcopy(o).to(dt)
        out[name] = measure(run(oh, dt), ref)
        del oh
    return out


def main():
    t0 = time.time()
    doc = {"what": "error of the reference's own 16-bit execution (oracle classes .half()/.bfloat16(), PyTorch CPU) against "
                   "its fp32 CPU path on identical inputs; same metrics as tests/parity_record.py:measure",
           "torch": torch.__version__, "cases": {}}
    tiny = UNetConfig.tiny()
    o = make_unet(tiny, seed=0)
    doc["cases"]["tiny UNet B2 16x32 s_kv9"] = unet_case(o, make_unet_inputs(tiny, batch=2, h=16, w=32, s_kv=9), 981)
    doc["cases"]["cfg1 shapes, tiny weights: n2 16x32 10 DDIM steps"] = pipeline_case(
        o, make_inputs(tiny, n=2, h=16, w=32, s_kv=9), 2, 10)
    print("tiny done", time.time() - t0, flush=True)
    full = UNetConfig.stage2()
    o = make_unet(full, seed=0)
    doc["cases"]["cfg1 UNet eval: 868.9M, B2 32x64 s_kv258"] = unet_case(
        o, make_unet_inputs(full, batch=2, h=32, w=64, s_kv=258), 981)
    print("full UNet done", time.time() - t0, flush=True)
    doc["cases"]["cfg1: 868.9M UNet, n1 (B2) 32x64 s_kv258, 10 DDIM steps, guidance 2"] = pipeline_case(
        o, make_inputs(full, n=1, h=32, w=64, s_kv=258), 1, 10)
    print("full pipeline done", time.time() - t0, flush=True)
    doc["cases"]["cfg2 UNet eval: 868.9M, B16 32x64 s_kv258"] = unet_case(
        o, make_unet_inputs(full, batch=16, h=32, w=64, s_kv=258), 481)
    print("B16 done", time.time() - t0, flush=True)
    OUT.write_text(json.dumps(doc, indent=1) + "\n")
    for case, d in doc["cases"].items():
        for name, m in d.items():
            print(f"{m['pct_outside_rtol1e-3_atol1e-4']:7.2f}%  max {m['max_err_over_max_ref']:.2e}  mean "
                  f"{m['mean_err_over_max_ref']:.2e}  {name:9s} {case}")


if __name__ == "__main__":
    main()
