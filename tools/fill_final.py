"""Fill the FINAL_* placeholders of DESIGN.md / README.md from the end-of-round verification run
(gpurun_out/r2_final_*.json, written by tools/final_run.sh) and copy the summaries that are judged to profiles/."""
import json
import re
import shutil
import sys
from pathlib import Path

root = Path(__file__).resolve().parent.parent
O = root / "gpurun_out"
b = json.loads((O / "r2_final_bench.json").read_text())
ref = json.loads((O / "r2_final_bench_reference_arm.json").read_text())
ab = json.loads((O / "r2_final_ablate.json").read_text())
tests = (O / "r2_final_tests.log").read_text()
npass = re.search(r"(\d+) passed", tests).group(1)
full = ab["full"]
d = lambda k: f"{full - ab[k]:.2f}"
vals = {
    "FINAL_IMGS": f"{b['value']:.2f}", "FINAL_E2E": f"{b['e2e']['value']:.2f}", "FINAL_STEP": f"{b['unet_step_ms']:.2f}",
    "FINAL_CONV": f"{b['roofline']['achieved']:.0f}", "FINAL_FRAC": f"{b['roofline']['frac']:.2f}",
    "FINAL_WHOLE": f"{b['roofline']['whole_unet_step']['frac_of_sustained_peak']:.2f}",
    "FINAL_IIO": f"{b['image_in_image_out']['value']:.1f}", "FINAL_CPU": f"{ref['value']:.4f}", "FINAL_TESTS": npass,
    "FINAL_AB_FULL": f"{full:.2f}", "FINAL_AB_CONV": d("no conv3x3"), "FINAL_AB_GEMM": d("no gemm"),
    "FINAL_AB_GEGLU": d("no GEGLU gemm"), "FINAL_AB_SHORTK": d("no gemm with K <= 640 (non-GEGLU)"),
    "FINAL_AB_ATT": d("no attention"), "FINAL_AB_GN": d("no groupnorm"),
}
for name in ("DESIGN.md", "README.md"):
    p = root / name
    s = p.read_text()
    for k in sorted(vals, key=len, reverse=True):
        s = s.replace(k, vals[k])
    p.write_text(s)
for f in ("bench.json", "bench_reference_arm.json", "ablate.json", "bench_configs.json", "launches.csv", "launches.md",
          "tests.log", "parity.txt", "smoke.log"):
    if (O / f"r2_final_{f}").exists():
        shutil.copy(O / f"r2_final_{f}", root / "profiles" / f"r2_final_{f}")
if (O / "r2_final_parity.json").exists():
    shutil.copy(O / "r2_final_parity.json", root / "profiles" / "r2_parity.json")
if (O / "r2_final_roofline_traffic.json").exists():
    shutil.copy(O / "r2_final_roofline_traffic.json", root / "profiles" / "roofline_traffic.json")
print(json.dumps(vals, indent=1))

# ncu --set full summary table and the nvidia-smi clock record of the bench run
import csv
import statistics
import subprocess

raw = O / "r2_final_kernels_raw.csv"
if raw.exists():
    md = subprocess.run([sys.executable, str(root / "tools" / "summarize_ncu_raw.py"), str(raw),
                         "Round 2 — ncu --set full of the round's kernels"], capture_output=True, text=True).stdout
    if md.strip():
        (root / "profiles" / "r2_final_ncu_kernels.md").write_text(md)
clk = O / "r2_final_clocks.csv"
if clk.exists():
    rows = list(csv.reader(clk.open()))[1:]
    sm = [int(r[1].split()[0]) for r in rows if len(r) > 8 and r[1].strip().split()[0].isdigit()]
    pw = [float(r[3].split()[0]) for r in rows if len(r) > 8 and r[3].strip().split()[0].replace(".", "").isdigit()]
    reasons = {}
    for r in rows:
        if len(r) > 8:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.strip() == "Active":
                    reasons[name] = reasons.get(name, 0) + 1
    (root / "profiles" / "r2_final_clocks.txt").write_text(
        f"nvidia-smi -lms 200 during bench.py --steps 5 --warmup 3 (gpurun_out/r2_final_clocks.csv): {len(sm)} samples, "
        f"SM clock median {statistics.median(sm):.0f} MHz (min {min(sm)}, max {max(sm)}; clocks.max.sm "
        f"{rows[0][2].strip() if rows else '?'}), power max {max(pw):.0f} W; samples with reason active: {reasons}\n")
