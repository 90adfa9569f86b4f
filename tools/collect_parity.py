"""gpurun_out/r2_parity.jsonl (written by the GPU parity tests, tests/parity_record.py) -> profiles/r2_parity.json.

    python tools/collect_parity.py [log.jsonl] [out.json]

One record per comparison label (the latest run wins), plus a summary of the north-star criterion
(rtol 1e-3 / atol 1e-4 against the fp32 CPU path) per BASELINE configuration.
"""
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def main():
    log = Path(sys.argv[1]) if len(sys.argv) > 1 else ROOT / "gpurun_out" / "r2_parity.jsonl"
    out = Path(sys.argv[2]) if len(sys.argv) > 2 else ROOT / "profiles" / "r2_parity.json"
    recs = {}
    for line in log.read_text().splitlines():
        if line.strip():
            r = json.loads(line)
            recs[r["label"]] = r
    rows = sorted(recs.values(), key=lambda r: (r["config"], r["dtype"], r["label"]))
    try:
        head = subprocess.run(["git", "rev-parse", "--short", "HEAD"], cwd=ROOT, capture_output=True, text=True).stdout.strip()
    except OSError:
        head = None
    doc = {
        "tolerance": {"rtol": 1e-3, "atol": 1e-4, "source": "BASELINE.json north_star (fp16 vs the fp32 CPU path)"},
        "note": "pct_outside = share of output elements with |got - ref| > atol + rtol * |ref|; asserted_* are the "
                "envelopes the tests enforce (fractions of max|ref|)",
        "git_head_when_collected": head,
        "records": rows,
    }
    out.write_text(json.dumps(doc, indent=1) + "\n")
    for r in rows:
        print(f"{r['pct_outside_rtol1e-3_atol1e-4']:7.2f}%  max {r['max_err_over_max_ref']:.2e}  mean {r['mean_err_over_max_ref']:.2e}  "
              f"{r['dtype']:9s} {r['label']}")


if __name__ == "__main__":
    main()
