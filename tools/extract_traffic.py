"""ncu CSV of tools/profile_convs.py (dram__bytes_read.sum, dram__bytes_write.sum, gpu__time_duration.sum per launch)
-> profiles/roofline_traffic.json: measured DRAM bytes per conv3x3 call of one UNet evaluation (bench.py reads
`conv3x3_igemm_dram_bytes_per_launch` into roofline.traffic).  usage: extract_traffic.py in.csv out.json [git head]"""
import collections
import csv
import json
import sys

CALLS = 52   # conv3x3 calls per UNet evaluation (49 convs + 3 Upsample2D convs); split-K adds finishing launches


def main(path, out, head=None):
    lines = [l for l in open(path) if not l.startswith("==")]
    r = csv.reader(lines)
    hdr = next(r)
    idx = {h: i for i, h in enumerate(hdr)}
    per = collections.OrderedDict()
    for row in r:
        if len(row) < len(hdr):
            continue
        key = row[idx["ID"]]
        d = per.setdefault(key, {"kernel": row[idx["Kernel Name"]]})
        v = float(row[idx["Metric Value"]].replace(",", ""))
        unit = row[idx["Metric Unit"]].lower()
        name = row[idx["Metric Name"]]
        if "bytes" in name:
            v *= {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(unit, 1)
        elif "time" in name:
            v *= {"ns": 1e-3, "us": 1, "usecond": 1, "ms": 1e3, "nsecond": 1e-3, "msecond": 1e3}.get(unit, 1)
        d[name] = v
    launches = list(per.values())
    rd = sum(l.get("dram__bytes_read.sum", 0) for l in launches)
    wr = sum(l.get("dram__bytes_write.sum", 0) for l in launches)
    us = sum(l.get("gpu__time_duration.sum", 0) for l in launches)
    doc = {
        "what": "DRAM bytes of the conv3x3 launches of one UNet evaluation (BASELINE config 2, B=16, bf16), ncu "
                "--clock-control none, default cache control (flushed before each launch: cold-cache traffic)",
        "source": path, "git_head": head, "kernel_launches": len(launches), "conv_calls": CALLS,
        "dram_bytes_read_total": rd, "dram_bytes_write_total": wr, "duration_us_total_cold": us,
        "conv3x3_igemm_dram_bytes_per_launch": (rd + wr) / CALLS,
        "algorithmic_bytes_per_launch_note": "inputs + weights + outputs (+ residual) of a call: see DESIGN.md §4; the "
                                             "32x64 320->320 conv is 64.8 MB",
    }
    json.dump(doc, open(out, "w"), indent=1)
    print(json.dumps({k: doc[k] for k in ("kernel_launches", "conv3x3_igemm_dram_bytes_per_launch", "duration_us_total_cold")}))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
