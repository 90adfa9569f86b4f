"""Summarise an `ncu --metrics gpu__time_duration.sum` launch list (CSV) of tools/profile_step.py into markdown:
per-kernel launch counts, total device time and share for the LAST UNet evaluation in the capture."""
import collections
import csv
import re
import sys


def load(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    r = csv.reader(lines)
    hdr = next(r)
    idx = {h: i for i, h in enumerate(hdr)}
    rows = [(row[idx["Kernel Name"]], float(row[idx["Metric Value"]].replace(",", "")) / 1e3, row[idx["Grid Size"]],
             row[idx["Block Size"]]) for row in r if len(row) >= len(hdr)]
    last = [i for i, (n, *_r) in enumerate(rows) if "timestep_embedding" in n][-1]
    return rows[last:]


def main(path, title):
    fw = [r for r in load(path) if "at::" not in r[0]]
    agg = collections.OrderedDict()
    for n, v, g, b in fw:
        k = re.sub(r"\(.*", "", n).replace("void pcdm::", "").replace("pcdm::", "")
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"# {title}\n")
    print(f"Source: `{path}` — `ncu --metrics gpu__time_duration.sum --clock-control none` over `python "
          f"tools/profile_step.py 2` (eager, B=16, 32x64 latents, 258 tokens, bf16); last UNet evaluation only. "
          f"ncu serialises launches and flushes caches, so compare SHARES, not absolutes.\n")
    print(f"launches: {len(fw)}, summed device time: {tot / 1e3:.2f} ms\n")
    print("| kernel | launches | total us | share |\n|---|---:|---:|---:|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {v[0]} | {v[1]:.1f} | {100 * v[1] / tot:.1f}% |")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "launch list")
