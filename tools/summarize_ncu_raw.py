"""`ncu -i X.ncu-rep --page raw --csv` -> a markdown table of the metrics the roofline discussion uses, one row per
profiled launch.  usage: summarize_ncu_raw.py raw.csv "title" > out.md"""
import csv
import sys

WANT = [("gpu__time_duration.sum", "duration", None),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe % (active)", None),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU pipe %", None),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots %", None),
        ("dram__bytes_read.sum", "DRAM read", None), ("dram__bytes_write.sum", "DRAM write", None),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM % of peak", None),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 % of peak", None),
        ("l1tex__m_xbar2l1tex_read_bytes.sum", "L2->SM read", None),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %", None),
        ("launch__registers_per_thread", "regs", None)]


def main(path, title):
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {}
    for i, h in enumerate(hdr):
        for key, _, _ in WANT:
            if h.endswith(key) and key not in col:
                col[key] = i
    k_name, k_grid, k_block = hdr.index("Kernel Name"), hdr.index("Grid Size"), hdr.index("Block Size")
    print(f"# {title}\n")
    print("Source: `ncu --set full --clock-control none --import-source on --profile-from-start off` over "
          "`python tools/profile_r2.py` (one warm-up pass, then the profiled pass; cold caches, serialised launches: "
          "absolute durations are upper bounds of what the kernels cost inside the step graph).  Raw report: "
          "`gpurun_out/r2_final_kernels.ncu-rep`.\n")
    heads = ["#", "kernel", "grid x block"] + [label for key, label, _ in WANT if key in col]
    print("| " + " | ".join(heads) + " |")
    print("|" + "---|" * len(heads))
    for n, r in enumerate(data):
        if len(r) < len(hdr):
            continue
        name = r[k_name].replace("void pcdm::", "").replace("pcdm::", "")
        name = name[: name.index("(")] if "(" in name else name
        cells = [str(n), f"`{name}`", f"{r[k_grid]} x {r[k_block]}"]
        for key, label, _ in WANT:
            if key not in col:
                continue
            v, u = r[col[key]], units[col[key]]
            try:
                f = float(v.replace(",", ""))
                cells.append(f"{f:,.1f} {u}".strip() if abs(f) < 1e6 else f"{f / 1e6:,.1f} M{u}".strip())
            except ValueError:
                cells.append(v)
        print("| " + " | ".join(cells) + " |")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "ncu summary")
