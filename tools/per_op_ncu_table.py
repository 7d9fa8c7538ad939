"""Markdown table of one forward + decode from an `ncu --csv` launch log (long format, one metric per row):

    ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,l1tex__m_xbar2l1tex_read_bytes.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__cycles_elapsed.avg.per_second \
        --clock-control none -s <launches of the first detect()> -c <launches of one detect()> --csv --log-file X.csv python tools/run_forward_once.py
    python tools/per_op_ncu_table.py X.csv [ops.txt] > X.md

ops.txt = the "ops: ..." line run_forward_once.py prints (op names in launch order); the stem expands to 3 launches and the
decode to 2.  Runs on a CPU-only box."""
import csv
import re
import sys


def main():
    path = sys.argv[1]
    ops = []
    if len(sys.argv) > 2:
        for line in open(sys.argv[2]):
            if line.startswith("ops:"):
                ops = line.split()[1:]
    rows = [r for r in csv.reader(open(path, newline="")) if len(r) >= 15 and r[0].isdigit()]
    launches = {}
    for r in rows:
        d = launches.setdefault(int(r[0]), {"kernel": r[4]})
        d[r[12]] = float(r[14].replace(",", ""))
    names = []
    for op in ops:
        names += [f"{op} (im2row)", f"{op} (4x1 conv)", f"{op} (3x3/2 max-pool)"] if op == "backbone.stem" else [op]
    names += ["decode (peaks)", "decode (select)"]
    print("| # | op | kernel | us | tensor pipe active % | SM clock GHz | L2->SM MB | DRAM read MB | DRAM write MB |")
    print("|---|---|---|---|---|---|---|---|---|")
    tot = 0.0
    for i, (k, d) in enumerate(sorted(launches.items())):
        kern = re.sub(r"\(.*", "", d["kernel"]).replace("void ", "").replace("cnl::", "")
        us = d.get("gpu__time_duration.sum", 0.0) / 1e3
        tot += us
        print(f"| {i} | {names[i] if i < len(names) else ''} | `{kern}` | {us:.1f} | "
              f"{d.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 0.0):.1f} | "
              f"{d.get('sm__cycles_elapsed.avg.per_second', 0.0) / 1e9:.2f} | {d.get('l1tex__m_xbar2l1tex_read_bytes.sum', 0.0) / 1e6:.0f} | "
              f"{d.get('dram__bytes_read.sum', 0.0) / 1e6:.0f} | {d.get('dram__bytes_write.sum', 0.0) / 1e6:.0f} |")
    print(f"\nSum of the {len(launches)} launches: {tot / 1e3:.3f} ms (serialised, cold caches, burst clocks: not the step time).")


if __name__ == "__main__":
    main()
