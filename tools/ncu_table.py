"""Developer tool: markdown table of the per-launch metrics of an `ncu --set full` report.
    ncu -i X.ncu-rep --page raw --csv > raw.csv; python tools/ncu_table.py raw.csv > profiles/NAME.md"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
cols = [("Kernel Name", "kernel"), ("Grid Size", "grid"), ("Block Size", "block"), ("gpu__time_duration.sum", "time"),
        ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %"),
        ("launch__registers_per_thread", "regs"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %")]
idx = [(hdr.index(c), n) for c, n in cols if c in hdr]
print("| " + " | ".join(f"{n} [{units[i]}]" if units[i] else n for i, n in idx) + " |")
print("|" + "---|" * len(idx))
for r in data:
    out = []
    for i, n in idx:
        v = r[i]
        if n == "kernel":
            v = v.replace("void ", "").replace("b200::", "").split("(")[0][:60]
        else:
            try:
                v = f"{float(v.replace(',', '')):.3f}".rstrip("0").rstrip(".") if "." in v else v
            except ValueError:
                pass
        out.append(v)
    print("| " + " | ".join(out) + " |")
