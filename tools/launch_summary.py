"""Summarise an ncu --csv launch list (one row per launch and metric): per-kernel totals of device time and DRAM
bytes, the step total, and the top launches.
    python tools/launch_summary.py gpurun_out/launches.csv [top] > profiles/NAME.md"""
import collections
import csv
import sys

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
rows = list(csv.reader(open(path)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr, data = rows[hi], rows[hi + 1:]
ki, mi, vi, ui, gi = (hdr.index(k) for k in ("Kernel Name", "Metric Name", "Metric Value", "Metric Unit", "Grid Size"))
launches = collections.OrderedDict()
for r in data:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(",", ""))
    u = r[ui]
    if r[mi] == "gpu__time_duration.sum":
        v = v / 1e3 if u in ("ns", "nsecond") else (v * 1e3 if u in ("ms", "msecond") else v)  # -> us
    else:
        v = v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    name = r[ki].split("(")[0].replace("void ", "").replace("b200::", "")
    L = launches.setdefault(r[0], {"name": name, "grid": r[gi]})
    L[r[mi]] = v
tot = collections.defaultdict(lambda: [0.0, 0.0, 0.0, 0])
for L in launches.values():
    t = tot[L["name"]]
    t[0] += L.get("gpu__time_duration.sum", 0.0)
    t[1] += L.get("dram__bytes_read.sum", 0.0)
    t[2] += L.get("dram__bytes_write.sum", 0.0)
    t[3] += 1
T = sum(t[0] for t in tot.values())
R = sum(t[1] for t in tot.values())
W = sum(t[2] for t in tot.values())
print(f"step total: {T / 1e3:.2f} ms over {len(launches)} launches; DRAM read {R / 1e9:.2f} GB, write {W / 1e9:.2f} GB")
print()
print("| kernel | launches | time [us] | share | DRAM read [GB] | DRAM write [GB] | GB/s while running |")
print("|---|---|---|---|---|---|---|")
for k, t in sorted(tot.items(), key=lambda kv: -kv[1][0]):
    bw = (t[1] + t[2]) / (t[0] * 1e-6) / 1e9 if t[0] > 0 else 0
    print(f"| `{k[:70]}` | {t[3]} | {t[0]:.1f} | {100 * t[0] / T:.1f} % | {t[1] / 1e9:.3f} | {t[2] / 1e9:.3f} | {bw:.0f} |")
print()
print("Top launches:")
print()
print("| time [us] | grid | DRAM read [MB] | DRAM write [MB] | kernel |")
print("|---|---|---|---|---|")
for L in sorted(launches.values(), key=lambda L: -L.get("gpu__time_duration.sum", 0.0))[:top]:
    print(f"| {L.get('gpu__time_duration.sum', 0):.1f} | {L['grid']} | {L.get('dram__bytes_read.sum', 0) / 1e6:.1f} | "
          f"{L.get('dram__bytes_write.sum', 0) / 1e6:.1f} | `{L['name'][:70]}` |")
