"""Summarise an ncu --csv launch list (gpu__time_duration.sum per launch): totals per kernel and the top launches."""
import collections
import csv
import sys

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
rows = list(csv.reader(open(path)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr, data = rows[hi], rows[hi + 1:]
ki, vi, ui, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("Grid Size")
tot, cnt, lst = collections.defaultdict(float), collections.Counter(), []
for r in data:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(",", ""))
    v = v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)
    name = r[ki].split("(")[0].replace("void ", "").replace("b200::", "")
    tot[name] += v
    cnt[name] += 1
    lst.append((v, name, r[gi]))
T = sum(tot.values())
print(f"total {T:.1f} us over {len(lst)} launches")
for k, v in sorted(tot.items(), key=lambda x: -x[1]):
    print(f"{v:10.1f} us {100 * v / T:5.1f}% n={cnt[k]:3d} {k[:90]}")
print()
for v, n, g in sorted(lst, reverse=True)[:top]:
    print(f"{v:9.1f} {g:22s} {n[:80]}")
