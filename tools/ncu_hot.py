"""Developer tool: top SASS lines by stall samples from `ncu -i rep --page source --csv` output.
    ncu -i X.ncu-rep --page source --csv --kernel-name regex:K --launch-count 1 > src.csv; python tools/ncu_hot.py src.csv [N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[hi]
data = rows[hi + 1:]
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0  # n-th kernel of the file
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
lo = starts[which]
end = starts[which + 1] if which + 1 < len(starts) else len(rows)
rows = rows[lo:end]
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
si = hdr.index("# Samples")
src = hdr.index("Source")
ie = hdr.index("Instructions Executed")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[si] or 0) for r in data if len(r) > si)
print("kernel:", rows[0][1][:100], "total samples", tot)
agg = {}
for r in data:
    for i, h in stall_cols:
        agg[h] = agg.get(h, 0) + int(r[i] or 0)
print("stall totals:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
idx = sorted(range(len(data)), key=lambda j: -int(data[j][si] or 0))[:top]
for j in sorted(idx):
    r = data[j]
    st = sorted(((int(r[i] or 0), h) for i, h in stall_cols), reverse=True)[:2]
    print(f"{j:5d} {int(r[si]):6d} {100 * int(r[si]) / max(tot, 1):5.1f}% exec={r[ie]:>8s} {r[src].strip()[:70]:70s} {st}")
