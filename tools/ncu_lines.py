"""Developer tool: stall samples per CUDA source line.
    ncu -i X.ncu-rep --page source --csv --print-source sass,cuda --kernel-name regex:K > s.csv; python tools/ncu_lines.py s.csv [N] [which]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
fn = [i for i, r in enumerate(rows) if r and r[0] == "Function Name"]
names = [rows[i][1] for i in fn]
# blocks of the same function name appear once per file; group by kernel instance = change of name sequence
uniq = []
for n in names:
    if not uniq or uniq[-1] != n:
        uniq.append(n)
target = uniq[which]
print("kernel:", target[:110])
lines = {}
tot = 0
cur_file = None
active = False
for i, r in enumerate(rows):
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
    elif r[0] == "Function Name":
        active = r[1] == target
    elif active and r[0].isdigit() and len(r) > 6:
        try:
            s = int(r[6])
        except ValueError:
            continue
        key = (cur_file, int(r[0]))
        if key in lines:
            lines[key] = (lines[key][0] + s, r[1])
        else:
            lines[key] = (s, r[1])
        tot += s
print("total samples", tot)
for (f, ln), (s, src) in sorted(lines.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{s:7d} {100 * s / max(tot, 1):5.1f}%  {f}:{ln:<4d} {src.strip()[:100]}")
