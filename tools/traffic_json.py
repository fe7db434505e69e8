"""Per-family device time and DRAM bytes of one step from an ncu launch list -> profiles/NAME.json (read by bench.py
for roofline.traffic).    python tools/traffic_json.py gpurun_out/launches.csv profiles/r1_traffic.json"""
import collections
import csv
import json
import re
import sys

FAMILIES = [("conv", r"gconv|nconv|pconv|wgrad|image_to_nhwc32|pack_weights"), ("norm", r"^in_"), ("resample", r"upsample"),
            ("head", r"^head_"), ("loss", r"^loss_"), ("stem_simt", r"^stem_")]
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[hi]
ki, mi, vi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
L = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) > vi:
        L.setdefault(r[0], {"name": r[ki].replace("void ", "").replace("b200::", "").split("(")[0]})[r[mi]] = float(r[vi].replace(",", ""))
out = {f: {"launches": 0, "time_us": 0.0, "dram_read_bytes": 0.0, "dram_write_bytes": 0.0} for f, _ in FAMILIES}
for d in L.values():
    for f, rx in FAMILIES:
        if re.search(rx, d["name"]):
            o = out[f]
            o["launches"] += 1
            o["time_us"] += d.get("gpu__time_duration.sum", 0.0) / 1e3
            o["dram_read_bytes"] += d.get("dram__bytes_read.sum", 0.0)
            o["dram_write_bytes"] += d.get("dram__bytes_write.sum", 0.0)
            break
out["_source"] = "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum over one step (tools/gpu_profile.sh), batch 32, 512^2"
json.dump(out, open(sys.argv[2], "w"), indent=1)
print(json.dumps(out, indent=1))
