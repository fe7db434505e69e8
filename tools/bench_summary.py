"""Pretty-print the JSON line(s) of bench.py read from stdin (developer convenience)."""
import json
import sys

for line in sys.stdin:
    line = line.rstrip()
    if line.startswith('{"metric"') or line.startswith('{"impl"'):
        d = json.loads(line)
        r = d.get("roofline") or {}
        print(f"value {d['value']:.1f} {d['unit']}  ms/step {d['ms_per_step']:.3f}  e2e {d['e2e']['value']:.1f}  "
              f"conv {r.get('achieved', 0):.0f} TF/s = {r.get('frac', 0):.3f} of peak  launches {d.get('gpu_launches')}  "
              f"clocks {d.get('clocks')}")
        for k, v in (d.get("breakdown_ms_per_step") or {}).items():
            print("   %-22s %7.3f ms  x%d" % (k, v["ms_per_step"], v["calls_per_step"]))
        if d.get("cpu_baseline"):
            print("   cpu_baseline", d["cpu_baseline"])
    else:
        print(line[:400])
