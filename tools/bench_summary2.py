"""Developer tool: one-line digest of bench.py JSON lines.  python tools/bench_summary2.py gpurun_out/b*.log"""
import json
import sys

for f in sys.argv[1:]:
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
    except Exception as e:  # noqa: BLE001
        print(f, "ERR", e)
        continue
    bd = d.get("breakdown_ms_per_step") or {}
    r = d.get("roofline") or {}
    print(f, "value", round(d["value"], 1), "ms", round(d["ms_per_step"], 3), "eager", d.get("eager") and round(d["eager"]["ms_per_step"], 3),
          "e2e", round(d["e2e"]["ms_per_step"], 3), "sgd", d.get("with_optimizer_step") and round(d["with_optimizer_step"]["ms_per_step"], 3),
          "launches/step", d["gpu_launches"] / d["steps"], "frac", r and round(r["frac"], 3), "serial", r and round(r["serial_ms_per_step"], 3),
          "clk", d["clocks"] and d["clocks"]["sm_mhz"], "per-rank", d.get("per_rank_ms_per_step"), "ar", d.get("allreduce_only_ms"),
          (d.get("graph") or {}).get("error"))
    print("   ", {k: round(v["ms_per_step"], 3) for k, v in bd.items() if k != "_note"})
