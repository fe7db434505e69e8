#!/bin/bash
# A/B of two builds of libb200unet.so on ONE box: alternating bench.py runs (value = graph replay of the step).
# usage: tools/ab_step.sh <base.so> [rounds]
BASE=$1; R=${2:-2}
mkdir -p gpurun_out
for i in $(seq 1 $R); do
  for v in base new; do
    if [ $v = base ]; then export B200UNET_LIB=$BASE; else unset B200UNET_LIB; fi
    timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-profile --no-e2e > gpurun_out/ab_$v$i.json 2> gpurun_out/ab_$v$i.err
    python - <<PY
import json
d = json.loads(open("gpurun_out/ab_$v$i.json").read().strip().splitlines()[-1])
print("$v $i: %.3f ms/step (eager %s)" % (d["ms_per_step"], d.get("eager", {}).get("ms_per_step")))
PY
  done
done
