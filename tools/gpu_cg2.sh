#!/bin/bash
# One GPU-box visit for the CTA-pair (cta_group::2) gconv configurations: operand-mapping probe (one process per case,
# a trap cannot take the rest down), conv tests, per-layer and whole-step A/B against B200UNET_CG2=0.
mkdir -p gpurun_out
L=gpurun_out/cg2_probe.log; : > $L
fail=0
for i in 0 1 2 3 4 5 6 7; do
  timeout 120 python tools/cg2_probe.py --case $i >> $L 2>&1; rc=$?
  echo "case $i rc=$rc" >> $L
  [ $rc -ne 0 ] && fail=$((fail+1))
  if [ $i -eq 1 ] && [ $fail -eq 2 ]; then echo "first two cases failed: stopping the probe" >> $L; break; fi
done
grep -v "^$" $L | tail -60
if [ $fail -ne 0 ] || grep -q "MISMATCH\|timed out" $L; then echo "probe not clean ($fail failed): skipping tests and A/B"; exit 0; fi
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_fullsize.py -x -q -k "conv" > gpurun_out/cg2_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/cg2_pytest.log
LAY=e2c2,e3c2,e4c2,e5c2,d0c1,d0c2,d1c1,d1c2,d2c1,d2c2,d3c1
B200UNET_CG2=0 timeout 300 python tools/conv_bench.py --only fprop,dgrad --layers $LAY > gpurun_out/convbench_cg1.log 2>&1
timeout 300 python tools/conv_bench.py --only fprop,dgrad --layers $LAY > gpurun_out/convbench_cg2.log 2>&1
paste -d'\n' gpurun_out/convbench_cg1.log gpurun_out/convbench_cg2.log
for i in 1 2; do
  for v in 0 1; do
    B200UNET_CG2=$v timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-profile --no-e2e > gpurun_out/cg2_step_$v$i.json 2> gpurun_out/cg2_step_$v$i.err
    python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/cg2_step_$v$i.json").read().strip().splitlines()[-1])
    print("CG2=$v run $i: %.3f ms/step (eager %s)" % (d["ms_per_step"], d.get("eager", {}).get("ms_per_step")))
except Exception as e:
    print("CG2=$v run $i: failed", e)
PY
  done
done
