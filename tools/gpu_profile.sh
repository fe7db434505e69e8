#!/bin/bash
# ncu evidence for one training step at the BASELINE size: (1) launch list with device times, (2) full capture of
# the heaviest kernels.  Output under gpurun_out/ (copied to profiles/ by hand).
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-profile"
$CMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain.log; exit 1; }
tail -c 600 gpurun_out/plain.log
L=$(python -c "import json;print(json.loads(open('gpurun_out/plain.log').read().strip().splitlines()[-1])['gpu_launches'])")
echo "launches per step: $L"
ncu --metrics gpu__time_duration.sum --clock-control none -s $((3*L)) -c $L --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu1.log 2>&1
echo "ncu launches rc=$?"; wc -l gpurun_out/launches.csv
if [ -n "$FULL" ]; then
ncu --set full --clock-control none --import-source on -k regex:"$FULL" -s ${FULL_SKIP:-0} -c ${FULL_COUNT:-6} -f -o gpurun_out/prof $CMD > gpurun_out/ncu2.log 2>&1
echo "ncu full rc=$?"; ls -la gpurun_out/prof.ncu-rep
fi
