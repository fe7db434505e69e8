#!/bin/bash
# ncu evidence for ONE training step at the BASELINE size (batch 32, 512^2): every launch with its device time and
# DRAM bytes (cold-cache, serialised: compare SHARES and bytes, not absolute times).  Output: gpurun_out/launches.csv
# (summarised into profiles/ by tools/launch_summary.py).  Per-kernel `--set full` captures: tools/prof_ops.py.
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-profile"
$CMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain.log; exit 1; }
tail -c 400 gpurun_out/plain.log
L=$(python -c "import json;print(json.loads(open('gpurun_out/plain.log').read().strip().splitlines()[-1])['gpu_launches'])")
echo "launches per step: $L"
# torch's own small kernels (fills, RNG) are interleaved: count ALL launches of the 3 warm-up steps by a first cheap pass
timeout ${NCU_TIMEOUT:-900} ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    -k regex:'gconv|nconv|pconv|wgrad|in_|stem|upsample|head_|loss_|pack_weights|image_to|stats_partial' -s $((3*L)) -c $L \
    --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu1.log 2>&1
echo "ncu launches rc=$?"; wc -l gpurun_out/launches.csv
