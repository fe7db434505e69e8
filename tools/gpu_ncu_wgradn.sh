#!/bin/bash
# ncu --set full of the narrow weight-gradient kernels at their real shapes (after the same command ran without ncu).
mkdir -p gpurun_out
CASES=${CASES:-e0c2,e1c2}
timeout 200 python tools/prof_ops.py --cases $CASES > gpurun_out/prof_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/prof_plain.log; exit 1; }
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'wgradn' -o /tmp/wg -f \
    python tools/prof_ops.py --cases $CASES > gpurun_out/ncu_wg.log 2>&1
echo "ncu rc=$?"
ncu -i /tmp/wg.ncu-rep --page raw --csv > gpurun_out/ncu_wg_raw.csv 2>/dev/null
ncu -i /tmp/wg.ncu-rep --page source --csv --launch-count 1 > gpurun_out/ncu_wg_src.csv 2>/dev/null
wc -l gpurun_out/ncu_wg_raw.csv gpurun_out/ncu_wg_src.csv
