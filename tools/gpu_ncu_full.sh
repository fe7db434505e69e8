#!/bin/bash
# `ncu --set full` over one launch per kernel family at its real layer shape (tools/prof_ops.py), after the same
# command exited 0 without ncu.  The report stays on the box; only its raw page (CSV) and the source-page hot lines of
# the top kernels travel back in gpurun_out/.
mkdir -p gpurun_out
CASES=${CASES:-e0c2,d4c1,d3c1,d1c1,e2c2,e1c1,norm512,up256,head,loss}
KRE='gconv|nconv|pconv|wgrad|in_|upsample|head_|loss_'
timeout 300 python tools/prof_ops.py --cases $CASES > gpurun_out/prof_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/prof_plain.log; exit 1; }
timeout ${NCU_TIMEOUT:-1200} ncu --set full --clock-control none --import-source on -k regex:"$KRE" -o /tmp/full_r2 -f \
    python tools/prof_ops.py --cases $CASES > gpurun_out/ncu_full.log 2>&1
echo "ncu rc=$?"
ncu -i /tmp/full_r2.ncu-rep --page raw --csv > gpurun_out/ncu_full_raw.csv 2>/dev/null
wc -l gpurun_out/ncu_full_raw.csv
for k in pconv_kernel upsample2x_fwd_kernel; do
  ncu -i /tmp/full_r2.ncu-rep --page source --csv --kernel-name regex:$k --launch-count 1 > gpurun_out/ncu_src_$k.csv 2>/dev/null
done
ls -la gpurun_out/*.csv
