#!/bin/bash
# Two-GPU visit: gradient check of the bucketed all-reduce (ddp_check) and the data-parallel bench with and without CTA pairs
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR tools/ddp_check.py > gpurun_out/ddp_check.log 2>&1; echo "ddp_check rc=$?"; tail -4 gpurun_out/ddp_check.log
for v in 1 0; do
  B200UNET_CG2=$v timeout 400 $TR bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline --no-profile > gpurun_out/ddp2_cg$v.json 2> gpurun_out/ddp2_cg$v.err; echo "bench CG2=$v rc=$?"
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/ddp2_cg$v.json").read().strip().splitlines()[-1])
    print("CG2=$v: %.3f ms/step, %.0f img/s, e2e %s, per-rank %s, without all-reduce %s" % (d["ms_per_step"], d["value"], d.get("e2e", {}).get("value"), d.get("per_rank_ms_per_step"), d.get("per_rank_ms_per_step_without_allreduce")))
except Exception as e:
    print("CG2=$v failed:", e)
PY
  tail -3 gpurun_out/ddp2_cg$v.err
done
