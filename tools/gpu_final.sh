#!/bin/bash
# Final GPU visit of a round: the whole GPU test suite, smoke, the default bench line, then the ncu evidence of the SAME
# build (launch list of one step; --set full over one launch per kernel family).
mkdir -p gpurun_out
STEPS=20 bash tools/gpu_check.sh
cp gpurun_out/bench.log gpurun_out/final_bench_1gpu.json
bash tools/gpu_profile.sh
bash tools/gpu_ncu_full.sh
