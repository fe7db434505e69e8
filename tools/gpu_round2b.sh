#!/bin/bash
# GPU visit: head backward in one wave + parallel loss finalize + refined CTA-pair dispatch, against the previous build
mkdir -p gpurun_out
BASE=tools/ab/lib_cg2_v1.so
echo "== head bench, previous build"; B200UNET_LIB=$BASE timeout 200 python tools/head_bench.py 2>&1 | tail -3
echo "== head bench, this build"; timeout 200 python tools/head_bench.py 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_loss.py tests/test_gpu_fullsize.py tests/test_gpu_model.py -x -q > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2b_pytest.log
timeout 400 python tools/conv_bench.py > gpurun_out/convbench_all.log 2>&1; cat gpurun_out/convbench_all.log
bash tools/ab_step.sh $BASE 2
