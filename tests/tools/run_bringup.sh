#!/bin/bash
# Runs every bring-up case in its own process (a faulting kernel only loses its own case).
# usage: tools/run_bringup.sh [case ...]   -> gpurun_out/bringup.log
mkdir -p gpurun_out
LOG=gpurun_out/bringup.log
: > $LOG
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.used --format=csv >> $LOG 2>&1
CASES="$@"
if [ -z "$CASES" ]; then CASES=$(python tests/tools/bringup.py list); fi
for c in $CASES; do
  echo "=== $c" >> $LOG
  timeout 240 python tests/tools/bringup.py $c >> $LOG 2>&1
  echo "exit=$?" >> $LOG
done
grep -c '"ok": true' $LOG; grep '"ok": false' $LOG | cut -c1-300; grep -E "exit=[1-9]|timed out|Error|error" $LOG | head -20
