#!/bin/bash
mkdir -p gpurun_out
T="timeout -s KILL"
run() { echo "=== $1"; shift; "$@" 2>&1 | tail -${TAILN:-4}; }
export PYTHONUNBUFFERED=1
run "head tests (defaults)" $T 300 python -m pytest tests/test_gpu_head.py -q -p no:cacheprovider
for c in cfg3 cfg4; do
  TAILN=1 run "breakdown $c" $T 300 python tools/head_breakdown.py $c | tee -a gpurun_out/head_breakdown4.jsonl
done
