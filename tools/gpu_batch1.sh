#!/bin/bash
# development batch: each variant in its own process under a hard timeout (a hung kernel must not eat the budget)
mkdir -p gpurun_out
T="timeout -s KILL"
run() { echo "=== $1"; shift; "$@" 2>&1 | tail -${TAILN:-6}; }
export PYTHONUNBUFFERED=1
run "A head tests, 1SM + unfused dW (8-warp epilogue only)" env LAFS_HEAD_1SM=1 LAFS_DW_UNFUSED=1 $T 200 python -m pytest tests/test_gpu_head.py -q -k "not variants" -p no:cacheprovider
run "B backward tests, fused dW no cluster" env LAFS_HEAD_1SM=1 LAFS_DW_CLUSTER=1 $T 120 python -m pytest tests/test_gpu_head.py -q -k "backward" -p no:cacheprovider
run "C backward tests, fused dW cluster 2" env LAFS_HEAD_1SM=1 LAFS_DW_CLUSTER=2 $T 120 python -m pytest tests/test_gpu_head.py -q -k "backward" -p no:cacheprovider
run "D backward tests, fused dW cluster 4" env LAFS_HEAD_1SM=1 LAFS_DW_CLUSTER=4 $T 120 python -m pytest tests/test_gpu_head.py -q -k "backward" -p no:cacheprovider
run "E head tests, CTA pairs, dW unfused" env LAFS_DW_UNFUSED=1 $T 200 python -m pytest tests/test_gpu_head.py -q -k "not variants" -p no:cacheprovider
run "F all head tests, defaults" $T 300 python -m pytest tests/test_gpu_head.py -q -p no:cacheprovider
run "G dino / ssl step tests" $T 300 python -m pytest tests/test_gpu_ema_dino.py tests/test_gpu_ssl_step.py -q -p no:cacheprovider
for v in "LAFS_HEAD_1SM=1 LAFS_DW_UNFUSED=1" "LAFS_HEAD_1SM=1 LAFS_DW_CLUSTER=1" "LAFS_HEAD_1SM=1 LAFS_DW_CLUSTER=2" "LAFS_HEAD_1SM=0 LAFS_DW_CLUSTER=4"; do
  TAILN=1 run "breakdown cfg3 $v" env $v $T 120 python tools/head_breakdown.py cfg3 | tee -a gpurun_out/head_breakdown.jsonl
done
TAILN=1 run "breakdown cfg4 1SM unfused" env LAFS_HEAD_1SM=1 LAFS_DW_UNFUSED=1 $T 120 python tools/head_breakdown.py cfg4 | tee -a gpurun_out/head_breakdown.jsonl
TAILN=1 run "breakdown cfg4 default" $T 120 python tools/head_breakdown.py cfg4 | tee -a gpurun_out/head_breakdown.jsonl
echo "=== microbench"; $T 200 python tools/microbench.py > gpurun_out/microbench.log 2>&1; python - <<'PY'
import json
d=json.load(open('gpurun_out/microbench.json'))
for k,v in d.items():
    if k.startswith(('dino','ema')): print(k, {a:round(b,4) if isinstance(b,float) else b for a,b in v.items()})
PY
