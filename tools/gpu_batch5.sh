#!/bin/bash
mkdir -p gpurun_out
T="timeout -s KILL"
run() { echo "=== $1"; shift; "$@" 2>&1 | tail -${TAILN:-4}; }
export PYTHONUNBUFFERED=1
run "head + patch-embed tests" $T 300 python -m pytest tests/test_gpu_head.py tests/test_gpu_patch_embed.py -q -p no:cacheprovider
for c in cfg3 cfg4; do
  TAILN=1 run "breakdown $c" $T 300 python tools/head_breakdown.py $c | tee -a gpurun_out/head_breakdown5.jsonl
done
echo "=== microbench"; $T 200 python tools/microbench.py > gpurun_out/microbench.log 2>&1; python - <<'PY'
import json
d=json.load(open('gpurun_out/microbench.json'))
for k,v in d.items():
    print(k, {a:round(b,4) if isinstance(b,float) else b for a,b in v.items()})
PY
