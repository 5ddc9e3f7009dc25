#!/bin/bash
mkdir -p gpurun_out
T="timeout -s KILL"
run() { echo "=== $1"; shift; "$@" 2>&1 | tail -${TAILN:-4}; }
export PYTHONUNBUFFERED=1
run "head tests (defaults)" $T 300 python -m pytest tests/test_gpu_head.py -q -p no:cacheprovider
run "head backward, dE no cluster" env LAFS_DE_CLUSTER=1 $T 120 python -m pytest tests/test_gpu_head.py -q -k "backward" -p no:cacheprovider
run "head backward, dE cluster 2" env LAFS_DE_CLUSTER=2 $T 120 python -m pytest tests/test_gpu_head.py -q -k "backward" -p no:cacheprovider
TAILN=12 run "patch-embed + vit_face + patches tests" $T 300 python -m pytest tests/test_gpu_patch_embed.py tests/test_gpu_vit_face.py tests/test_gpu_patches.py -q -p no:cacheprovider
for c in cfg3 cfg4; do
  TAILN=1 run "breakdown $c" $T 200 python tools/head_breakdown.py $c | tee -a gpurun_out/head_breakdown2.jsonl
done
