#!/bin/bash
# focused GPU pass for SURVEY 8f row 1 (fused DINO head): its tests, the smoke, a timing probe
mkdir -p gpurun_out
T="timeout -s KILL"
export PYTHONUNBUFFERED=1
echo "=== pytest dino head"; $T 420 python -m pytest tests/test_gpu_dino_head.py -q -p no:cacheprovider -x 2>&1 | tail -25
echo "=== probe"; $T 240 python tools/dino_head_probe.py > gpurun_out/f1_probe.json 2> gpurun_out/f1_probe.err; tail -c 400 gpurun_out/f1_probe.err; cat gpurun_out/f1_probe.json
echo "=== smoke"; $T 240 python __graft_entry__.py --smoke 2>&1 | tail -3
