#!/bin/bash
# focused GPU pass for SURVEY 8f row 1 (fused DINO head): its tests, a per-call breakdown, a timing probe
mkdir -p gpurun_out
T="timeout -s KILL"
export PYTHONUNBUFFERED=1
echo "=== pytest dino head"; $T 420 python -m pytest tests/test_gpu_dino_head.py -q -p no:cacheprovider 2>&1 --tb=short | tail -70
echo "=== breakdown"; $T 240 python tools/dino_head_breakdown.py 2>&1 | tail -2 | tee gpurun_out/f1_breakdown.json
echo "=== probe"; $T 240 python tools/dino_head_probe.py > gpurun_out/f1_probe.json 2> gpurun_out/f1_probe.err; cat gpurun_out/f1_probe.json
