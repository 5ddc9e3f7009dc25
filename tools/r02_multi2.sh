#!/bin/bash
# 2-GPU pass after the fused DINO head: N>1 parity tests (incl. the fused head's centre exchange) + the bench at N = 2
mkdir -p gpurun_out
T="timeout -s KILL"
export PYTHONUNBUFFERED=1
N=$(nvidia-smi -L | wc -l)
echo "=== $N GPUs: pytest multi"; $T 400 python -m pytest tests/test_gpu_multi.py -q -p no:cacheprovider --tb=short 2>&1 | tail -30 | tee gpurun_out/pytest_multi.txt
echo "=== bench --gpus 2"
$T 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
grep -v "^\[rank.\]:\[W\|^W1017\|^\*\*\*" gpurun_out/bench_n2.err | tail -c 500
python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/bench_n2.json").read().strip().splitlines() if l.startswith("{")][-1])
    print({k: d[k] for k in ("n_gpus", "value", "ms_per_step")}, "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"])
    print(d["config"]["centre_exchange"])
    for k, v in d["head"].items(): print(k, v["ms_fwd_bwd"], v.get("ms_by_exchange"), v.get("parity_max_rel"))
except Exception as e:
    print("no bench line:", e)
PY
