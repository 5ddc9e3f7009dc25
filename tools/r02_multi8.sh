#!/bin/bash
# 8-GPU pass: the scaling bench at N = 4 and N = 8 (sharded head parity is checked inside bench.py at every N)
mkdir -p gpurun_out
T="timeout -s KILL"
export PYTHONUNBUFFERED=1
N=$(nvidia-smi -L | wc -l)
for n in 4 8; do
  if [ $n -gt $N ]; then continue; fi
  echo "=== bench --gpus $n"
  $T 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 50 --warmup 5 > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err
  grep -v "^\[rank.\]:\[W\|^W1017\|^\*\*\*" gpurun_out/bench_n$n.err | tail -c 300
  python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/bench_n$n.json").read().strip().splitlines() if l.startswith("{")][-1])
    print({k: d[k] for k in ("n_gpus", "value", "ms_per_step")}, "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"])
    print(d["config"]["centre_exchange"])
    for k, v in d["head"].items(): print(k, v["ms_fwd_bwd"], v.get("ms_by_exchange"), v.get("parity_max_rel"))
except Exception as e:
    print("no bench line:", e)
PY
done
