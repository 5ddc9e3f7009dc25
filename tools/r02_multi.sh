#!/bin/bash
# multi-GPU batch: N>1 parity tests + the scaling bench at N = number of visible GPUs
mkdir -p gpurun_out
T="timeout -s KILL"
export PYTHONUNBUFFERED=1
N=$(nvidia-smi -L | wc -l)
echo "=== $N GPUs: pytest multi"; $T 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_optim.py tests/test_gpu_patch_embed.py tests/test_gpu_vit_face.py -q -p no:cacheprovider 2>&1 | tail -8
for n in 2 $N; do
  if [ $n -gt $N ]; then continue; fi
  echo "=== bench --gpus $n"
  $T 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 50 --warmup 5 > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err
  grep -v "^\[rank.\]:\[W\|^W1017\|^\*\*\*" gpurun_out/bench_n$n.err | tail -c 500
  python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/bench_n$n.json").read().strip().splitlines() if l.startswith("{")][-1])
    print({k: d[k] for k in ("n_gpus", "value", "ms_per_step")}, "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"].get("host_numa"))
    print(d["config"]["centre_exchange"])
    for k, v in d["kernels"].items(): print(k, v["ms"])
    for k, v in d["head"].items(): print(k, v["ms_fwd_bwd"], v.get("ms_by_exchange"), v.get("parity_vs_unsharded"), v.get("parity_max_rel"))
except Exception as e:
    print("no bench line:", e)
PY
  if [ $n -eq $N ]; then break; fi
done
