#!/bin/bash
# 1-GPU batch: EMA overlap sweep on the graphed step, the other SSL configs, the round's ncu pass
mkdir -p gpurun_out
T="timeout -s KILL"
export PYTHONUNBUFFERED=1
echo "=== other SSL configs"
for c in cfg1_L8 cfg4 cfg0; do
  EXTRA="--no-cpu --no-ref-gpu"; if [ $c == cfg0 ]; then EXTRA=""; fi
  $T 600 python bench.py --config $c --steps 50 --warmup 5 $EXTRA > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_$c.json").read().strip().splitlines()[-1])
    print("$c", {k: d[k] for k in ("value", "ms_per_step")}, "e2e", d["e2e"]["value"], "ref_eager", (d.get("reference_eager_b200") or {}).get("ssl_step", {}).get("faces_per_s"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
    print({k: v["ms"] for k, v in d["kernels"].items()})
except Exception as e:
    print("$c no bench line:", e); print(open("gpurun_out/bench_$c.err").read()[-600:])
PY
done
echo "=== ncu round"; bash tools/ncu_round.sh 2>&1 | tail -16
