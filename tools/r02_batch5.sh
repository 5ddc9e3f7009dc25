#!/bin/bash
mkdir -p gpurun_out
T="timeout -s KILL"
export PYTHONUNBUFFERED=1
echo "=== pytest -m gpu (PDL launches everywhere)"; $T 1200 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -15
echo "=== smoke"; $T 300 python __graft_entry__.py --smoke 2>&1 | tail -2
for v in 1 0; do
echo "=== bench ours LAFS_PDL=$v"; LAFS_PDL=$v $T 900 python bench.py --no-cpu --no-ref-gpu > gpurun_out/bench_pdl$v.json 2> gpurun_out/bench_pdl$v.err; grep -v "^\[rank0\]:\[W" gpurun_out/bench_pdl$v.err | tail -c 300; python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_pdl$v.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step")}, "eager", d["value_eager_launches"], "e2e", d["e2e"]["value"])
    for k, v in d["kernels"].items(): print(k, v["ms"])
    for k, v in d["head"].items(): print(k, v["ms_fwd_bwd"], v["ms_fwd_bwd_eager"], v["frac_tc"])
    print("extras", d.get("extras"))
except Exception as e:
    print("no bench line:", e)
PY
done
