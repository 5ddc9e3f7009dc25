#!/bin/bash
# GPU batch 3 of round 2: ablations of the patch-embed kernel in both orientations, the full suite (student embed
# backward, saved tokens), bench ours (+ reference_eager_b200) and the reference arm.
mkdir -p gpurun_out
T="timeout -s KILL"
export PYTHONUNBUFFERED=1
echo "=== pytest -m gpu"; $T 1200 python -m pytest tests -m gpu -q -p no:cacheprovider -x 2>&1 | tail -5
echo "=== pe ablations, token-major (default)"; $T 300 python tools/pe_ablate.py 2>&1 | tee gpurun_out/pe_ablate_tokm.txt
echo "=== pe ablations, dims-major (LAFS_PE_TOKN=1)"; LAFS_PE_TOKN=1 $T 300 python tools/pe_ablate.py 2>&1 | tee gpurun_out/pe_ablate_tokn.txt
echo "=== bench ours"; $T 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; grep -v "^\[rank0\]:\[W" gpurun_out/bench_n1.err | tail -c 600; python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_n1.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step", "e2e", "roofline", "clocks")})
    for k, v in d["kernels"].items(): print(k, v)
    for k, v in d["head"].items(): print(k, v)
    print("reference_eager_b200", d.get("reference_eager_b200")); print("cpu_baseline", d.get("cpu_baseline"))
except Exception as e:
    print("no bench line:", e)
PY
echo "=== bench reference"; $T 600 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; head -c 1200 gpurun_out/bench_ref.json; tail -c 300 gpurun_out/bench_ref.err; echo
