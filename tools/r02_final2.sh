#!/bin/bash
# final 1-GPU pass of round 2 (after the fused DINO head): full GPU suite, smoke, bench (ours + reference arm),
# ncu launch lists (fused DINO head vs its unfused form with DRAM bytes; the bench command)
mkdir -p gpurun_out
T="timeout -s KILL"
export PYTHONUNBUFFERED=1
echo "=== pytest -m gpu"; $T 1200 python -m pytest tests -m gpu -q -p no:cacheprovider --tb=short 2>&1 | tail -30 | tee gpurun_out/pytest_gpu.txt
echo "=== smoke"; $T 300 python __graft_entry__.py --smoke 2>&1 | tail -2
echo "=== bench ours"; $T 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; grep -v "^\[rank0\]:\[W" gpurun_out/bench_n1.err | tail -c 300; python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_n1.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step", "roofline", "clocks")}); print("e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "eager", d["value_eager_launches"])
    for k, v in d["kernels"].items(): print(k, {a: b for a, b in v.items() if a != "note"})
    for k, v in d["head"].items(): print(k, v["ms_fwd_bwd"], v["frac_tc"])
    print("extras", d.get("extras")); print("reference_eager_b200", d.get("reference_eager_b200")); print("cpu_baseline", d.get("cpu_baseline"))
except Exception as e:
    print("no bench line:", e)
PY
echo "=== bench reference"; $T 600 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; head -c 600 gpurun_out/bench_ref.json; echo
echo "=== ncu fused DINO head"; bash tools/ncu_f1.sh 2>&1 | tail -70
echo "=== ncu launch list of the bench"
$T 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-ref-gpu > gpurun_out/launches_bench.log 2>&1
tail -c 300 gpurun_out/launches_bench.log; wc -l gpurun_out/launches_bench.csv
