#!/bin/bash
# final 1-GPU pass of round 2: full GPU suite, smoke, bench (ours + reference arm), the other SSL configs, ncu pass
mkdir -p gpurun_out
T="timeout -s KILL"
export PYTHONUNBUFFERED=1
echo "=== pytest -m gpu"; $T 1200 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -12
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
echo "=== bench reference"; $T 600 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; head -c 700 gpurun_out/bench_ref.json; echo
echo "=== other SSL configs"
for c in cfg1_L8 cfg4 cfg0; do
  EXTRA="--no-cpu --no-ref-gpu"; if [ $c == cfg0 ]; then EXTRA=""; fi
  $T 600 python bench.py --config $c --steps 50 --warmup 5 $EXTRA > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_$c.json").read().strip().splitlines()[-1])
    print("$c", {k: d[k] for k in ("value", "ms_per_step")}, "e2e", d["e2e"]["value"], "ref_eager", (d.get("reference_eager_b200") or {}).get("ssl_step", {}).get("faces_per_s"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
except Exception as e:
    print("$c no bench line:", e)
PY
done
echo "=== ncu round"; bash tools/ncu_round.sh 2>&1 | tail -18
