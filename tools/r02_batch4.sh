#!/bin/bash
mkdir -p gpurun_out
T="timeout -s KILL"
export PYTHONUNBUFFERED=1
echo "=== pytest -m gpu"; $T 1200 python -m pytest tests -m gpu -q -p no:cacheprovider -x 2>&1 | tail -30
echo "=== bench ours"; $T 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; grep -v "^\[rank0\]:\[W" gpurun_out/bench_n1.err | tail -c 400; python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_n1.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step", "roofline", "clocks")}); print("e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"])
    for k, v in d["kernels"].items(): print(k, {a: b for a, b in v.items() if a != "note"})
    for k, v in d["head"].items(): print(k, v["ms_fwd_bwd"], v["frac_tc"])
    print("extras", d.get("extras")); print("reference_eager_b200", d.get("reference_eager_b200")); print("cpu_baseline", d.get("cpu_baseline"))
except Exception as e:
    print("no bench line:", e)
PY
cap() {  # name regex driver-mode skip
  $T 300 ncu --set full --clock-control none --import-source on -k regex:$2 -s $4 -c 1 -o gpurun_out/$1 \
      python tools/prof_driver.py $3 3 > gpurun_out/$1.log 2>&1
}
echo "=== ncu captures"
rm -f gpurun_out/*.ncu-rep
cap pe_global_u8 gather_embed_kernel pe_global_u8 1
cap pe_local_u8 gather_embed_kernel pe_local_u8 1
cap dino_fwd dino_fwd_partial dino 1
cap dino_finish dino_finish dino 1
cap pe_bwd gemm_bwd_kernel pe_bwd 1
cap optim adamw_ema_kernel optim 1
ls -la gpurun_out/*.ncu-rep
