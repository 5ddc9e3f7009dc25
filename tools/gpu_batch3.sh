#!/bin/bash
mkdir -p gpurun_out
T="timeout -s KILL"
run() { echo "=== $1"; shift; "$@" 2>&1 | tail -${TAILN:-4}; }
export PYTHONUNBUFFERED=1
run "head tests (defaults)" $T 300 python -m pytest tests/test_gpu_head.py -q -p no:cacheprovider
for c in cfg3 cfg4; do
  TAILN=1 run "breakdown $c" $T 300 python tools/head_breakdown.py $c | tee -a gpurun_out/head_breakdown3.jsonl
done
# one full ncu capture of the new kernels (dW fused, dE, grad, fwd) at cfg 3
cap() { $T 200 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -o gpurun_out/$1 -f \
      python tools/prof_driver.py head_bwd 3 > gpurun_out/$1.log 2>&1; }
cap head_dwf dw_fused_kernel 1
cap head_de gemm_bwd_kernel 1
cap head_grad head_gemm_kernel 3
cap head_fwd head_gemm_kernel 2
ls -la gpurun_out/head_*.ncu-rep
