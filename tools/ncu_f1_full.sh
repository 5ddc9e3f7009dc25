#!/bin/bash
# `ncu --set full` captures of the fused DINO head's kernels (one launch each); raw pages exported as CSV on the box
mkdir -p gpurun_out/ncu_raw
rm -f gpurun_out/*.ncu-rep gpurun_out/ncu_raw/*.csv
T="timeout -s KILL"
cap() {  # name regex skip
  $T 150 ncu --set full --clock-control none --profile-from-start off -k regex:$2 -s $3 -c 1 -o gpurun_out/$1 \
      python tools/dino_head_once.py fused > gpurun_out/$1.log 2>&1
  tail -1 gpurun_out/$1.log
}
cap dh_prep_weight_teacher dh_prep_weight_kernel 0
cap dh_wn_bwd dh_wn_bwd_kernel 0
cap dh_probs_student head_gemm_kernel 3      # launches: statistics (student), statistics (teacher), Q, P_s
cap dh_dw_gemm_tn "gemm_bwd_kernel<1" 0
for r in gpurun_out/*.ncu-rep; do
  ncu -i $r --page raw --csv > gpurun_out/ncu_raw/$(basename $r .ncu-rep).csv 2>/dev/null
  rm -f $r
done
ls -la gpurun_out/ncu_raw
