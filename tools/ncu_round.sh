#!/bin/bash
# Profiling pass for a round (run under gpurun, ONE GPU).  Writes everything to gpurun_out/.
#   1. launch list of the bench command (per-launch device time, cold-cache & serialised)
#   2. one `--set full` capture of each hot kernel (no source import: gpurun_out/ must stay under 64 MiB)
set -u
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
T="timeout -s KILL"
$T 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-ref-gpu > gpurun_out/launches_bench.log 2>&1
cap() {  # name regex driver-mode skip
  $T 300 ncu --set full --clock-control none -k regex:$2 -s $4 -c 1 -o gpurun_out/$1 \
      python tools/prof_driver.py $3 3 > gpurun_out/$1.log 2>&1
}
# the dominant kernel of the bench step, captured from the bench command itself (full 147-tensor list)
$T 600 ncu --set full --clock-control none -k regex:ema_multi_kernel -s 4 -c 1 -o gpurun_out/ema_bench \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-ref-gpu > gpurun_out/ema_bench.log 2>&1
cap dino_fwd dino_fwd_partial dino 1
cap dino_finish dino_finish dino 1
cap dino_bwd dino_bwd_kernel dino 1
cap pe_global_u8 gather_embed_kernel pe_global_u8 1
cap pe_local_u8 gather_embed_kernel pe_local_u8 1
cap pe_bwd gemm_bwd_kernel pe_bwd 1
cap optim adamw_ema_kernel optim 1
cap head_fwd head_gemm_kernel head_bwd 2     # launches per repetition: forward (statistics), gradient + class dots
cap head_grad head_gemm_kernel head_bwd 3
cap head_de gemm_bwd_kernel head_bwd 1
cap head_dw dw_diag_kernel head_bwd 1
# gpurun brings back at most 64 MiB: keep the raw metric pages as CSV and drop the reports (the DINO ones are 15 MB each)
mkdir -p gpurun_out/ncu_raw
for r in gpurun_out/*.ncu-rep; do
  ncu -i $r --page raw --csv > gpurun_out/ncu_raw/$(basename $r .ncu-rep).csv 2>/dev/null
  rm -f $r
done
ls -la gpurun_out/ncu_raw; du -sh gpurun_out
