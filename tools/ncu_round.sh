#!/bin/bash
# Profiling pass for a round (run under gpurun, ONE GPU).  Writes everything to gpurun_out/.
#   1. launch list of the bench command (per-launch device time, cold-cache & serialised)
#   2. one `--set full` capture of each hot kernel
set -u
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/launches_bench.log 2>&1
cap() {  # name regex driver-mode skip
  ncu --set full --clock-control none --import-source on -k regex:$2 -s $4 -c 1 -o gpurun_out/$1 \
      python tools/prof_driver.py $3 3 > gpurun_out/$1.log 2>&1
}
# the dominant kernel of the bench step, captured from the bench command itself (full 147-tensor list)
ncu --set full --clock-control none --import-source on -k regex:ema_multi_kernel -s 4 -c 1 -o gpurun_out/ema_bench \
    python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ema_bench.log 2>&1
cap ema ema_multi_kernel ema 1
cap dino_fwd dino_fwd_partial dino 1
cap dino_bwd dino_bwd_kernel dino 1
cap pe_global gather_embed_kernel pe_global 1
cap pe_local gather_embed_kernel pe_local 1
cap pe_global_u8 gather_embed_kernel pe_global_u8 1
cap dino_finish dino_finish dino 1
cap head_fwd head_gemm_kernel head_bwd 2     # launches per repetition: <256,0,pair> (forward), <256,2,pair> (grad logits)
cap head_grad head_gemm_kernel head_bwd 3
cap head_de gemm_bwd_kernel head_bwd 2       # <0,4> (dE, split-K, W_hat multicast), <1,1> (dW)
cap head_dw gemm_bwd_kernel head_bwd 3
rm -f gpurun_out/head_dwf.ncu-rep gpurun_out/head_dwf.log
ls -la gpurun_out/*.ncu-rep
