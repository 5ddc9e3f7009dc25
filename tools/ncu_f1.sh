#!/bin/bash
# ncu launch lists (device time + DRAM bytes per launch) of one fused DINO-head step and of its unfused form
mkdir -p gpurun_out
T="timeout -s KILL"
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
for mode in fused unfused; do
  $T 300 ncu --metrics $M --clock-control none --profile-from-start off -c 300 --csv --log-file gpurun_out/f1_ncu_$mode.csv \
      python tools/dino_head_once.py $mode > gpurun_out/f1_ncu_$mode.log 2>&1
  tail -1 gpurun_out/f1_ncu_$mode.log
done
python tools/ncu_f1_summary.py gpurun_out/f1_ncu_fused.csv gpurun_out/f1_ncu_unfused.csv | tail -60
