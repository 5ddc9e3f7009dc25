#!/bin/bash
# re-capture of the patch-embed kernels and the bench launch list only
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/launches_bench.log 2>&1
for m in pe_global pe_local pe_global_u8; do
  ncu --set full --clock-control none --import-source on -k regex:gather_embed_kernel -s 1 -c 1 -f -o gpurun_out/$m \
      python tools/prof_driver.py $m 3 > gpurun_out/$m.log 2>&1
done
ls -la gpurun_out/pe_*.ncu-rep
