#!/bin/bash
# ncu evidence for round 2 (one GPU): plain run first, then the launch list and one --set full capture of the three ws kernels
cd "$(dirname "$0")/.."
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-sweep"
$CMD > gpurun_out/r2_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 60 --csv --log-file gpurun_out/r2_launches_final.csv $CMD > gpurun_out/r2_ncu_launches.log 2>&1
$CMD > gpurun_out/r2_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:stream_kernel_ws -s 9 -c 3 -o gpurun_out/r2_prof_ws_final $CMD > gpurun_out/r2_ncu_full.log 2>&1
ls -la gpurun_out/r2_prof_ws_final.ncu-rep
