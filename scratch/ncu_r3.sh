#!/bin/bash
# ncu evidence, end of round 2 (one GPU): launch list of the default bench, launch list + one --set full capture of the FoG kernels
cd "$(dirname "$0")/.."
O=gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-sweep"
$CMD > $O/r3_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 60 --csv --log-file $O/r3_launches.csv $CMD --graph 0 > $O/r3_ncu_launches.log 2>&1
FOG="python bench.py --workload fog --steps 2 --warmup 3 --no-cpu-baseline --no-sweep --graph 0"
$FOG > $O/r3_plain_fog.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 40 --csv --log-file $O/r3_launches_fog.csv $FOG > $O/r3_ncu_launches_fog.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:stream_kernel -s 12 -c 2 -o $O/r3_fog_final -f $FOG > $O/r3_ncu_full_fog.log 2>&1
ls -la $O/r3_fog_final.ncu-rep $O/r3_launches.csv $O/r3_launches_fog.csv
