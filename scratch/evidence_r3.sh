#!/bin/bash
# final evidence of the round (one GPU): tests, smoke, both bench arms, the other workloads
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 600 python -m pytest tests -q -m gpu > $O/r3_pytest_gpu.log 2>&1; tail -2 $O/r3_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r3_smoke.log 2>&1; tail -6 $O/r3_smoke.log
timeout 600 python bench.py > $O/r3_bench.json 2> $O/r3_bench.err; tail -c 600 $O/r3_bench.json
timeout 600 python bench.py --impl reference > $O/r3_bench_reference_arm.json 2> $O/r3_bench_reference_arm.err; tail -c 400 $O/r3_bench_reference_arm.json
for w in fog weargait_async weargait_relaxed scaled; do
  timeout 300 python bench.py --workload $w --no-cpu-baseline --no-sweep > $O/r3_wl_$w.json 2> $O/r3_wl_$w.err
  python -c "
import json; d=json.load(open('$O/r3_wl_$w.json')); print('$w', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline'].get('per_stream_ms'))"
done
timeout 300 python bench.py --dtype f32 --no-cpu-baseline --no-sweep > $O/r3_wl_weargait_f32.json 2> $O/r3_wl_weargait_f32.err
python -c "
import json; d=json.load(open('$O/r3_wl_weargait_f32.json')); print('weargait f32', d['value'], d['ms_per_step'])"
