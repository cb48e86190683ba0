#!/bin/bash
# final tree: the driver's own sequence (GPU tests, smoke, default bench, CPU arm)
mkdir -p gpurun_out
( time python -m pytest tests -x -q -m gpu > gpurun_out/tests22.txt 2>&1 ) 2> gpurun_out/tests22.time; tail -2 gpurun_out/tests22.txt; grep real gpurun_out/tests22.time
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
( time python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref22.json 2> gpurun_out/bench_ref22.err ) 2> gpurun_out/bench_ref22.time; grep real gpurun_out/bench_ref22.time
( time python bench.py > gpurun_out/bench_n1_22.json 2> gpurun_out/bench_n1_22.err ) 2> gpurun_out/bench_n1_22.time; grep real gpurun_out/bench_n1_22.time; tail -c 200 gpurun_out/bench_n1_22.err
python - <<'P'
import json
d=json.load(open('gpurun_out/bench_n1_22.json')); r=json.load(open('gpurun_out/bench_ref22.json'))
print(d['value'], d['ms_per_step'], d['kernel_ms_per_launch'], d['roofline']['frac'], d['clocks'])
print('e2e', d['e2e']['value'], d['e2e']['frac_of_ceiling'], 'stats', d['e2e_statistics_only']['value'], 'scen', d['e2e_scenario_inputs']['value'])
print('ref', r['value'], r['cpu_baseline']['cores'], 'ratio e2e', d['e2e']['value']/r['value'], 'ratio device', d['value']/r['value'])
P
