#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -x -k "host_pipeline or guard or bench_contract" > gpurun_out/tests19.txt 2>&1; tail -3 gpurun_out/tests19.txt
PYTHONPATH=$PWD python tools/experiments/scenario_e2e_chunks.py 2>&1 | tee gpurun_out/scen_chunks2.txt
( time python bench.py > gpurun_out/bench_n1c.json 2> gpurun_out/bench_n1c.err ) 2> gpurun_out/bench_n1c.time
tail -c 300 gpurun_out/bench_n1c.err; grep real gpurun_out/bench_n1c.time
python - <<'P'
import json
d=json.load(open('gpurun_out/bench_n1c.json'))
print(d['value'], d['ms_per_step'], d['roofline']['frac'])
for k in ('e2e','e2e_statistics_only','e2e_scenario_inputs'):
    print(k, d[k].get('value'), d[k].get('frac_of_ceiling'), d[k].get('error'))
P
