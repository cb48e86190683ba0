#!/bin/bash
mkdir -p gpurun_out
( time python -m pytest tests -q -m gpu > gpurun_out/tests16.txt 2>&1 ) 2> gpurun_out/tests16.time; tail -4 gpurun_out/tests16.txt; cat gpurun_out/tests16.time
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu --no-lit --no-configs > gpurun_out/plain16.json 2> gpurun_out/plain16.err; python -c "
import json; d=json.load(open('gpurun_out/plain16.json')); print(d['kernel_ms_per_launch'], d['ms_per_step'], d['roofline']['frac'])"
