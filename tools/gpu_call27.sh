#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --no-lit --no-configs > gpurun_out/b27.json 2> gpurun_out/b27.err; tail -c 300 gpurun_out/b27.err
python -c "
import json; d=json.load(open('gpurun_out/b27.json')); print(d['roofline']['frac'], d['roofline'].get('clock_note'), d['clocks'])"
python -m pytest tests/test_gpu_bench_contract.py -q -m gpu 2>&1 | tail -1
