#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 --no-cpu --no-lit --no-configs > gpurun_out/bench_e2e23.json 2> gpurun_out/bench_e2e23.err; tail -c 300 gpurun_out/bench_e2e23.err
python - <<'P'
import json
d=json.load(open('gpurun_out/bench_e2e23.json')); e=d['e2e']
print(e['value'], e['ceiling_value'], e['frac_of_ceiling'], e['mix_estimate_value'], e['pcie_ceiling_gbs']['d2h_alone'], e['pcie_ceiling_gbs']['h2d_alone'], e['achieved_gbs_per_rank'])
P
python -m pytest tests/test_gpu_bench_contract.py -q -m gpu 2>&1 | tail -2
