#!/bin/bash
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29658 bench.py --gpus 8 --steps 3 --warmup 3 --no-lit --no-configs > gpurun_out/bench_n8g.json 2> gpurun_out/bench_n8g.err
python - <<'P'
import json
d=json.loads([l for l in open('gpurun_out/bench_n8g.json') if l.startswith('{')][-1]); e=d['e2e']
print(json.dumps(e['pcie_ceiling_gbs'])); print(e['achieved_gbs_per_rank'], '%.3e %.3e %.3e'%(e['value'], e['ceiling_value'], e['mix_estimate_value']), e['frac_of_ceiling'], e['frac_of_mix_estimate'])
P
