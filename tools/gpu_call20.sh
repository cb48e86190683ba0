#!/bin/bash
# 8 GPUs: final-build lines at N = 8, 4, 2 and the CPU arm
mkdir -p gpurun_out
for n in 8 4 2; do
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2954$n bench.py --gpus $n > gpurun_out/bench_n${n}d.json 2> gpurun_out/bench_n${n}d.err ) 2> gpurun_out/bench_n${n}d.time
tail -c 300 gpurun_out/bench_n${n}d.err; grep real gpurun_out/bench_n${n}d.time
done
( time python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_refd.json 2> gpurun_out/bench_refd.err ) 2> gpurun_out/bench_refd.time; grep real gpurun_out/bench_refd.time
python - <<'P'
import json
for n in (8,4,2):
    try:
        d=json.loads([l for l in open('gpurun_out/bench_n%dd.json'%n) if l.startswith('{')][-1])
        print(n, d['value'], d['ms_per_step'], d['roofline']['frac'], d.get('multi_gpu_bitwise'), d['e2e']['value'], d['e2e']['frac_of_ceiling'], d['e2e_statistics_only']['value'], d.get('e2e_scenario_inputs',{}).get('value'))
    except Exception as e: print(n, 'FAILED', e)
d=json.load(open('gpurun_out/bench_refd.json')); print('ref', d['value'], d['cpu_baseline']['cores'])
P
