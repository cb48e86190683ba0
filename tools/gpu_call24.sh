#!/bin/bash
# final lines for profiles/: N = 8, 4, 2, 1 with the final bench.py
mkdir -p gpurun_out
for n in 8 4 2; do
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2964$n bench.py --gpus $n > gpurun_out/bench_n${n}f.json 2> gpurun_out/bench_n${n}f.err ) 2> gpurun_out/bench_n${n}f.time
grep real gpurun_out/bench_n${n}f.time
done
( time python bench.py > gpurun_out/bench_n1f.json 2> gpurun_out/bench_n1f.err ) 2> gpurun_out/bench_n1f.time; grep real gpurun_out/bench_n1f.time
python - <<'P'
import json
for n in (8,4,2,1):
    try:
        d=json.loads([l for l in open('gpurun_out/bench_n%df.json'%n) if l.startswith('{')][-1]); e=d['e2e']
        print(n, '%.3e'%d['value'], round(d['ms_per_step'],2), round(d['roofline']['frac'],3), d.get('multi_gpu_bitwise'), '%.3e'%e['value'], round(e['frac_of_ceiling'],2), round(e['frac_of_mix_estimate'],2), '%.3e'%d['e2e_statistics_only']['value'], '%.3e'%d['e2e_scenario_inputs']['value'])
    except Exception as ex: print(n, 'FAILED', ex)
P
