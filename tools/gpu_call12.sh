#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/tests12.txt 2>&1; tail -2 gpurun_out/tests12.txt
UFAIR_FUZZ_CASES=600 python -m pytest tests/test_gpu_fuzz.py -x -q -m gpu > gpurun_out/fuzz600.txt 2>&1; tail -2 gpurun_out/fuzz600.txt
( time python bench.py > gpurun_out/bench_n1b.json 2> gpurun_out/bench_n1b.err ) 2> gpurun_out/bench_n1b.time
tail -c 300 gpurun_out/bench_n1b.err; grep real gpurun_out/bench_n1b.time
B="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-lit --no-configs"
$B > gpurun_out/plain12.json 2> gpurun_out/plain12.err && ncu --set full --clock-control none --import-source on -k regex:ufair_integrate -s 1 -c 1 -o gpurun_out/r2_final $B > gpurun_out/ncu12.log 2>&1
$B > gpurun_out/plain12b.json 2> gpurun_out/plain12b.err && ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_final_launches.csv $B > gpurun_out/ncu12b.log 2>&1
tail -2 gpurun_out/ncu12.log
