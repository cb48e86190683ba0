#!/bin/bash
# 2 GPUs: new GPU tests, N=2 bench (reducer, canary, link ceiling), N=1 full bench, reference arm
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/tests10.txt 2>&1; tail -3 gpurun_out/tests10.txt
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err ) 2> gpurun_out/bench_n2.time
tail -c 400 gpurun_out/bench_n2.err; cat gpurun_out/bench_n2.time
( time python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err ) 2> gpurun_out/bench_n1.time
tail -c 400 gpurun_out/bench_n1.err; cat gpurun_out/bench_n1.time
( time python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err ) 2> gpurun_out/bench_ref.time
cat gpurun_out/bench_ref.time
