#!/bin/bash
mkdir -p gpurun_out
{ nproc; free -g | head -2; nvidia-smi topo -m | head -12; ls /sys/devices/system/node/ | tr '\n' ' '; } > gpurun_out/topo8.txt 2>&1
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err ) 2> gpurun_out/bench_n8.time
tail -c 600 gpurun_out/bench_n8.err; cat gpurun_out/bench_n8.time
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 4 > gpurun_out/bench_n4.json 2> gpurun_out/bench_n4.err ) 2> gpurun_out/bench_n4.time
cat gpurun_out/bench_n4.time
