#!/bin/bash
mkdir -p gpurun_out
L=fiveeqscm_b200/libufair_
tools/sweep.sh fs --steps 10 --warmup 3 --no-e2e --no-cpu --no-lit --no-configs -- ${L}base.so ${L}fF13.so ${L}f313.so ${L}fFFF.so ${L}base.so 2>&1 | tee gpurun_out/fs_sweep.txt
