#!/bin/bash
mkdir -p gpurun_out
L=fiveeqscm_b200/libufair_
tools/sweep.sh bb --steps 10 --warmup 3 --no-e2e --no-cpu --no-lit --no-configs -- ${L}base.so ${L}nopatch.so ${L}uorder.so ${L}nopatch_uorder.so ${L}base.so 2>&1 | tee gpurun_out/bb_sweep.txt
