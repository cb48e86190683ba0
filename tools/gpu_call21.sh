#!/bin/bash
mkdir -p gpurun_out
L=fiveeqscm_b200/libufair_
tools/sweep.sh kd --steps 10 --warmup 3 --no-e2e --no-cpu --no-lit --no-configs -- ${L}base.so ${L}kdyn.so ${L}base.so ${L}kdyn.so 2>&1 | tee gpurun_out/kd_sweep.txt
