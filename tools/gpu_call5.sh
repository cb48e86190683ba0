#!/bin/bash
mkdir -p gpurun_out
tools/micro/dfma_forms > gpurun_out/dfma_forms.txt 2>&1; cat gpurun_out/dfma_forms.txt
for lib in libufair_gpl1; do UFAIR_LIB=$PWD/fiveeqscm_b200/$lib.so python tools/time_plan.py --members 20000,40000,80000,160000,320000,640000 --reps 10; done > gpurun_out/small_ensembles2.txt 2>&1
python tools/time_plan.py --members 160000,320000,640000 --reps 10 >> gpurun_out/small_ensembles2.txt 2>&1
cut -c1-140 gpurun_out/small_ensembles2.txt
