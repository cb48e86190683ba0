#!/bin/bash
mkdir -p gpurun_out
tools/micro/issue_model > gpurun_out/issue_model2.txt 2>&1
python -m pytest tests -x -q -m gpu > gpurun_out/tests4.txt 2>&1; tail -3 gpurun_out/tests4.txt
for lib in libufair libufair_big; do UFAIR_LIB=$PWD/fiveeqscm_b200/$lib.so python tools/time_plan.py --members 5000,10000,20000,40000,80000 --reps 20; done > gpurun_out/small_ensembles.txt 2>&1
for lib in libufair libufair_nk3; do UFAIR_LIB=$PWD/fiveeqscm_b200/$lib.so python tools/time_plan.py --members 1250000 --alpha newton --newton-iters 3 --reps 4; done > gpurun_out/newton.txt 2>&1
cat gpurun_out/small_ensembles.txt gpurun_out/newton.txt
B="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-lit --no-configs"
$B > gpurun_out/plain4.json 2> gpurun_out/plain4.err && ncu --set full --clock-control none --import-source on -k regex:ufair_integrate -s 1 -c 1 -o gpurun_out/r2_gplall $B > gpurun_out/ncu4.log 2>&1
$B > gpurun_out/plain4b.json 2> gpurun_out/plain4b.err && ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_launches.csv $B > gpurun_out/ncu4b.log 2>&1
tail -3 gpurun_out/ncu4.log
