#!/bin/bash
mkdir -p gpurun_out
tools/sweep.sh s7 --steps 5 --warmup 3 --no-e2e --no-cpu --no-lit --no-configs -- fiveeqscm_b200/libufair.so fiveeqscm_b200/libufair_m10.so fiveeqscm_b200/libufair_m10rc7.so fiveeqscm_b200/libufair_m9rc7.so fiveeqscm_b200/libufair_m8rc7.so fiveeqscm_b200/libufair_m11.so fiveeqscm_b200/libufair_m10tt2.so | tee gpurun_out/sweep7.txt
tools/sweep.sh s7b --steps 5 --warmup 3 --no-e2e --no-cpu --no-lit --no-configs -- fiveeqscm_b200/libufair.so fiveeqscm_b200/libufair_m10.so | tee -a gpurun_out/sweep7.txt
python -m pytest tests/test_gpu_math.py -x -q -m gpu 2>&1 | tail -2
