#!/bin/bash
mkdir -p gpurun_out
tools/sweep.sh s9 --steps 5 --warmup 3 --no-e2e --no-cpu --no-lit --no-configs -- fiveeqscm_b200/libufair.so fiveeqscm_b200/libufair_tt16.so fiveeqscm_b200/libufair_tt12.so | tee gpurun_out/sweep9.txt
tools/sweep.sh s9sparse --sparse --steps 5 --warmup 3 --no-e2e --no-cpu --no-lit --no-configs -- fiveeqscm_b200/libufair.so fiveeqscm_b200/libufair_form8.so fiveeqscm_b200/libufair_form8tt4.so | tee -a gpurun_out/sweep9.txt
tools/sweep.sh s9f32 --precision f32 --steps 5 --warmup 3 --no-e2e --no-cpu --no-lit --no-configs -- fiveeqscm_b200/libufair.so | tee -a gpurun_out/sweep9.txt
python -m pytest tests -x -q -m gpu 2>&1 | tail -2
