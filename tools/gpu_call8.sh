#!/bin/bash
mkdir -p gpurun_out
tools/sweep.sh s8 --steps 5 --warmup 3 --no-e2e --no-cpu --no-lit --no-configs -- fiveeqscm_b200/libufair_m8rc7.so fiveeqscm_b200/libufair_m8rc7tt2.so fiveeqscm_b200/libufair_m8rc7tt4.so fiveeqscm_b200/libufair_m8rc7tt8.so fiveeqscm_b200/libufair_m7rc7.so fiveeqscm_b200/libufair_m6rc7.so fiveeqscm_b200/libufair_m8rc6.so fiveeqscm_b200/libufair_m8rc7.so | tee gpurun_out/sweep8.txt
