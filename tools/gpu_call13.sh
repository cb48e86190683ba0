#!/bin/bash
# decay_small experiment: bit-identity of each variant against the base build, then the bench sweep
mkdir -p gpurun_out
export PYTHONPATH=$PWD
for v in base sF37 s337 s777 sF3F s007 sdyn; do
  UFAIR_LIB=$PWD/fiveeqscm_b200/libufair_$v.so python tools/experiments/small_decay_check.py $v 2>&1 | tail -2
done > gpurun_out/sd_check.txt 2>&1
python tools/experiments/small_decay_check.py compare base sF37 s337 s777 sF3F s007 sdyn >> gpurun_out/sd_check.txt 2>&1
cat gpurun_out/sd_check.txt
tools/sweep.sh sd --steps 10 --warmup 3 --no-e2e --no-cpu --no-lit --no-configs -- fiveeqscm_b200/libufair_base.so fiveeqscm_b200/libufair_sF37.so fiveeqscm_b200/libufair_s337.so fiveeqscm_b200/libufair_s777.so fiveeqscm_b200/libufair_sF3F.so fiveeqscm_b200/libufair_s007.so fiveeqscm_b200/libufair_sdyn.so fiveeqscm_b200/libufair_base.so 2>&1 | tee gpurun_out/sd_sweep.txt
