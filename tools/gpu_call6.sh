#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
tools/sweep.sh s6 --steps 5 --warmup 3 --no-e2e --no-cpu --no-lit --no-configs -- fiveeqscm_b200/libufair.so fiveeqscm_b200/libufair_ph_rc4.so fiveeqscm_b200/libufair_ph_rc0.so fiveeqscm_b200/libufair_ph_tt2.so fiveeqscm_b200/libufair_ph_m10.so fiveeqscm_b200/libufair_ph_m14.so fiveeqscm_b200/libufair_ph_m16.so fiveeqscm_b200/libufair_ph_m16rc5.so fiveeqscm_b200/libufair_ph_m16rc4.so fiveeqscm_b200/libufair_m2.so | tee gpurun_out/sweep6.txt
python -m pytest tests/test_gpu_parity.py tests/test_gpu_forms.py tests/test_gpu_inverse.py -x -q -m gpu 2>&1 | tail -2
UFAIR_LIB=$PWD/fiveeqscm_b200/libufair_m2.so python -m pytest tests/test_gpu_math.py -x -q -m gpu 2>&1 | tail -15
UFAIR_LIB=$PWD/fiveeqscm_b200/libufair_m2.so python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -1
