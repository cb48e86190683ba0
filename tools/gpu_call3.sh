#!/bin/bash
mkdir -p gpurun_out
tools/micro/issue_model > gpurun_out/issue_model.txt 2>&1
python -m pytest tests/test_gpu_math.py tests/test_gpu_parity.py tests/test_gpu_forms.py -x -q -m gpu > gpurun_out/tests3.txt 2>&1; tail -3 gpurun_out/tests3.txt
for v in g3_tt1 g3_tt1_rc5; do UFAIR_LIB=$PWD/fiveeqscm_b200/libufair_$v.so python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$v.txt 2>&1; tail -1 gpurun_out/smoke_$v.txt; done
tools/sweep.sh s3 --steps 5 --warmup 3 --no-e2e --no-cpu --no-lit --no-configs -- fiveeqscm_b200/libufair.so fiveeqscm_b200/libufair_g3_tt1.so fiveeqscm_b200/libufair_g3_tt2.so fiveeqscm_b200/libufair_g3_tt1_rc4.so fiveeqscm_b200/libufair_g3_tt1_rc5.so fiveeqscm_b200/libufair_g3_tt1_rc5_m10.so fiveeqscm_b200/libufair_g3_tt1_sq0.so | tee gpurun_out/sweep3.txt
( time python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err ) 2> gpurun_out/bench_full.time
tail -c 600 gpurun_out/bench_full.err; cat gpurun_out/bench_full.time
