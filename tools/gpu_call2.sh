#!/bin/bash
mkdir -p gpurun_out
python - > gpurun_out/pin_speed.txt 2>&1 <<'PY'
import time, torch
torch.cuda.init()
for gb in (1, 8):
    t0 = time.perf_counter(); x = torch.empty(gb << 30, dtype=torch.uint8, pin_memory=True); t1 = time.perf_counter()
    print("pin %d GB: %.2f s (%.2f GB/s)" % (gb, t1 - t0, gb / (t1 - t0)))
    del x
PY
for v in gplall gplall_nospec; do UFAIR_LIB=$PWD/fiveeqscm_b200/libufair_$v.so python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$v.txt 2>&1; tail -1 gpurun_out/smoke_$v.txt; done
tools/sweep.sh s2 --steps 5 --warmup 3 --no-e2e --no-cpu --no-lit -- fiveeqscm_b200/libufair.so fiveeqscm_b200/libufair_nospec.so fiveeqscm_b200/libufair_gplall.so fiveeqscm_b200/libufair_gplall_nospec.so fiveeqscm_b200/libufair_gplall_nospec_tt1.so fiveeqscm_b200/libufair_gplall_nospec_tt4.so fiveeqscm_b200/libufair_gplall_nospec_m10.so fiveeqscm_b200/libufair_gplall_nospec_m14tt1.so | tee gpurun_out/sweep2.txt
