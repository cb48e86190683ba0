"""Experiment (kept for the record): does the number of privatised histogram copies matter?
Measured on B200, configs[3] shard: 1 copy 38.5 ms, 4: 39.9, 16: 39.9, 64: 39.6, 256: 40.4 per step --
no contention effect, so the default stays small.  Run on a GPU box: python tools/exp_hist_copies.py"""
import sys, torch
sys.path.insert(0, '.')
import bench
from fiveeqscm_b200 import concentrations as conc
torch.cuda.set_device(0)
E, gp, tp = bench.device_ensemble(torch, 1_250_000, 736, 0, True)
for copies in (1, 4, 16, 64, 256):
    plan = conc.DevicePlan(E, gp, tp, stats=conc.HistSpec(copies=copies))
    for _ in range(2):
        plan.run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        plan.reset_stats(); plan.launch(); plan.stats_pass(); plan.finalize_stats()
    e1.record(); torch.cuda.synchronize()
    print(copies, e0.elapsed_time(e1) / 5)
    del plan
