"""Experiment: CUPTI timeline (torch.profiler) of one scenario-input / statistics-only host-pipeline call."""
import sys
import time

import torch
from torch.profiler import ProfilerActivity, profile

from fiveeqscm_b200 import concentrations as conc, params as P

n_t, Me = 736, 786432
chunk = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
spec = conc.HistSpec()
gp, tp, esc, idx = P.sample_on_device(Me, 20261018, n_scen=4, dense_pools=True)
pin = lambda x: x.cpu().contiguous().pin_memory()
scen_h = torch.from_numpy(P.scenario_emissions(n_t)).pin_memory()
gph, tph, esch = pin(gp), pin(tp), pin(esc)
idxh = torch.empty(idx.shape, dtype=torch.int32, pin_memory=True); idxh.copy_(idx)
del gp, tp, esc, idx
ws = conc.Workspace(0, chunk)
out = conc.pinned_result(3, n_t, Me, outputs=(), stats=spec, return_state=False)
call = lambda: conc.run_ensemble(scen_h, gph, tph, scen_idx=idxh.numpy(), e_scale=esch, stats=spec, outputs=(), workspace=ws, out=out,
                                 return_state=False)
call(); call()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    t0 = time.perf_counter()
    call()
    wall = time.perf_counter() - t0
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
base = ev[0].time_range.start
print("wall %.2f ms, %d device events" % (1e3 * wall, len(ev)))
for e in ev:
    print("%9.3f ms  +%8.3f ms  %s" % ((e.time_range.start - base) / 1e3, (e.time_range.end - e.time_range.start) / 1e3, e.name[:70]))
cpu = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CPU and (e.time_range.end - e.time_range.start) > 200]
cpu.sort(key=lambda e: e.time_range.start)
for e in cpu[:40]:
    print("cpu %9.3f ms  +%8.3f ms  %s" % ((e.time_range.start - base) / 1e3, (e.time_range.end - e.time_range.start) / 1e3, e.name[:70]))
