"""Experiment: e2e of the scenario-input / statistics-only feed (bench.py `e2e_scenario_inputs`) against the chunk size
of the host pipeline, and the fixed per-call overhead (a 128-member call)."""
import time

import torch

from fiveeqscm_b200 import concentrations as conc, params as P

n_t, Me = 736, 786432
spec = conc.HistSpec()
gp, tp, esc, idx = P.sample_on_device(Me, 20261018, n_scen=4, dense_pools=True)
pin = lambda x: x.cpu().contiguous().pin_memory()
scen_h = torch.from_numpy(P.scenario_emissions(n_t)).pin_memory()
gph, tph, esch = pin(gp), pin(tp), pin(esc)
idxh = torch.empty(idx.shape, dtype=torch.int32, pin_memory=True); idxh.copy_(idx)
del gp, tp, esc, idx


def timeit(M, chunk, reps=5):
    ws = conc.Workspace(0, chunk)
    out = conc.pinned_result(3, n_t, M, outputs=(), stats=spec, return_state=False)
    call = lambda: conc.run_ensemble(scen_h, gph[..., :M], tph[..., :M], scen_idx=idxh.numpy()[:M], e_scale=esch[..., :M], stats=spec,
                                     outputs=(), workspace=ws, out=out, return_state=False)
    call(); call()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        call()
    dt = (time.perf_counter() - t0) / reps
    ws.close()
    return dt


print("fixed overhead, 128 members: %.3f ms" % (1e3 * timeit(128, 128)))
for chunk in (32768, 65536, 131072, 196608, 262144, 393216, 786432):
    dt = timeit(Me, chunk)
    print("chunk %7d: %.2f ms per call -> %.3e member-steps/s" % (chunk, 1e3 * dt, Me * n_t / dt))
