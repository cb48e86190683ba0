"""Experiment: do the decay_small variants give the bits of the base build?  Run once per library (UFAIR_LIB):
    UFAIR_LIB=... python tools/experiments/small_decay_check.py <tag>
writes /tmp/sd_<tag>.npz; `compare base x y ...` compares them bit for bit."""
import sys

import numpy as np


def run(tag):
    import torch
    from fiveeqscm_b200 import concentrations as api, params as P
    out = {}
    for name, M, n_t, twist in (("phys", 150_016, 96, None), ("wide", 131_072 + 37, 64, "wide"), ("sat", 120_000, 40, "sat")):
        gp, tp, esc, idx = P.sample_on_device(M, 77, dense_pools=True)
        gp = gp.contiguous().clone()
        if twist == "wide":      # alpha from 1e-3 to 1e3 across members: the fallback runs in most warps and steps
            g = torch.Generator(device="cuda").manual_seed(5)
            gp[:, 8] = gp[:, 8] * (1.0 + 0.6 * torch.randn(3, M, generator=g, device="cuda", dtype=torch.float64))
        if twist == "sat":
            gp[:, 8] = -5000.0
            gp[0, 7] = 0.002
        scen = torch.from_numpy(P.scenario_emissions(n_t)).cuda()
        E = (scen[:, :, idx.long()] * esc[:, None, :]).contiguous()
        res = api.run_ensemble(E, gp, tp.contiguous())
        torch.cuda.synchronize()
        for k in ("C", "RF", "T", "state"):
            out[name + "_" + k] = getattr(res, k).cpu().numpy()
    np.savez("/tmp/sd_%s.npz" % tag, **out)
    print(tag, "done", {k: float(np.nanmax(np.abs(v))) for k, v in out.items() if k.endswith("_T")})


def compare(base, others):
    a = np.load("/tmp/sd_%s.npz" % base)
    for o in others:
        b = np.load("/tmp/sd_%s.npz" % o)
        bad = [k for k in a.files if not np.array_equal(a[k].view(np.uint64), b[k].view(np.uint64))]
        # -0 / +0 only?
        worse = [k for k in bad if not np.array_equal(a[k], b[k], equal_nan=True)]
        print("compare", base, o, "bitwise-different:", bad, "value-different:", worse)


if __name__ == "__main__":
    if sys.argv[1] == "compare":
        compare(sys.argv[2], sys.argv[3:])
    else:
        run(sys.argv[1])
