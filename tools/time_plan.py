#!/usr/bin/env python
"""Integrator launch time for a list of ensemble sizes (perf experiments; UFAIR_LIB picks the library variant).
    python tools/time_plan.py --members 5000,10000,20000 [--alpha newton --newton-iters 3] [--n-t 736]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--members", default="10000")
    ap.add_argument("--n-t", type=int, default=736)
    ap.add_argument("--alpha", default="exp")
    ap.add_argument("--newton-iters", type=int, default=0)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--precision", default="f64")
    a = ap.parse_args()
    import torch
    from fiveeqscm_b200 import concentrations as conc
    for M in [int(x) for x in a.members.split(",")]:
        E, gp, tp = bench.device_ensemble(torch, M, a.n_t, 0, True)
        plan = conc.DevicePlan(E, gp, tp, stats=conc.HistSpec(), alpha_mode=a.alpha, newton_iters=a.newton_iters,
                               precision=a.precision)
        kms, sms = bench.time_launches(torch, plan, a.reps, warm=3)
        print("lib=%s M=%d alpha=%s K=%d: kernel %.4f ms, step %.4f ms, %.3e member-steps/s, variant %s" % (
            os.path.basename(os.environ.get("UFAIR_LIB", "libufair.so")), M, a.alpha, a.newton_iters, kms, sms,
            M * a.n_t / (kms * 1e-3), bench.variant_of(plan)), flush=True)
        del plan, E, gp, tp
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
