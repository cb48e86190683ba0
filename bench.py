#!/usr/bin/env python
"""bench.py -- headline benchmark of the Universal-FaIR ensemble integrator on B200.

Metric (BASELINE.json): ensemble member-timesteps/sec.  Workload: one GPU's shard of configs[3]
(10^7-member x 736-year 3-gas ensemble over 8 GPUs = 1.25e6 members per GPU), FP64, per-member
emissions resident in HBM, full C/RF/T output to HBM, per-step temperature histogram + moments,
NCCL all-reduce of the statistics when N > 1.  Weak scaling: per-GPU work is fixed.

    python bench.py [--gpus N] [--steps K] [--warmup W]            our arm
    python bench.py --impl reference ...                             the CPU arm (oracle on host cores)
    torchrun --nproc-per-node N bench.py --gpus N ...                N > 1

One JSON line on stdout (rank 0).  `value` = device-resident throughput; `e2e` = the same metric
through the public host-buffer API (pinned host inputs -> H2D -> kernel -> D2H of all outputs);
`roofline` = the integrate kernel against the FP64-pipe peak measured in this run (and the HBM
view); `cpu_baseline` = the C oracle timed on this box's cores on a bounded sample.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

N_GAS = 3
# algorithmic work per member-step (BASELINE.md section 5 / DESIGN.md): 3 gases
#   bytes: 3 E in + 3 C + 3 RF + 1 T out = 10 words (80 B in f64, 40 B in f32)
FLOPS_PER_STEP = 751.0               # 163 simple + 15 exp(28) + 3 log(36) + 3 sqrt(10) + 3 div(10)
METRIC = "ensemble member-timesteps/sec"


def algorithmic_flops(gas_form):
    """SURVEY.md 8(d) convention (exp = 28, log = 36, sqrt = 10, div = 10 FP64 flops), per member-step,
    for the pools and forcing terms the parameters actually use: per gas 18 + 8 n_pool simple flops,
    1 + n_pool exponentials, one division, a log / sqrt where that term exists; 13 for the thermal
    step.  Four pools and all three terms in every gas give the survey's 751."""
    total = 13.0
    for f in gas_form:
        n_pool = (f & 7) or 4
        terms = (f >> 4) & 7 or 7
        total += 18 + 8 * n_pool + 28 * (1 + n_pool) + 10 + (36 if terms & 1 else 0) + (10 if terms & 4 else 0)
    return total


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--members-per-gpu", type=int, default=1_250_000)
    ap.add_argument("--n-t", type=int, default=736)
    ap.add_argument("--precision", default="f64", choices=["f64", "f32"])
    ap.add_argument("--e2e-members", type=int, default=262_144)
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--e2e-chunk", type=int, default=16384, help="members per chunk of the host pipeline")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="target CPU work for the cpu_baseline sample")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-lit", action="store_true", help="skip the secondary literature-parameter measurement")
    ap.add_argument("--sparse", action="store_true", help="literature-style sparse parameters (1-pool CH4/N2O)")
    ap.add_argument("--general-kernel", action="store_true", help="experiment: never pick a specialised per-gas form")
    ap.add_argument("--fext", action="store_true", help="experiment: add a shared external-forcing series")
    ap.add_argument("--iirf-max", type=float, default=None, help="experiment: switch the iIRF ceiling on (general kernel)")
    ap.add_argument("--no-stats", action="store_true", help="experiment: integrate without histogram/moments")
    ap.add_argument("--outputs", default="C,RF,T", help="experiment: comma list of outputs written to HBM ('' = none)")
    return ap.parse_args()


def workload_config(args, n_gpus):
    return {
        "workload": "configs[3] shard: 3-gas (CO2/CH4/N2O) x %d steps x %d members per GPU (x%d GPUs = %d members), "
                    "per-member emissions, full C/RF/T output + per-step T histogram (1024 bins) and moments"
                    % (args.n_t, args.members_per_gpu, n_gpus, args.members_per_gpu * n_gpus),
        "members_per_gpu": args.members_per_gpu, "n_t": args.n_t, "n_gas": N_GAS, "dt_years": 1.0,
        "alpha_mode": "exp", "t_mode": "mid", "parameters": "sparse" if args.sparse else "dense (4 active pools, 3 forcing terms per gas)",
        "sharding": "member axis, %d rank(s), no data-path collective; one all-reduce of histogram+moments" % n_gpus,
        "l2": "inputs (%.1f GB per step) exceed L2; no explicit flush" % (N_GAS * args.n_t * args.members_per_gpu * 8 / 1e9),
    }


# ------------------------------------------------------------------------------------------------
# synthetic inputs on the device (same recipe as fiveeqscm_b200.params, torch RNG)
# ------------------------------------------------------------------------------------------------
def device_ensemble(torch, M, n_t, rank, dense):
    """Synthetic inputs of one rank, generated on its GPU by the product's own sampler
    (ufair_sample_f64: Philox stream keyed by the GLOBAL member index, so the N ranks together hold
    the one ensemble a single GPU would generate): perturbed parameters, a scenario index and an
    emission scale per member; the per-member emission rows are the scenario rows times the scale."""
    from fiveeqscm_b200 import params as P
    dev = torch.device("cuda", torch.cuda.current_device())
    gp, tp, scale, idx = P.sample_on_device(M, 20261018, first_member=rank * M, n_scen=4, dense_pools=dense)
    scen = torch.from_numpy(P.scenario_emissions(n_t)).to(dev)                  # [3][n_t][4]
    E = scen[:, :, idx.long()]                                                    # [3][n_t][M]
    E *= scale[:, None, :]
    return E.contiguous(), gp.contiguous(), tp.contiguous()


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "power_w_max": None, "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 7:
                continue
            try:
                sm.append(float(c[0])); mx.append(float(c[1])); pw.append(float(c[2]))
            except ValueError:
                continue
            for nme, v in zip(names, c[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        self.f.close()
        os.unlink(self.f.name)
        if sm:
            busy = [s for s, p in zip(sm, pw) if p > 0.5 * max(pw)] or sm
            out.update(sm_mhz=statistics.median(busy), sm_max_mhz=max(mx), reasons=sorted(reasons),
                       power_w_max=max(pw), samples=len(sm))
        return out


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle on the host cores (the only place bench.py executes oracle/)
# ------------------------------------------------------------------------------------------------
def host_threads():
    """Host threads the CPU arm may use: the cores this process can run on (torchrun exports
    OMP_NUM_THREADS=1, which is about the GPU ranks, not about this leg)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_oracle_rate(E, gp, tp, target_seconds, threads=0):
    """member-steps/s of the C oracle on a bounded sample (first members of the workload)."""
    from oracle import c_oracle as co
    n_thr = host_threads() if threads <= 0 else threads
    n_t = E.shape[1]
    probe = min(E.shape[2], 64 * n_thr)
    t0 = time.perf_counter()
    co.oxfair(E[:, :, :probe], gp[:, :, :probe], tp[:, :probe], outputs=("C", "RF", "T"), n_threads=n_thr)
    dt0 = max(time.perf_counter() - t0, 1e-6)
    n = int(min(E.shape[2], max(probe, probe * target_seconds / dt0)))
    t0 = time.perf_counter()
    out = co.oxfair(E[:, :, :n], gp[:, :, :n], tp[:, :n], outputs=("C", "RF", "T"), n_threads=n_thr)
    dt1 = time.perf_counter() - t0
    return n * n_t / dt1, out["threads"], n, dt1


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from fiveeqscm_b200 import params as P
    from oracle import c_oracle as co
    n_thr = host_threads()
    n = max(256, 128 * n_thr)
    ens = P.sample_ensemble(n, n_t=args.n_t, dense_pools=not args.sparse)
    E = P.member_emissions(ens["scen"], ens["scen_idx"], ens["e_scale"])
    gp, tp = ens["gas_params"], ens["thermal_params"]
    # size each step at ~2 s of work
    t0 = time.perf_counter()
    co.oxfair(E, gp, tp, n_threads=n_thr)
    dt0 = time.perf_counter() - t0
    reps = max(1, int(2.0 / max(dt0, 1e-3)))
    def step():
        for _ in range(reps):
            out = co.oxfair(E, gp, tp, n_threads=n_thr)
            h, m = co.temperature_stats(out["T"], -5.0, 25.0, 1024)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    el = time.perf_counter() - t0
    work = args.steps * reps * n * args.n_t
    val = work / el
    sample = "%d members x %d steps x %d repeats per step (C oracle, OpenMP, incl. histogram)" % (n, args.n_t, reps)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "member-timesteps/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * el / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, args.gpus),
            "cpu_baseline": {"value": val, "unit": "member-timesteps/s", "cores": n_thr, "kind": "port", "sample": sample,
                             "host_cpus": os.cpu_count()},
            "e2e": {"value": val, "unit": "member-timesteps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "note": "the reference ships no implementation of the 5-equation loop (SURVEY.md 0); this arm times the "
                    "in-repo C restatement (oracle/ufair_oracle.c) on the host cores"}
    emit(line)


# ------------------------------------------------------------------------------------------------
_REAL_STDOUT = None


def claim_stdout():
    """Keep stdout for the ONE JSON line: everything else any library prints to fd 1 (NCCL's version
    banner, for one) is sent to stderr for the life of the process."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    claim_stdout()
    args = parse()
    if args.impl == "reference":
        run_reference(args)
        return
    import torch
    import torch.distributed as dist
    from fiveeqscm_b200 import _abi, concentrations as conc, dist as D

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py (our arm) needs a GPU; there is no CPU fallback"
    torch.cuda.set_device(local)
    # one rank per GPU: keep each rank (and the pinned host buffers it is about to allocate) on its GPU's
    # NUMA node.  Not at N = 1, where the CPU-baseline leg wants every core of the box.
    numa = D.bind_to_gpu_numa_node(local) if world > 1 else None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n_gpus = world
    L = _abi.lib()
    M, n_t = args.members_per_gpu, args.n_t

    E, gp, tp = device_ensemble(torch, M, n_t, rank, dense=not args.sparse)
    spec = None if args.no_stats else conc.HistSpec()
    outs = tuple(o for o in args.outputs.split(",") if o)
    fx = None
    if args.fext:
        yr = torch.arange(n_t, device=E.device, dtype=torch.float64)
        fx = 0.1 * torch.sin(2.0 * 3.141592653589793 * yr / 11.0)
    plan = conc.DevicePlan(E, gp, tp, stats=spec, precision=args.precision, outputs=outs, f_ext=fx,
                           iirf_max=args.iirf_max, gas_form=None if args.general_kernel else "auto")
    vform, vgpl, vmw = plan.kernel_variant()
    flops_step = algorithmic_flops(plan.gas_form if not args.general_kernel else (0,) * N_GAS)
    assert args.sparse or flops_step == FLOPS_PER_STEP
    res = plan.result
    if args.precision == "f32":
        pass  # DevicePlan converted the inputs; E/gp/tp (f64) are only kept for the CPU sample

    ev = lambda: torch.cuda.Event(enable_timing=True)
    kern_ev = [(ev(), ev()) for _ in range(args.steps)]

    def step(i=None):
        plan.reset_stats()
        if i is not None:
            kern_ev[i][0].record()
        plan.launch()                      # the fused integrator (+ in-loop histogram): the dominant kernel
        if i is not None:
            kern_ev[i][1].record()
        plan.stats_pass()                     # second statistics pass over the T rows
        plan.finalize_stats()
        if world > 1 and spec is not None:
            D.allreduce_stats(res.hist, res.moments)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 0)):
        step()
    barrier()
    clocks = ClockSampler(local) if rank == 0 else None
    e0, e1 = ev(), ev()
    e0.record()
    for i in range(args.steps):
        step(i)
    e1.record()
    barrier()
    clk = clocks.stop() if clocks else None
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
    kms = torch.tensor([sum(a.elapsed_time(b) for a, b in kern_ev) / args.steps], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(kms, op=dist.ReduceOp.MAX)
    total_ms, kernel_ms = float(ms.item()), float(kms.item())
    ms_per_step = total_ms / args.steps
    units_per_step = float(M) * n_t * n_gpus
    value = units_per_step / (ms_per_step * 1e-3)

    # sanity on the result of the last step (cheap, outside the timed region)
    if spec is not None:
        assert bool((res.hist.sum(dim=1) == M * n_gpus).all()), "histogram rows must count every member"
    if res.T is not None:
        assert bool(torch.isfinite(res.T[-1]).all())

    line = {"metric": METRIC, "value": value, "unit": "member-timesteps/s", "n_gpus": n_gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": args.precision, "data": "synthetic", "config": workload_config(args, n_gpus),
            # per step: stats_reset, ufair_integrate_kernel, moments_pass, stats_finalize (1 without statistics)
            "gpu_launches": (4 if spec is not None else 1) * args.steps, "kernel_ms_per_launch": kernel_ms}

    if rank == 0:
        line["clocks"] = clk
        # ---- roofline of the dominant kernel (ufair_integrate_kernel), measured live
        es = 8 if args.precision == "f64" else 4
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        hbm_src = "MEASURED_PEAKS.json (measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        # FMA-pipe peak, measured here: the first launch is the BURST figure (cold, full clock); the
        # timed region above is a seconds-long step under the power cap, so the denominator is the
        # SUSTAINED figure: the same microbenchmark back to back for ~1 s, median of the second half
        msd, fl = __import__("ctypes").c_double(), __import__("ctypes").c_double()
        peak_fn = L.ufair_peak_fp64 if args.precision == "f64" else L.ufair_peak_fp32
        iters = 400_000 if args.precision == "f64" else 800_000
        rates = []
        for _ in range(10):   # each call = warm-up launch + timed launch, ~53 ms apiece
            _abi.check(peak_fn(iters, msd, fl, None))
            rates.append(fl.value / (msd.value * 1e-3) / 1e12)
        fpeak_burst = rates[0]
        fpeak = statistics.median(rates[5:])
        bytes_per_launch = (N_GAS + 2 * N_GAS + 1) * es * float(M) * n_t
        flops_per_launch = flops_step * float(M) * n_t
        a_hbm = bytes_per_launch / (kernel_ms * 1e-3) / 1e9
        a_fp = flops_per_launch / (kernel_ms * 1e-3) / 1e12
        traffic = None  # DRAM bytes per launch, scaled from the committed ncu --set full capture
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(args.precision)
            if tj and spec is not None and outs == ("C", "RF", "T") and not args.sparse:
                traffic = tj["bytes_per_member_step"] * float(M) * n_t
        except Exception:
            pass
        t_hbm, t_fp = bytes_per_launch / (hbm_peak * 1e9), flops_per_launch / (fpeak * 1e12)
        bound = "fp64" if args.precision == "f64" else "fp32"
        # FP32 mode evaluates its exponentials / log / sqrt / reciprocals on the MUFU pipe, so the FMA-flop
        # convention does not describe it; its binding roofline is HBM
        if t_hbm > t_fp or args.precision == "f32":
            roof = {"bound": "hbm", "achieved": a_hbm, "peak": hbm_peak, "unit": "GB/s", "frac": a_hbm / hbm_peak}
        else:
            roof = {"bound": bound, "achieved": a_fp, "peak": fpeak, "unit": "TFLOP/s", "frac": a_fp / fpeak}
        roof.update({
            "traffic": traffic, "kernel": "ufair_integrate_kernel<%s,3,EXP>" % ("double" if es == 8 else "float"),
            "kernel_ms": kernel_ms, "kernel_share_of_step": kernel_ms / ms_per_step,
            "algorithmic_bytes_per_member_step": (N_GAS + 2 * N_GAS + 1) * es,
            "algorithmic_flops_per_member_step": flops_step,
            "kernel_variant": {"form": list(vform), "gases_per_lane": vgpl, "members_per_warp": vmw,
                               "loop": plan.loop_variant},
            "hbm": {"achieved": a_hbm, "peak": hbm_peak, "unit": "GB/s", "frac": a_hbm / hbm_peak, "peak_source": hbm_src},
            bound: {"achieved": a_fp, "peak": fpeak, "unit": "TFLOP/s", "frac": a_fp / fpeak,
                    "peak_burst": fpeak_burst, "frac_of_burst": a_fp / fpeak_burst,
                    "peak_source": "FMA microbenchmark in this run (ufair_peak_%s): sustained = median of launches "
                                   "6-10 of ten back-to-back ~2x%.0f ms calls, burst = the first" % (bound, msd.value)},
        })
        line["roofline"] = roof

    # ---- the same shard with literature-style (sparse) parameters: one-pool CH4 / N2O, one forcing term
    # per gas, on the specialised kernel the library picks for them.  Secondary figure; the headline
    # above keeps the dense parameters, where nothing can be skipped.
    if (rank == 0 and world == 1 and not args.sparse and not args.general_kernel and spec is not None
            and not args.fext and args.iirf_max is None and not args.no_lit):
        try:
            del plan, res
            torch.cuda.empty_cache()
            from fiveeqscm_b200 import params as P
            gp2, tp2, _, _ = P.sample_on_device(M, 20261018, dense_pools=False, precision="f64")
            plan2 = conc.DevicePlan(E, gp2.contiguous(), tp2.contiguous(), stats=spec, precision=args.precision, outputs=outs)
            f2, g2, m2 = plan2.kernel_variant()
            for _ in range(3):
                plan2.reset_stats(); plan2.launch()
            k0, k1 = ev(), ev()
            n2 = 5
            k0.record()
            for _ in range(n2):
                plan2.launch()
            k1.record()
            torch.cuda.synchronize()
            ms2 = k0.elapsed_time(k1) / n2
            fl2 = algorithmic_flops(plan2.gas_form)
            line["literature_parameters"] = {
                "what": "same shard and outputs, CH4 / N2O with one pool and the sqrt term only, CO2 four pools and "
                        "the log term: integrator launches only",
                "kernel_ms": ms2, "value": float(M) * n_t / (ms2 * 1e-3), "unit": "member-timesteps/s",
                "kernel_variant": {"form": list(f2), "gases_per_lane": g2, "members_per_warp": m2},
                "algorithmic_flops_per_member_step": fl2,
                "frac_of_fp64_peak": fl2 * float(M) * n_t / (ms2 * 1e-3) / 1e12 / line["roofline"][bound]["peak"]
                if bound in line["roofline"] else None}
            del plan2, gp2, tp2
            torch.cuda.empty_cache()
        except Exception as exc:  # never lose the headline line to the secondary figure
            line["literature_parameters"] = {"error": repr(exc)}

    # ---- e2e: the public host-buffer API, pinned host inputs, H2D + kernel + D2H of every output
    if not args.no_e2e:
        Me = min(args.e2e_members, M)
        hdt = torch.float64 if args.precision == "f64" else torch.float32
        pin = lambda x: x.to(hdt).cpu().pin_memory()   # host arrays already in the run's precision
        Eh, gph, tph = pin(E[:, :, :Me]), pin(gp[:, :, :Me]), pin(tp[:, :Me])
        outs = ("C", "RF", "T")
        out = conc.pinned_result(N_GAS, n_t, Me, outputs=outs, stats=spec, precision=args.precision)
        ws = conc.Workspace(local, args.e2e_chunk)
        call = lambda: conc.run_ensemble(Eh, gph, tph, stats=spec, outputs=outs, precision=args.precision,
                                         workspace=ws, out=out)
        call()  # warm-up: staging allocation
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            r = call()
            if world > 1:
                hh, mm = torch.from_numpy(r.hist).cuda(), torch.from_numpy(r.moments).cuda()
                D.allreduce_stats(hh, mm)
                torch.cuda.synchronize()
        el = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(el, op=dist.ReduceOp.MAX)
        es = 8 if args.precision == "f64" else 4
        h2d = (N_GAS * n_t + N_GAS * 17 + 4) * Me * es
        d2h = ((2 * N_GAS + 1) * n_t + 18) * Me * es + n_t * spec.bins * 8 + n_t * 32
        line["e2e"] = {"value": float(Me) * n_t * n_gpus * args.e2e_steps / float(el.item()),
                       "unit": "member-timesteps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                       "members_per_gpu": Me, "steps": args.e2e_steps, "numa_binding": numa,
                       "what": "run_ensemble(host pinned E/params -> all of C, RF, T, state + histogram back on the host), "
                               "chunked %d members, H2D/kernel/D2H overlapped on 3 streams; wall clock, max over ranks" % args.e2e_chunk}
        # secondary: the configs[3] use case proper -- host inputs in, only the ensemble statistics back
        # (histogram + moments; no trajectory leaves the GPU), same chunked pipeline
        if spec is not None:
            try:
                out2 = conc.pinned_result(N_GAS, n_t, Me, outputs=(), stats=spec, precision=args.precision, return_state=False)
                call2 = lambda: conc.run_ensemble(Eh, gph, tph, stats=spec, outputs=(), precision=args.precision,
                                                  workspace=ws, out=out2, return_state=False)
                call2()
                barrier()
                t0 = time.perf_counter()
                for _ in range(args.e2e_steps):
                    call2()
                el2 = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
                if world > 1:
                    dist.all_reduce(el2, op=dist.ReduceOp.MAX)
                line["e2e_statistics_only"] = {
                    "value": float(Me) * n_t * n_gpus * args.e2e_steps / float(el2.item()), "unit": "member-timesteps/s",
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": n_t * spec.bins * 8 + n_t * 32,
                    "what": "same host inputs, outputs=(): only the per-step T histogram and moments return to the host"}
            except Exception as exc:
                line["e2e_statistics_only"] = {"error": repr(exc)}
        ws.close()

    # ---- CPU baseline: the oracle on this box's cores, bounded sample of the same workload
    if rank == 0 and world == 1 and not args.no_cpu:   # reported at N = 1 only
        n_cpu = min(M, 262144)
        rate, thr, n_used, secs = cpu_oracle_rate(E[:, :, :n_cpu].cpu().numpy(), gp[:, :, :n_cpu].cpu().numpy(),
                                                  tp[:, :n_cpu].cpu().numpy(), args.cpu_seconds)
        line["cpu_baseline"] = {"value": rate, "unit": "member-timesteps/s", "cores": thr, "kind": "port",
                                "sample": "first %d members of rank 0's shard x %d steps, C oracle with OpenMP, %.1f s"
                                          % (n_used, n_t, secs), "host_cpus": os.cpu_count()}
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        emit(line)


if __name__ == "__main__":
    main()
