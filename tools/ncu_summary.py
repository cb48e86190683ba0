"""Summarise an ncu report for profiles/: key raw metrics + executed-instruction mix + top stall sites.

    ncu -i X.ncu-rep --page raw --csv > raw.csv ; ncu -i X.ncu-rep --page source --csv > src.csv
    python tools/ncu_summary.py raw.csv src.csv
"""
import collections
import csv
import re
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__average_warp_latency_per_inst_issued.ratio",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active", "launch__shared_mem_per_block_dynamic",
        "launch__grid_size", "launch__block_size", "sm__maximum_warps_per_active_cycle_pct"]


def raw(path):
    rows = list(csv.reader(open(path)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    for h, u, v in zip(hdr, units, vals):
        if h in KEYS or h.startswith("smsp__average_warps_issue_stalled"):
            if h.startswith("smsp__average_warps_issue_stalled") and float(v or 0) < 0.05:
                continue
            print(f"{h} [{u}] = {v}")


def src(path, top=25):
    rows = list(csv.reader(open(path)))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    ops, samp = collections.Counter(), collections.Counter()
    data, tot, tots = [], 0, 0
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    for k, r in enumerate(rows[2:]):
        if len(r) < len(hdr):
            continue
        s = r[ix["Source"]].strip()
        m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", s)
        op = m.group(2) if m else s
        keep2 = ("IMAD.MOV", "IMAD.WIDE", "MUFU", "BAR", "LDS", "STS", "LDG", "STG", "SYNCS", "LDC")
        op = ".".join(op.split(".")[:2]) if op.startswith(keep2) else op.split(".")[0]
        n, sm = int(r[ix["Instructions Executed"]]), int(r[ix["# Samples"]])
        ops[op] += n
        samp[op] += sm
        tot += n
        tots += sm
        data.append((sm, k, s[:64], n, {x: int(r[ix[x]]) for x in stalls}))
    print(f"\nexecuted warp-instructions: {tot}   stall samples: {tots}")
    print(f"{'opcode':14s} {'executed':>12s} {'%':>6s} {'samples%':>8s}")
    for op, n in ops.most_common(22):
        print(f"{op:14s} {n:12d} {100 * n / tot:6.2f} {100 * samp[op] / tots:8.2f}")
    print("\ntop stall sites (samples, % of all, SASS, executed, top reasons)")
    for sm, k, s, n, st in sorted(data, reverse=True)[:top]:
        t3 = sorted(st.items(), key=lambda x: -x[1])[:2]
        print(f"{sm:7d} {100 * sm / tots:5.2f}%  {s:64s} n={n:10d} {t3}")


if __name__ == "__main__":
    raw(sys.argv[1])
    if len(sys.argv) > 2:
        src(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 25)
