"""bench.py on a small workload: one JSON line with every key the measurement contract names."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_bench_line_has_the_contract_keys():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--members-per-gpu", "40000", "--steps", "3",
                          "--warmup", "3", "--e2e-members", "8192", "--cpu-seconds", "1", "--configs-scale", "0.01"],
                         capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-3000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout[-2000:]
    d = json.loads(lines[0])
    assert d["metric"] == "ensemble member-timesteps/sec" and d["unit"] == "member-timesteps/s"
    assert d["n_gpus"] == 1 and d["steps"] == 3 and d["warmup"] == 3 and d["value"] > 0 and d["ms_per_step"] > 0
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["dtype"] == "f64" and d["data"] == "synthetic" and "workload" in d["config"] and "model" not in d["config"]
    assert d["gpu_launches"] == 4 * 3
    r = d["roofline"]
    assert r["bound"] in ("fp64", "hbm") and 0 < r["frac"] < 1.5 and r["peak"] > 0 and r["achieved"] > 0
    assert r["unit"] in ("TFLOP/s", "GB/s") and "traffic" in r and r["hbm"]["peak"] > 0 and r["fp64"]["peak"] > 20
    assert abs(r["kernel_share_of_step"] - r["kernel_ms"] / d["ms_per_step"]) < 1e-9
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert e["value"] < d["value"]            # host copies are inside the e2e timed region
    c = d["cpu_baseline"]
    assert c["kind"] == "port" and c["cores"] >= 1 and c["value"] > 0 and "sample" in c
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    # round 2: the link ceiling e2e is stated against, the measured peaks, the other BASELINE configurations
    assert e["ceiling_value"] > 0 and 0 < e["frac_of_ceiling"] < 1.5 and e["pcie_ceiling_gbs"]["d2h_alone"][0] > 1
    assert d["peaks"]["fp64_tflops"] > 20 and d["peaks"]["hbm_gbs"] > 1000
    cfg = d["configs"]
    for key in ("configs1_1e4_members", "configs2_1e6x4_scenario_shared_f64", "configs2_1e6x4_scenario_shared_f32",
                "configs4_dt0.1_T_and_statistics", "configs4_dt0.1_full_output_chunk", "newton_k3"):
        assert "error" not in cfg[key], cfg[key]
        assert cfg[key]["kernel_ms"] > 0 and cfg[key]["value"] > 0 and cfg[key]["frac"] > 0 and "kernel_variant" in cfg[key]
    assert c["textbook_scalar_loop"]["value"] > 0
