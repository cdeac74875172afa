"""CPU check of bench.py's reference arm: one JSON line on stdout with the contract's keys (no GPU, no CUDA library)."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def test_reference_arm_prints_one_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--cpu-sample", "200", "--samples", "1500", "--points", "1000"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["unit"] == "points/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["gpu_launches"] == 0 and line["dtype"] == "f64"
    for key in ("metric", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "vs_baseline", "data", "config"):
        assert key in line
    # the unmodified reference when it is importable (here: /root/reference; on the GPU box: oracle/_ref), else the port
    from oracle import ref_loader

    assert line["cpu_baseline"]["kind"] == ("reference" if ref_loader.reference_root() else "port")
    assert line["cpu_baseline"]["cores"] >= 1 and line["config"]["workload"].startswith("KernelSR predict, 2-D integrator")
    assert line["e2e"] == {"value": line["value"], "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in line["config"]


def test_non_zero_ranks_of_the_reference_arm_stay_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


import pytest


@pytest.mark.gpu
def test_gpu_arm_prints_one_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "2", "--warmup", "3", "--samples", "2000",
                          "--points", "20000", "--cpu-sample", "200", "--no-extras"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["value"] > 0 and line["n_gpus"] == 1 and line["steps"] == 2 and line["warmup"] >= 3
    assert line["gpu_launches"] == 8  # 4 kernels per predict step x 2 steps
    assert line["dtype"] == "f64" and line["data"] == "synthetic"
    r = line["roofline"]
    assert r["bound"] in ("hbm", "tensor") and r["unit"] == "TFLOP/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    e = line["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] == 20000 * 2 * 8 and e["d2h_bytes_per_step"] == 15 * 20000 * 8
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["value"] > 0
    assert line["cpu_baseline"]["port_value"] > 0 and line["e2e"]["pageable_input_value"] > 0
    assert line["fit_recursion_gbs"] > 0 and line["potrf_inv_tflops"] > 0
    assert "sm_mhz" in line["clocks"] and "reasons" in line["clocks"]
    assert line["fit_s"] > 0 and line["gram_solve_s"] > 0
