"""bench.py's host-side logic (no GPU): the roofline arithmetic of DESIGN.md section 3 / SURVEY 8d, the measured-peak and
ncu-traffic look-ups, the synthetic state builder, and the CPU reference arm's JSON contract on a tiny sample."""
import importlib.util
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("die_bench", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    argv, sys.argv = sys.argv, ["bench.py"]
    try:
        spec.loader.exec_module(mod)
    finally:
        sys.argv = argv
    return mod


def test_roofline_arithmetic():
    b = _bench()
    M = C = 65536
    B = 4096
    alive = 6554 * B
    ms = {"physarum_forward": 5.0, "move_claim": 2.5, "field_step": 3.6, "agent_feed": 2.6, "finalize_stats": 0.03}
    meas = dict(M=M, C=C, alive_local=alive, kernel_ms=dict(ms), ms_per_step=13.75)
    r, step_bytes = b.roofline_of(meas, B, "physarum_batched_4096x256x256")
    # three-kernel path: 72 (forward with the float32 gradient pair) + 56 + 40 B per slot, 48 B per cell (+ 24 B per alive
    # agent); the survey's model counts 96 B for the forward kernel: 240 B per cell-update with M = C
    assert step_bytes == (72 + 56 + 40) * M * B + 48 * C * B + 24 * alive
    assert r["step"]["survey_bytes"] == (96 + 56 + 40) * M * B + 48 * C * B + 24 * alive
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and r["kernel"] == "physarum_forward"
    k = r["kernels"]["physarum_forward"]
    assert abs(k["gbs"] - 72 * M * B / 5.0e-3 / 1e9) < 0.1 and abs(r["achieved"] - k["gbs"]) < 1e-9
    assert abs(r["frac"] - k["gbs"] / r["peak"]) < 1e-3
    assert abs(r["step"]["gbs"] - step_bytes / 13.75e-3 / 1e9) < 0.1
    assert r["traffic"] is None or r["traffic"] > 0
    # the cluster-fused step: one kernel, 112 B per cell-update (144 B in the survey's model)
    meas = dict(M=M, C=C, alive_local=alive, kernel_ms={"physarum_forward": 4.0, "env_step_fused": 6.0}, ms_per_step=10.0,
                grad_kind=1)
    r, step_bytes = b.roofline_of(meas, B, "physarum_batched_4096x256x256")
    assert step_bytes == (80 + 112) * M * B + 24 * alive and r["kernel"] == "env_step_fused"
    assert r["step"]["survey_bytes"] == 240 * M * B + 24 * alive


def test_peak_and_traffic_lookups():
    b = _bench()
    peak, src = b.measured_hbm_peak()
    assert peak > 1000 and isinstance(src, str)
    assert b.ncu_traffic("no such workload", "physarum_forward") is None
    t = b.ncu_traffic("physarum_single_field_4096x4096", "field_step")
    assert t is None or t > 1e8


def test_synthetic_state_builder():
    b = _bench()
    med, ag = b.build_host_state((32, 48), 3, seed=5)
    assert med.shape == (3, 3, 32, 48) and ag.shape == (3, 4, 32 * 48)
    assert (med[:, 2] == 0).all() and med[:, 1].min() >= 0 and set(np.unique(med[:, 0])) <= {0.0, 1.0}
    alive = ag[:, 2] > 0
    assert np.array_equal(alive.sum(axis=1), med[:, 0].sum(axis=(1, 2)))
    assert not np.array_equal(med[0], med[1])
    med2, ag2 = b.build_host_state((32, 48), 3, seed=5)
    assert np.array_equal(med, med2) and np.array_equal(ag, ag2)


def test_reference_arm_contract_on_a_tiny_sample():
    """`bench.py --impl reference`: one JSON line with the contract's keys (a 32x32 sample, one process)."""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "3", "--warmup", "3",
                          "--cpu-field", "32", "--cpu-procs", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "cell-updates/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["n_gpus"] == 1 and line["steps"] == 3
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] == 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["e2e"]["value"] == line["value"] and line["gpu_launches"] == 0
