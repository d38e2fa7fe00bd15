"""bench.py contract checks that need no GPU: the reference arm prints ONE JSON line with the agreed keys,
non-zero ranks of a multi-process launch stay silent, and the algorithmic byte count matches SURVEY 8(d)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True,
                          env=e, timeout=600)


def test_reference_arm_prints_one_json_line():
    r = _run(["--impl", "reference", "--gpus", "1", "--steps", "4", "--warmup", "3", "--ref-envs", "512"])
    assert r.returncode == 0, r.stderr
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "batched env-steps/sec" and d["unit"] == "env-steps/s"
    assert d["higher_is_better"] is True and d["steps"] == 4 and d["warmup"] == 3 and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_silently():
    r = _run(["--impl", "reference", "--gpus", "2", "--steps", "3", "--warmup", "3", "--ref-envs", "256"],
             env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_algorithmic_bytes_match_survey():
    sys.path.insert(0, ROOT)
    import bench
    assert bench.algorithmic_bytes_per_env_step(10, 1, 1) == 377
    assert bench.algorithmic_bytes_per_env_step(4, 1, 1) == 185
    assert bench.algorithmic_bytes_per_env_step(64, 1, 1) == 2105


def test_reference_arm_names_the_b200_arms_configuration():
    """Both arms print the same `config` (the driver's same_config check); the CPU sample is described in cpu_baseline."""
    sys.path.insert(0, ROOT)
    import bench
    r = _run(["--impl", "reference", "--gpus", "1", "--steps", "3", "--warmup", "3", "--ref-envs", "256", "--ref-seconds", "0.2"])
    d = json.loads(r.stdout.strip())
    assert d["config"] == bench.workload_config("c4", 1048576, 1, 377)
    assert d["repeats"] >= 1 and d["timed_seconds"] >= 0.2 and d["sample_envs"] == 256
    assert "256 envs" in d["cpu_baseline"]["sample"]


def test_traffic_table_lookup():
    sys.path.insert(0, ROOT)
    import bench
    assert bench.ncu_traffic_bytes("c4", 1048576) > 1e8
    assert bench.ncu_traffic_bytes("c4", 12345) is None


import pytest  # noqa: E402


@pytest.mark.gpu
def test_b200_arm_contract_on_the_gpu():
    """on a GPU the line's ms_per_step x steps is the event-timed region, gpu_launches == steps,
    the roofline is recomputable from the line, the end-to-end leg declares its copies and its PCIe ceiling, and the
    legs carry their own roofline and clocks."""
    r = _run(["--gpus", "1", "--steps", "48", "--warmup", "24", "--envs", "262144", "--no-cpu", "--legs", "c2,c4_strong",
              "--e2e-steps", "4"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["metric"] == "batched env-steps/sec" and d["n_gpus"] == 1 and d["steps"] == 48 and d["gpu_launches"] == 48
    assert d["scaling"] == "weak" and d["dtype"] == "f32" and d["vs_baseline"] is None
    assert abs(d["value"] - 262144 * 48 / (d["ms_per_step"] * 48 * 1e-3)) < 1e-6 * d["value"]
    rf = d["roofline"]
    assert rf["bound"] == "hbm" and rf["algorithmic_bytes_per_env_step"] == 377
    assert abs(rf["achieved"] - 377 * 262144 / (rf["kernel_ms"] * 1e-3) / 1e9) < 1e-6 * rf["achieved"]
    assert abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-9
    assert abs(rf["kernel_ms"] - d["ms_per_step"]) < 1e-9          # one launch per step, one rank
    assert 1e9 < d["value"] < 3e10                                  # between 6 % and 170 % of the HBM roofline
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] == 262144 * 11 * 4 and e["d2h_bytes_per_step"] == 262144 * (29 * 4 + 4 + 1)
    assert 0 < e["value"] < d["value"] and 0 < e["frac"] <= 1.2 and e["pcie_ceiling"]["value"] > 0
    assert d["clocks"]["sm_mhz"] and not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    legs = d["legs"]
    assert legs["launch_floor_us"] > 0.5
    for name in ("c2", "c2_rollout_kernel", "c4_strong"):
        leg = legs[name]
        assert "error" not in leg, leg
        assert leg["value"] > 0 and leg["ms_per_step"] > 0 and leg["roofline"]["frac"] > 0 and "sm_mhz" in leg["clocks"]
    assert legs["c2"]["roofline"]["l2_resident"] and legs["c2"]["envs_per_gpu"] == 4096
    assert legs["c4_strong"]["total_envs"] == 1048576 and legs["c4_strong"]["scaling"] == "strong"
    assert legs["c2_rollout_kernel"]["value"] > legs["c2"]["value"]       # no per-step launch latency
