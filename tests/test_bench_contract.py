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
