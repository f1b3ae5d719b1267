"""bench.py's reference arm on CPU: one JSON line with the keys the driver reads (the b200 arm needs a GPU and says so)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(*args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                          timeout=600, env=e)


def test_reference_arm_line(po):
    r = run_bench("--impl", "reference", "--steps", "1", "--warmup", "0", "--ref-size", "256")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [x for x in r.stdout.splitlines() if x.startswith("{")]
    assert len(lines) == 1
    line = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
              "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in line, k
    assert line["impl"] == "reference" and line["unit"] == "GFLOP/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["vs_baseline"] is None and "workload" in line["config"]
    cb = line["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == line["value"] and cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_do_nothing(po):
    r = run_bench("--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0", "--ref-size", "256",
                  env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_b200_arm_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("GPU present")
    r = run_bench("--steps", "1", "--warmup", "3")
    assert r.returncode != 0 and "no CPU fallback" in (r.stdout + r.stderr)
