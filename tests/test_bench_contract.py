"""CPU checks of the bench.py contract: the reference arm prints one JSON line with the agreed keys, non-zero ranks of a
torchrun launch stay silent, and the B200 arm refuses to run (loudly) where there is no GPU instead of falling back."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None, timeout=300):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, env=e, timeout=timeout)


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_reference_arm_line():
    r = _run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-sample-rows", "20000"], env={"OPENBLAS_NUM_THREADS": "2"})
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["metric"] == "grid_point_posterior_safe_set_evals_per_sec" and j["unit"] == "evals/s"
    assert j["higher_is_better"] is True and j["dtype"] == "f64" and j["data"] == "synthetic" and j["vs_baseline"] is None
    assert j["value"] > 0 and j["ms_per_step"] > 0 and j["steps"] == 1 and j["warmup"] == 0 and j["n_gpus"] == 1
    assert "C4" in j["config"]["workload"] and j["config"]["rows"] == 50 ** 4
    assert j["cpu_baseline"]["kind"] == "port" and j["cpu_baseline"]["cores"] >= 1 and j["cpu_baseline"]["value"] == j["value"]
    assert j["e2e"] == {"value": j["value"], "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_are_silent():
    r = _run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"], env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_strong_scaling_is_the_default_and_keeps_the_named_grid():
    """BASELINE config 4 is "50^4 ... sharded 8xB200": with N ranks the FIXED grid is split, per-GPU rows shrink."""
    r = _run(["--impl", "reference", "--gpus", "8", "--steps", "1", "--warmup", "0", "--cpu-sample-rows", "5000"],
             env={"RANK": "0", "WORLD_SIZE": "8", "LOCAL_RANK": "0", "OPENBLAS_NUM_THREADS": "2"})
    assert r.returncode == 0, r.stderr[-2000:]
    j = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][0])
    assert j["scaling"] == "strong" and j["config"]["rows"] == 50 ** 4 and j["config"]["rows_per_gpu"] == 50 ** 4 // 8
    assert "50x50x50x50" in j["config"]["workload"] and j["n_gpus"] == 8


def test_reference_arm_config5_line():
    r = _run(["--impl", "reference", "--config", "C5", "--steps", "1", "--warmup", "0", "--particles", "3000"],
             env={"OPENBLAS_NUM_THREADS": "2"})
    assert r.returncode == 0, r.stderr[-2000:]
    j = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][0])
    assert j["impl"] == "reference" and "C5" in j["config"]["workload"] and j["config"]["n_train"] == 512 and j["config"]["n_gps"] == 2
    assert j["value"] > 0 and j["cpu_baseline"]["kind"] == "port" and j["e2e"]["value"] == j["value"]


def test_weak_scaling_names_the_larger_grid():
    r = _run(["--impl", "reference", "--gpus", "4", "--steps", "1", "--warmup", "0", "--cpu-sample-rows", "5000", "--scaling", "weak"],
             env={"RANK": "0", "WORLD_SIZE": "4", "LOCAL_RANK": "0", "OPENBLAS_NUM_THREADS": "2"})
    assert r.returncode == 0, r.stderr[-2000:]
    j = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][0])
    assert j["scaling"] == "weak" and j["config"]["rows"] == 4 * 50 ** 4 and j["config"]["rows_per_gpu"] == 50 ** 4
    assert "50x200x50x50" in j["config"]["workload"]


@pytest.mark.skipif(_has_gpu(), reason="checks the no-GPU failure mode")
def test_b200_arm_has_no_cpu_fallback():
    r = _run(["--steps", "1", "--warmup", "0", "--num-samples", "6"])
    assert r.returncode != 0
    assert ("NativeLibraryError" in r.stderr or "no NVIDIA driver" in r.stderr) and not any(l.startswith("{") for l in r.stdout.splitlines())
