"""bench.py's driver contract, as far as it can be exercised without a GPU: the reference arm (`--impl reference`, the
oracle port of predictor.py:898-1023 on host cores) prints ONE JSON line with the keys the driver reads, only rank 0
works under torchrun, and the GPU arm refuses to run without a CUDA device instead of falling back to anything."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TINY = ["--size", "128", "--slices", "2", "--ddim-steps", "2"]


def _bench(*args, env=None):
    e = dict(os.environ)
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
        e.pop(k, None)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, env=e, timeout=600)


@pytest.fixture(scope="module")
def reference_line():
    r = _bench("--impl", "reference", "--gpus", "1", "--steps", "2", "--warmup", "1", *TINY)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines          # ONE JSON line on stdout, nothing else
    return json.loads(lines[0])


def test_reference_arm_line_has_the_contract_keys(reference_line):
    d = reference_line
    assert d["impl"] == "reference"
    assert d["metric"] == "3D flow-field predictions/sec" and d["unit"] == "predictions/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 1
    assert d["dtype"] == "f32"
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d["config"]["ddim_steps"] == 2 and d["config"]["slices"] == 2 and d["config"]["size"] == 128
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] == (os.cpu_count() or 1) and cb["value"] == d["value"] and "nothing scaled" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_times_whole_predictions(reference_line):
    """Every timed step is one full prediction (nothing extrapolated): value = 1 / mean(step seconds)."""
    d = reference_line
    assert len(d["step_seconds"]) == d["steps"]
    mean = sum(d["step_seconds"]) / len(d["step_seconds"])
    assert d["value"] == pytest.approx(1.0 / mean, rel=1e-9)
    assert d["ms_per_step"] == pytest.approx(1e3 * mean, rel=1e-9)
    assert d["wall_s"] >= sum(d["step_seconds"])


def test_reference_arm_other_ranks_exit_silently():
    r = _bench("--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0", *TINY, env={"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_both_arms_quote_the_same_config():
    """`config` is built by one function from the command line alone, so the two arms' lines carry the same dict."""
    sys.path.insert(0, ROOT)
    import bench
    argv = sys.argv
    try:
        sys.argv = ["bench.py"]
        a = bench.parse()
        sys.argv = ["bench.py", "--impl", "reference"]
        b = bench.parse()
    finally:
        sys.argv = argv
    ca, cb = bench.workload_config(a), bench.workload_config(b)
    assert ca == cb and ca["ddim_steps"] == 50 and ca["slices"] == 11 and ca["size"] == 256
    assert "configs[2]" in ca["workload"] and "1.48 GB" in ca["l2"]
    assert a.gpus == 1 and a.warmup >= 3 and a.steps >= 1          # the no-flag defaults of the contract


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the refusal on a box without a GPU")
def test_gpu_arm_refuses_to_run_without_a_device():
    r = _bench("--steps", "1", "--warmup", "1", *TINY)
    assert r.returncode != 0 and r.stdout.strip() == ""
    assert "no CPU fallback" in r.stderr


_FAKE_SMI = """#!/bin/sh
# prints one CSV sample every 50 ms like `nvidia-smi --query-gpu=... -lms`
i=0
while true; do
  i=$((i+1))
  echo "$((1500+i)), 1965, Not Active, Not Active, Not Active, Active, 900.5"
  sleep 0.05
done
"""


def _fake_smi(tmp_path, monkeypatch):
    exe = tmp_path / "nvidia-smi"
    exe.write_text(_FAKE_SMI)
    exe.chmod(0o755)
    monkeypatch.setenv("PATH", f"{tmp_path}{os.pathsep}{os.environ['PATH']}")


def test_clock_sampler_windows(tmp_path, monkeypatch):
    """`clocks` comes from the samples taken DURING the timed region; a region shorter than one sampling period falls back
    to the last warm-up samples (same workload) and says so; no nvidia-smi -> stated, not invented."""
    import time
    sys.path.insert(0, ROOT)
    import bench
    _fake_smi(tmp_path, monkeypatch)
    s = bench.ClockSampler(0)
    s.start()
    time.sleep(0.4)           # warm-up steps
    s.mark()
    n_before = s.mark_idx
    time.sleep(0.4)           # timed region
    c = s.stop()
    assert n_before >= 2 and c["window"] == "timed region" and c["samples"] >= 2
    assert c["sm_max_mhz"] == 1965.0 and c["reasons"] == ["sw_power_cap"] and c["power_w_max"] == 900.5
    assert c["sm_mhz"] > 1500 + n_before                       # only samples after the mark count
    s = bench.ClockSampler(0)
    s.start()
    time.sleep(0.4)
    s.mark()
    c = s.stop()              # an (almost) empty timed region
    if c["window"] != "timed region":                          # (a sample may still slip in between mark and stop)
        assert "warm-up" in c["window"] and 1 <= c["samples"] <= 3 and c["sm_mhz"] is not None
    monkeypatch.setenv("PATH", str(tmp_path / "nowhere"))
    s = bench.ClockSampler(0)
    s.start()
    s.mark()
    assert s.stop()["reasons"] == ["nvidia-smi unavailable"]
