"""The persistent conv kernel's work decomposition (csrc/conv_work.cuh: strided / contiguous unit walks, split-K, stream-K)
enumerated on the host: exact cover of every (tile, K step), consistent piece counts, distinct workspace slots."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("nvcc") is None and not os.path.exists("/usr/local/cuda/bin/nvcc"), reason="needs nvcc (host compile only)")
def test_work_partition_exact_cover(tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    exe = str(tmp_path / "work_partition_check")
    r = subprocess.run([nvcc, "-O1", "-std=c++17", "-Wno-deprecated-gpu-targets", "-o", exe, os.path.join(ROOT, "tests", "cpu", "work_partition_check.cu")],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    r = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout[-2000:]
    assert "0 failed" in r.stdout
