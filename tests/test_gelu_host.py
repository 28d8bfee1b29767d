"""Numerics of the device GELU helpers, checked on the host: csrc/common.cuh is compiled by nvcc as host code
(tests/host/gelu_host.cu) and swept against erf() in double precision. No GPU needed."""
import json
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if c and os.path.exists(c):
            return c
    return None


@pytest.mark.skipif(_nvcc() is None, reason="nvcc not available")
def test_gelu_helpers_match_erf_in_double(tmp_path):
    exe = tmp_path / "gelu_host"
    r = subprocess.run([_nvcc(), "-O2", "-std=c++17", "-o", str(exe), os.path.join(ROOT, "tests", "host", "gelu_host.cu")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = json.loads(subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout)
    assert out["n"] > 70000
    # gelu(x) = x * Phi(x): relative accuracy on both tails (bf16 output rounding is 3.9e-3)
    assert out["max_abs"] < 1.5e-6 and out["max_rel"] < 1e-5
    # d/dx gelu(x) = Phi(x) + x * phi(x)
    assert out["grad_max_abs"] < 5e-6
