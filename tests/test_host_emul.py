"""Runs the warp-FFT code of csrc/fft_core.cuh on the CPU (tests/host_emul.cu)."""
import os
import shutil
import subprocess

import pytest

from conftest import ROOT


def test_warp_fft_host_emulation(tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = str(tmp_path / "host_emul")
    subprocess.run([nvcc, "-std=c++17", "-O2", "-Wno-deprecated-gpu-targets", "-o", exe,
                    os.path.join(ROOT, "tests", "host_emul.cu")], check=True, capture_output=True)
    res = subprocess.run([exe], capture_output=True, text=True)
    assert res.returncode == 0 and "HOST_EMUL OK" in res.stdout, res.stdout + res.stderr
