"""CPU: run the unit-FFT data flow of csrc/fft_core.cuh lane by lane on the host
(tests/host_emul.cu) against a float64 DFT - catches index-mapping errors without a GPU."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("nvcc") is None, reason="nvcc not available")
def test_unit_fft_host_emulation(tmp_path):
    exe = tmp_path / "host_emul"
    subprocess.run(["nvcc", "-std=c++17", "-O1", "--expt-relaxed-constexpr", "-w", "-o", str(exe),
                    os.path.join(ROOT, "tests", "host_emul.cu")], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "OK" in out.stdout
