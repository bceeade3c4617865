"""Host-compiled check of the register butterflies (radix 2/4/8/16, both directions) against a naive DFT."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_radix_butterflies_on_host(tmp_path):
    exe = tmp_path / "test_radix"
    src = os.path.join(ROOT, "tests", "native", "test_radix_host.cu")
    subprocess.run(["nvcc", "-O1", "-std=c++17", "-Wno-deprecated-gpu-targets", "-o", str(exe), src], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "worst" in out.stdout


def test_fft_design_model():
    """the index model the kernels were derived from (tools/fft_model.py) stays self-consistent"""
    import importlib.util
    spec = importlib.util.spec_from_file_location("fft_model", os.path.join(ROOT, "tools", "fft_model.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    import numpy as np
    rng = np.random.default_rng(0)
    for N, R, rad in [(64, 16, [16, 4]), (512, 16, [16, 16, 2]), (128, 16, [16, 8])]:
        x = rng.standard_normal(N) + 1j * rng.standard_normal(N)
        assert np.allclose(m.stockham(x, rad, R, -1), np.fft.fft(x), atol=1e-10)
        for wordbytes in (8, 16):      # 8: Float32 complex; 16: Float64 complex, exchanged as whole words in one phase
            w, rd = m.bank_conflicts(N, R, rad, wordbytes, 16, 1)
            assert max(w + [rd]) == 1
    for N, rad in [(4096, [16, 16, 16]), (8192, [16, 16, 8, 4]), (2048, [16, 16, 8])]:
        w, rd = m.bank_conflicts(N, 16, rad, 16, 16, 1)
        assert max(w + [rd]) == 1
