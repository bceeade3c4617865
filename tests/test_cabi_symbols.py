"""CPU-side checks of the drop-in boundary: the shared library loads, exports every symbol declared in
include/fourierflows_b200.h, the ctypes table covers them all, and compute entry points fail loudly (FFB_ECUDA)
when no device is present -- there is no CPU fallback."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "fourierflows_b200.h")


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge
    ge.build()
    import fourierflows_jl_b200 as ff
    return ff._lib.load()


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ffb_[a-z0-9_]+)\s*\(", src)) - {"ffb_calcN_fn"})


def test_header_symbols_are_exported(lib):
    names = declared_functions()
    assert len(names) >= 45
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/fourierflows_b200.h but not exported"


def test_ctypes_table_covers_header(lib):
    import fourierflows_jl_b200 as ff
    missing = set(declared_functions()) - set(ff._lib.SIGNATURES) - {"ffb_last_error"}
    assert not missing, missing
    extra = set(ff._lib.SIGNATURES) - set(declared_functions())
    assert not extra, extra


def test_header_cites_reference_lines():
    src = open(HEADER).read()
    assert src.count("src/") >= 20, "every entry point cites the reference interface it replaces"


def test_header_compiles_as_c(tmp_path):
    c = tmp_path / "t.c"
    c.write_text('#include "fourierflows_b200.h"\nint main(void){ffb_desc d; d.ndim = 1; return d.ndim - 1;}\n')
    assert os.system(f"gcc -std=c99 -Wall -Werror -I{ROOT}/include -c {c} -o {tmp_path}/t.o") == 0


def test_no_cpu_fallback(lib):
    import fourierflows_jl_b200 as ff
    if ff.have_device():
        pytest.skip("a device is present")
    p = C.c_void_p()
    n = (C.c_int64 * 3)(8, 1, 1)
    rc = lib.ffb_plan_create(C.byref(p), 1, n, 1, 0, 1, 0)
    assert rc == ff._lib.FFB_ECUDA and b"CUDA" in lib.ffb_last_error()
    with pytest.raises(ff.FFBError):
        ff.OneDGrid(ff.GPU(), nx=8, Lx=1.0)
    # argument validation happens before any device work
    n5 = (C.c_int64 * 3)(5, 1, 1)
    assert lib.ffb_plan_create(C.byref(p), 1, n5, 1, 0, 1, 0) == ff._lib.FFB_EDOMAIN


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under the package may import or execute it."""
    pkg = os.path.join(ROOT, "fourierflows_jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".jl")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", text, flags=re.M), f"{f} imports the oracle"
                assert "scipy" not in text or f == "__init__.py", f"{f} must not lean on a CPU FFT"


def test_new_entry_points_validate_arguments_without_a_device(lib):
    """ffb_fft_inverse_multi / ffb_pipeline_* reject NULL handles with FFB_EINVAL before they touch the device (host-side contract,
    checked here on the CPU; the compute behaviour is covered by the -m gpu tests)."""
    import fourierflows_jl_b200 as ff
    L = ff._lib
    L.load()
    null = C.c_void_p()
    outs = (C.c_void_p * 1)(None)
    fuse = (L.ffb_fuse * 1)()
    assert lib.ffb_fft_inverse_multi(None, None, 1, outs, fuse) == L.FFB_EINVAL
    assert lib.ffb_pipeline_create(C.byref(null), None, 2) == L.FFB_EINVAL and not null.value
    t = C.c_int(-1)
    assert lib.ffb_pipeline_submit(None, None, None, 1, C.byref(t)) == L.FFB_EINVAL
    assert lib.ffb_pipeline_wait(None, 0) == L.FFB_EINVAL
    assert lib.ffb_pipeline_destroy(None) == L.FFB_OK
    # pinned host memory needs the CUDA runtime: without a device the allocation fails loudly, it does not fall back to malloc
    if not ff.have_device():
        with pytest.raises(ff.FFBError):
            ff.PinnedBuffer((16,), "float64")


def test_free_fft_functions_check_their_arguments_on_the_host():
    """`rfft / irfft / fft / ifft` (src/FourierFlows.jl:72): type and shape errors are raised before any plan is created; without a
    device the transform itself fails loudly (no NumPy fallback)."""
    import numpy as np
    import fourierflows_jl_b200 as ff

    class Shaped:   # only shape / dtype are looked at before the plan is created
        def __init__(self, shape, dtype):
            self.shape, self.dtype = shape, np.dtype(dtype)
    for fn, arg in ((ff.rfft, Shaped((8, 8), np.complex128)), (ff.fft, Shaped((8, 8), np.float64)), (ff.ifft, Shaped((8,), np.float32))):
        with pytest.raises(TypeError):
            fn(arg)
    with pytest.raises(ValueError):
        ff.irfft(Shaped((5, 8), np.complex128), 16)
    if not ff.have_device():
        with pytest.raises(ff.FFBError):
            ff.rfft(Shaped((8, 8), np.float64))
