"""Host-side logic of the Python mirror that needs no device: alias ranges, DomainError ordering, clock arithmetic,
Diagnostic bookkeeping, LSRK54 tableau, stepper-name resolution."""
import math

import numpy as np
import pytest

import oracle as fo


@pytest.fixture(scope="module")
def ff():
    import __graft_entry__ as ge
    ge.build()
    import fourierflows_jl_b200 as ff
    return ff


@pytest.mark.parametrize("a", [0, 1 / 3, 1 / 2, 1 / 4, 0.1, 0.9])
@pytest.mark.parametrize("n", [4, 6, 10, 16, 30, 32, 34, 128, 8192, 2048])
def test_alias_ranges_match_oracle(ff, a, n):
    assert ff.getaliasedwavenumbers(n, n // 2 + 1, a) == fo.getaliasedwavenumbers(n, n // 2 + 1, a)


def test_aliased_fraction_must_be_below_one(ff):
    with pytest.raises(ff.FFBError):
        ff.getaliasedwavenumbers(16, 9, 1.0)


def test_domain_error_raised_before_any_device_work(ff):
    for ctor, kw in ((ff.OneDGrid, dict(nx=5, Lx=1)), (ff.TwoDGrid, dict(nx=4, Lx=1, ny=5, Ly=2)),
                     (ff.ThreeDGrid, dict(nx=4, Lx=1, ny=6, Ly=2, nz=5, Lz=3))):
        with pytest.raises(ff.DomainError):
            ctor(ff.GPU(), **kw)


def test_clock_uses_grid_float_type(ff):
    c = ff.Clock(np.float32, 0.1)
    assert c.dt.dtype == np.float32 and c.t.dtype == np.float32 and c.step == 0
    c.t = c.T(c.t + c.dt)
    assert c.t == np.float32(0.1)


def test_stepper_names(ff):
    assert ff.STEPPERS == fo.STEPPERS
    for s in ff.STEPPERS:
        assert ff.isexplicit(s) == fo.isexplicit(s)


def test_lsrk54_tableau_matches_oracle(ff):
    from fourierflows_jl_b200 import timesteppers as ts
    assert ts._A == fo.LSRK54_A and ts._B == fo.LSRK54_B and ts._Cc == fo.LSRK54_C


def test_diagnostic_bookkeeping_without_device(ff):
    class FakeClock:
        t, step = 0.0, 0

    class FakeProb:
        clock = FakeClock()

    p = FakeProb()
    d = ff.Diagnostic(lambda pr: pr.clock.step * 10, p, freq=3, nsteps=10)
    assert len(d.data) == math.ceil(11 / 3) and d.i == 1 and d[0] == 0
    for s in range(1, 31):
        p.clock.step = s
        p.clock.t = 0.1 * s
        ff.increment(d)
    assert d.i == 11 and d[-1] == 300 and d.steps[10] == 30 and len(d.data) >= 11  # extended past ndata


def test_range_matches_julia_twice_precision(ff):
    from fourierflows_jl_b200.domains import _range
    z = _range(-1.5, 0.3, 10, np.float64)
    assert repr(float(z[-1])) == "1.2"
    assert np.array_equal(z, fo.fforacle._range(-1.5, 0.3, 10, np.float64))
