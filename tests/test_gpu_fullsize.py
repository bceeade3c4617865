"""Parity at BASELINE.json's full sizes through size-independent properties (round trip, Parseval, linearity,
analytic derivatives) plus direct oracle comparisons where the CPU oracle still finishes in seconds."""
import numpy as np
import pytest

import oracle as fo
from util import relerr

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ff():
    import fourierflows_jl_b200 as ff
    assert ff.have_device()
    return ff


def test_config_c2_derivative_round_trip_4096(ff):
    """BASELINE.json configs[1]: TwoDGrid 4096^2 Float64 rfft / (ik, il) / irfft, vs analytic derivative and the oracle"""
    n = 4096
    g = ff.TwoDGrid(ff.GPU(), nx=n, Lx=2 * np.pi)
    og = fo.TwoDGrid(nx=n, Lx=2 * np.pi, dense=False)
    x, y = og.x.reshape(-1, 1), og.y.reshape(1, -1)
    u = np.asfortranarray(np.sin(3 * x + 2 * y) + 0.5 * np.cos(7 * x - 5 * y))
    ux_exact = 3 * np.cos(3 * x + 2 * y) - 3.5 * np.sin(7 * x - 5 * y)
    uy_exact = 2 * np.cos(3 * x + 2 * y) + 2.5 * np.sin(7 * x - 5 * y)
    du = ff.DevArray.from_numpy(u)
    uh = g.rfftplan * du
    uxh = ff.DevArray((g.nkr, g.nl), np.complex128)
    uyh = ff.DevArray((g.nkr, g.nl), np.complex128)
    ff.spectral_mul(uxh, uh, g, coef=1j, px=1)
    ff.spectral_mul(uyh, uh, g, coef=1j, py=1)
    ux, uy = g.rfftplan.solve(uxh).to_numpy(), g.rfftplan.solve(uyh).to_numpy()
    assert relerr(ux, ux_exact) <= 1e-12 and relerr(uy, uy_exact) <= 1e-12
    # random-phase field against the oracle
    r = fo.random_phase_field((n, n), 2 * np.pi, 64.0, slope=1.0, seed=1234)
    rh = (g.rfftplan * ff.DevArray.from_numpy(r)).to_numpy()
    assert relerr(rh, og.rfftplan * r) <= 1e-12


@pytest.mark.parametrize("T,n", [(np.float64, 8192), (np.float32, 8192)])
def test_2d_8192_properties(ff, T, n):
    """8192^2 (config C3 grid): round trip, Parseval, linearity; Float64 <= 1e-12, Float32 <= 1e-5"""
    tol = 1e-12 if T == np.float64 else 1e-5
    rng = np.random.default_rng(11)
    a = np.asfortranarray(rng.standard_normal((n, n), dtype=T))
    plan = ff.Plan((n, n), T, ff._lib.FFB_R2C)
    g = ff.TwoDGrid(ff.GPU(), nx=n, Lx=2 * np.pi, T=T)
    da = ff.DevArray.from_numpy(a)
    ah = plan * da
    back = plan.solve(ah)
    assert relerr(back.to_numpy(), a) <= tol
    # Parseval: sum |a|^2 dx dy == parsevalsum2(ah)
    lhs = float(np.sum(a.astype(np.float64) ** 2)) * float(g.dx) * float(g.dy)
    assert abs(ff.parsevalsum2(ah, g) - lhs) <= 10 * tol * lhs
    # linearity: F(2a - 3b) = 2F(a) - 3F(b) with b a shifted copy
    b = np.asfortranarray(np.roll(a, 17, axis=1))
    db = ff.DevArray.from_numpy(b)
    bh = plan * db
    comb = ff.DevArray((n, n), T)
    ff.axpby(comb, 2.0, da, -3.0, db)
    ch = plan * comb
    lin = ff.DevArray(ah.shape, ah.dtype)
    ff.axpby(lin, 2.0, ah, -3.0, bh)
    assert relerr(ch.to_numpy(), lin.to_numpy()) <= 10 * tol


@pytest.mark.parametrize("T,n,tol", [(np.float32, 512, 1e-5), (np.float64, 256, 1e-12)])
def test_3d_properties_and_oracle(ff, T, n, tol):
    rng = np.random.default_rng(12)
    a = np.asfortranarray(rng.standard_normal((n, n, n), dtype=T))
    plan = ff.Plan((n, n, n), T, ff._lib.FFB_R2C)
    da = ff.DevArray.from_numpy(a)
    ah = plan * da
    assert relerr(plan.solve(ah).to_numpy(), a) <= tol
    if n <= 256:
        assert relerr(ah.to_numpy(), fo.RfftPlan((n, n, n), T) * a) <= tol


def test_vorticity_etdrk4_2048_vs_oracle(ff):
    """config C3 equation at 2048^2 Float64 ETDRK4, C-driven: two steps against the CPU oracle (<= 1e-12 per step)"""
    n, nu, dt = 2048, 1e-4, 1e-3
    cp = ff.CProblem((n, n), 2 * np.pi, stepper="ETDRK4", dt=dt, calcN="vorticity2d", nu=nu)
    oprob = fo.TwoDNavierStokes.Problem(nx=n, nu=nu, dt=dt, stepper="ETDRK4")
    z0 = fo.random_phase_field((n, n), 2 * np.pi, 64.0, slope=-1.0, seed=1234)
    cp.set_physical(z0)
    oprob.grid.rfftplan.mul(oprob.sol, z0)
    for s in range(2):
        cp.stepforward(1)
        fo.stepforward(oprob, 1)
        assert relerr(cp.sol.to_numpy(), oprob.sol) <= (s + 1) * 1e-12
