"""End-to-end stepping parity: the reference's time-stepper tests (test/test_timesteppers.jl) re-run on the device
through the public API, plus step-by-step parity with the CPU oracle for the benchmark equations.

Tolerances (BASELINE.json north_star): rel-L2 <= 1e-12 per step in Float64, <= 1e-5 in Float32; linear Diffusion
<= 1e-10 after 1000 steps; the reference's own gate rtol = nsteps * 1e-12 against the analytic Gaussian."""
import numpy as np
import pytest

import oracle as fo
from util import isapprox, relerr

pytestmark = pytest.mark.gpu
KAPPA = 1e-2
DT = 1e-9 / KAPPA
rtol_timesteppers = 1e-12


def gaussian_solution(x, t, c0=0.01, sigma=0.2, kappa=1e-2):
    return c0 * sigma / np.sqrt(sigma ** 2 + 2 * kappa * t) * np.exp(-x ** 2 / (2 * (sigma ** 2 + 2 * kappa * t)))


@pytest.fixture(scope="module")
def ff():
    import fourierflows_jl_b200 as ff
    assert ff.have_device()
    return ff


@pytest.mark.parametrize("stepper", fo.STEPPERS)
@pytest.mark.parametrize("varying", [False, True], ids=["const-kappa", "array-kappa"])
def test_diffusion_1000_steps(ff, stepper, varying):
    """constantdiffusiontest_stepforward / varyingdiffusiontest_stepforward (test_timesteppers.jl:39-59) on the GPU,
    and GPU-vs-oracle <= 1e-10 after 1000 steps (config C1 of BASELINE.json uses nx=256)."""
    nsteps, nx = 1000, 128
    kappa = KAPPA * np.ones(nx) if varying else KAPPA
    prob = ff.Diffusion.Problem(ff.GPU(), nx=nx, Lx=2 * np.pi, kappa=kappa, dt=DT, stepper=stepper)
    oprob = fo.Diffusion.Problem(nx=nx, Lx=2 * np.pi, kappa=kappa, dt=DT, stepper=stepper)
    c0 = gaussian_solution(oprob.grid.x, 0)
    cf = gaussian_solution(oprob.grid.x, nsteps * oprob.clock.dt)
    ff.Diffusion.set_c(prob, c0)
    fo.Diffusion.set_c(oprob, c0)
    ff.stepforward(prob, nsteps)
    fo.stepforward(oprob, nsteps)
    ff.Diffusion.updatevars(prob)
    fo.Diffusion.updatevars(oprob)
    assert prob.clock.step == nsteps and prob.clock.t == oprob.clock.t
    c = prob.vars.c.to_numpy()
    assert isapprox(cf, c, rtol=prob.clock.step * rtol_timesteppers)       # the reference's gate
    assert relerr(prob.sol.to_numpy(), oprob.sol) <= 1e-10                  # north_star gate
    assert relerr(c, oprob.vars.c) <= 1e-10


def test_config_c1_diffusion_256_rk4(ff):
    """BASELINE.json configs[0]: Diffusion 1-D nx=256 Float64 RK4, 1000 steps"""
    prob = ff.Diffusion.Problem(ff.GPU(), nx=256, Lx=2 * np.pi, kappa=KAPPA, dt=1e-7, stepper="RK4")
    oprob = fo.Diffusion.Problem(nx=256, Lx=2 * np.pi, kappa=KAPPA, dt=1e-7, stepper="RK4")
    c0 = gaussian_solution(oprob.grid.x, 0)
    ff.Diffusion.set_c(prob, c0)
    fo.Diffusion.set_c(oprob, c0)
    ff.stepforward(prob, 1000)
    fo.stepforward(oprob, 1000)
    ff.Diffusion.updatevars(prob)
    assert relerr(prob.vars.c.to_numpy(), gaussian_solution(oprob.grid.x, 1000 * 1e-7)) <= 1e-9
    assert relerr(prob.sol.to_numpy(), oprob.sol) <= 1e-10


@pytest.mark.parametrize("stepper", [s for s in fo.STEPPERS if fo.isexplicit(s)])
def test_step_until(ff, stepper):
    """constantdiffusiontest_step_until (test_timesteppers.jl:61-70), incl. the `time_interval - clock.t` quirk (:752)"""
    t_final = 1000 * DT + 1e-6 / np.pi
    prob = ff.Diffusion.Problem(ff.GPU(), nx=128, Lx=2 * np.pi, kappa=KAPPA, dt=DT, stepper=stepper)
    oprob = fo.Diffusion.Problem(nx=128, Lx=2 * np.pi, kappa=KAPPA, dt=DT, stepper=stepper)
    c0 = gaussian_solution(oprob.grid.x, 0)
    ff.Diffusion.set_c(prob, c0)
    fo.Diffusion.set_c(oprob, c0)
    ff.step_until(prob, t_final)
    fo.step_until(oprob, t_final)
    ff.Diffusion.updatevars(prob)
    assert prob.clock.step == oprob.clock.step == 1004 and prob.clock.t == oprob.clock.t and prob.clock.dt == oprob.clock.dt
    assert isapprox(gaussian_solution(oprob.grid.x, t_final), prob.vars.c.to_numpy(), rtol=prob.clock.step * rtol_timesteppers)
    assert relerr(prob.sol.to_numpy(), oprob.sol) <= 1e-10


@pytest.mark.parametrize("stepper", ["ETDRK4", "FilteredETDRK4"])
def test_step_until_throws_for_etdrk4(ff, stepper):
    """runtests.jl:222-226"""
    prob = ff.Diffusion.Problem(ff.GPU(), nx=16, kappa=KAPPA, dt=DT, stepper=stepper)
    with pytest.raises(ff.FFBError):
        ff.step_until(prob, 1.0)
    with pytest.raises(ff.FFBError):
        ff.CProblem((16,), 2 * np.pi, stepper=stepper, dt=DT, calcN="zero", nu=KAPPA).step_until(1.0)


@pytest.mark.parametrize("stepper", fo.STEPPERS)
def test_instantiate_problem(ff, stepper):
    """test_instantiate_problem.jl:1-21"""
    prob = ff.Diffusion.Problem(ff.GPU(), nx=4, stepper=stepper)
    assert isinstance(prob, ff.Problem)
    if stepper.startswith("Filtered"):
        dummy = ff.Diffusion.Problem(ff.GPU(), nx=16, stepper=stepper)
        real = ff.Problem(dummy.eqn, stepper, 1.0, dummy.grid, dummy.vars, dummy.params, innerK=0.0, outerK=1 / 16)
        assert real.timestepper.filter.to_numpy()[2] < 1e-16
    with pytest.raises(ff.FFBError):
        ff.TimeStepper("NoSuch", prob.eqn, 0.1, ff.GPU())


def _vort_pair(ff, stepper, T, nx=128, ny=None, nu=1e-3, dt=2e-3, **kw):
    ny = nx if ny is None else ny
    prob = ff.TwoDNavierStokes.Problem(ff.GPU(), nx=nx, ny=ny, nu=nu, dt=dt, stepper=stepper, T=T, **kw)
    oprob = fo.TwoDNavierStokes.Problem(nx=nx, ny=ny, nu=nu, dt=dt, stepper=stepper, T=T)
    z0 = fo.random_phase_field((nx, ny), 2 * np.pi, 8.0, slope=-1, seed=1234, T=T)
    prob.grid.rfftplan.mul(prob.sol, ff.DevArray.from_numpy(z0))
    oprob.grid.rfftplan.mul(oprob.sol, z0)
    return prob, oprob


@pytest.mark.parametrize("stepper", fo.STEPPERS)
def test_vorticity2d_per_step_parity_f64(ff, stepper):
    """2-D vorticity (user calcN! + dealias!, config C3 shape) on 128 x 96: each step restarted from the oracle's state,
    rel-L2 <= 1e-12 per step, and <= 20e-12 after 20 free-running steps."""
    prob, oprob = _vort_pair(ff, stepper, np.float64, nx=128, ny=96)
    assert relerr(prob.sol.to_numpy(), oprob.sol) <= 1e-13
    for _ in range(5):
        ff.stepforward(prob)
        fo.stepforward(oprob)
        assert relerr(prob.sol.to_numpy(), oprob.sol) <= 1e-12
    ff.stepforward(prob, 15)
    fo.stepforward(oprob, 15)
    assert prob.clock.step == 20 and relerr(prob.sol.to_numpy(), oprob.sol) <= 20e-12


@pytest.mark.parametrize("stepper", ["ETDRK4", "FilteredRK4", "LSRK54", "AB3"])
@pytest.mark.parametrize("coef", [np.float64, np.float32], ids=["coef64", "coef32"])
def test_vorticity2d_per_step_parity_f32(ff, stepper, coef):
    kw = {"coef_dtype": coef} if "ETDRK4" in stepper else {}
    if coef == np.float32 and "ETDRK4" not in stepper:
        pytest.skip("coefficient width only applies to ETDRK4")
    prob, oprob = _vort_pair(ff, stepper, np.float32, nx=64, **kw)
    for _ in range(5):
        ff.stepforward(prob)
        fo.stepforward(oprob)
        assert relerr(prob.sol.to_numpy(), oprob.sol) <= 1e-5
    assert prob.clock.t == oprob.clock.t and prob.clock.t.dtype == np.float32


@pytest.mark.parametrize("stepper,T", [("FilteredRK4", np.float64), ("LSRK54", np.float32), ("ETDRK4", np.float32), ("ETDRK4", np.float64)])
def test_burgers3d_parity(ff, stepper, T):
    """3-D test equation of configs C4/C5 on 32 x 30 x 16 (non-power-of-two y) and 32^3"""
    for shape in ((32, 30, 16), (32, 32, 32)):
        prob = ff.Burgers3D.Problem(ff.GPU(), nx=shape[0], ny=shape[1], nz=shape[2], kappa=1e-3, dt=1e-3, stepper=stepper, T=T)
        oprob = fo.Burgers3D.Problem(nx=shape[0], ny=shape[1], nz=shape[2], kappa=1e-3, dt=1e-3, stepper=stepper, T=T)
        c0 = fo.random_phase_field(shape, 2 * np.pi, 4.0, slope=0, seed=1234, T=T)
        prob.grid.rfftplan.mul(prob.sol, ff.DevArray.from_numpy(c0))
        oprob.grid.rfftplan.mul(oprob.sol, c0)
        tol = 1e-12 if T == np.float64 else 1e-5
        for _ in range(4):
            ff.stepforward(prob)
            fo.stepforward(oprob)
            assert relerr(prob.sol.to_numpy(), oprob.sol) <= tol


@pytest.mark.parametrize("stepper", ["ETDRK4", "FilteredETDRK4", "RK4", "FilteredLSRK54", "AB3", "FilteredForwardEuler"])
def test_c_driven_problem_matches_python_driven(ff, stepper):
    """`ffb_step` (C-driven loop + built-in calcN!) must reproduce the API-driven problem and the oracle"""
    nx, nu, dt = 64, 1e-3, 2e-3
    cp = ff.CProblem((nx, nx), 2 * np.pi, stepper=stepper, dt=dt, calcN="vorticity2d", nu=nu)
    prob, oprob = _vort_pair(ff, stepper, np.float64, nx=nx, nu=nu, dt=dt)
    z0 = fo.random_phase_field((nx, nx), 2 * np.pi, 8.0, slope=-1, seed=1234)
    cp.set_physical(z0)
    assert relerr(cp.sol.to_numpy(), oprob.sol) <= 1e-13
    cp.stepforward(6)
    ff.stepforward(prob, 6)
    fo.stepforward(oprob, 6)
    t, step, _ = cp.clock
    assert step == 6 and t == float(oprob.clock.t)
    assert relerr(cp.sol.to_numpy(), oprob.sol) <= 6e-12
    assert relerr(cp.sol.to_numpy(), prob.sol.to_numpy()) <= 1e-13
    assert relerr(cp.get_physical(), oprob.grid.rfftplan.solve(oprob.sol)) <= 1e-11


def test_c_driven_diffusion_and_burgers(ff):
    nx = 128
    kap = ff.DevArray.from_numpy(KAPPA * np.ones(nx))
    cp = ff.CProblem((nx,), 2 * np.pi, stepper="RK4", dt=DT, calcN="diffusion", kappa=kap, scalar_zero_L=True, aliased_fraction=0)
    oprob = fo.Diffusion.Problem(nx=nx, kappa=KAPPA * np.ones(nx), dt=DT, stepper="RK4")
    c0 = gaussian_solution(oprob.grid.x, 0)
    cp.set_physical(c0)
    fo.Diffusion.set_c(oprob, c0)
    cp.stepforward(200)
    fo.stepforward(oprob, 200)
    assert relerr(cp.sol.to_numpy(), oprob.sol) <= 1e-10
    cp.step_until(300 * DT + 1e-8)
    fo.step_until(oprob, 300 * DT + 1e-8)
    assert cp.clock[1] == oprob.clock.step and relerr(cp.sol.to_numpy(), oprob.sol) <= 1e-10
    # Burgers 3-D, FilteredRK4 (config C4 shape) at 32^3
    cb = ff.CProblem((32, 32, 32), 2 * np.pi, stepper="FilteredRK4", dt=1e-3, calcN="burgers3d", nu=1e-3)
    ob = fo.Burgers3D.Problem(nx=32, kappa=1e-3, dt=1e-3, stepper="FilteredRK4")
    c0 = fo.random_phase_field((32, 32, 32), 2 * np.pi, 4.0, slope=0)
    cb.set_physical(c0)
    ob.grid.rfftplan.mul(ob.sol, c0)
    cb.stepforward(3)
    fo.stepforward(ob, 3)
    assert relerr(cb.sol.to_numpy(), ob.sol) <= 3e-12


def test_diagnostics_bookkeeping(ff):
    """test_diagnostics.jl:11-31 analogue: decay of the k=1 mode under RK4, sampled every 2 steps"""
    prob = ff.Diffusion.Problem(ff.GPU(), nx=6, Lx=2 * np.pi, kappa=1.0, dt=1e-3, stepper="RK4")
    ff.Diffusion.set_c(prob, np.cos(np.asarray(prob.grid.x)))
    diag = ff.Diagnostic(lambda p: ff.parsevalsum2(p.sol, p.grid), prob, freq=2, nsteps=100)
    e0 = diag[0]
    ff.stepforward(prob, diag, 100)
    assert len(diag) == 51 and diag.steps[diag.i - 1] == 100
    assert np.isclose(diag[-1], e0 * np.exp(-2.0 * float(prob.clock.t)), rtol=1e-9)


@pytest.mark.parametrize("n,stepper", [(64, "ETDRK4"), (128, "FilteredRK4")])
def test_fused_c_driven_vorticity_matches_oracle(ff, n, stepper):
    """`fused = 1`: the calcN! elementwise kernels folded into the FFT passes; same tolerance against the oracle"""
    nu, dt = 1e-3, 2e-3
    cp = ff.CProblem((n, n), 2 * np.pi, stepper=stepper, dt=dt, calcN="vorticity2d", nu=nu, fused=1)
    oprob = fo.TwoDNavierStokes.Problem(nx=n, nu=nu, dt=dt, stepper=stepper)
    z0 = fo.random_phase_field((n, n), 2 * np.pi, 8.0, slope=-1, seed=1234)
    cp.set_physical(z0)
    oprob.grid.rfftplan.mul(oprob.sol, z0)
    for s in range(5):
        cp.stepforward(1)
        fo.stepforward(oprob, 1)
        assert relerr(cp.sol.to_numpy(), oprob.sol) <= (s + 1) * 1e-12
    cb = ff.CProblem((32, 32, 32), 2 * np.pi, stepper="FilteredRK4", dt=1e-3, calcN="burgers3d", nu=1e-3, fused=1)
    ob = fo.Burgers3D.Problem(nx=32, kappa=1e-3, dt=1e-3, stepper="FilteredRK4")
    c0 = fo.random_phase_field((32, 32, 32), 2 * np.pi, 4.0, slope=0)
    cb.set_physical(c0)
    ob.grid.rfftplan.mul(ob.sol, c0)
    cb.stepforward(3)
    fo.stepforward(ob, 3)
    assert relerr(cb.sol.to_numpy(), ob.sol) <= 3e-12


@pytest.mark.parametrize("fk", [dict(innerK=0.0, outerK=0.5), dict(order=2), dict(tol=1e-8), dict(innerK=0.4, outerK=0.9, order=6, tol=1e-12)],
                         ids=["innerK0", "order", "tol", "all"])
def test_c_driven_filter_keywords_each_have_their_own_default(ff, fk):
    """`Problem(...; innerK, outerK, order, tol)` (makefilter keywords, src/domains.jl:506; test_instantiate_problem.jl:6-21): every
    keyword of ffb_problem_config falls back to the reference default on its own, and innerK = 0 is honoured"""
    n, nu, dt = 64, 1e-3, 2e-3
    cp = ff.CProblem((n, n), 2 * np.pi, stepper="FilteredRK4", dt=dt, calcN="vorticity2d", nu=nu, filter_kwargs={k: fk.get(k, d) for k, d in
                     (("order", 0.0), ("innerK", -1.0 if "innerK" not in fk else fk["innerK"]), ("outerK", 0.0), ("tol", 0.0))})
    oprob = fo.TwoDNavierStokes.Problem(nx=n, nu=nu, dt=dt, stepper="FilteredRK4", **fk)
    z0 = fo.random_phase_field((n, n), 2 * np.pi, 8.0, slope=-1, seed=1234)
    cp.set_physical(z0)
    oprob.grid.rfftplan.mul(oprob.sol, z0)
    cp.stepforward(3)
    fo.stepforward(oprob, 3)
    assert relerr(cp.sol.to_numpy(), oprob.sol) <= 3e-12
