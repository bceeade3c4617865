"""Pins the CPU oracle against the reference's own known-answer tests (no GPU needed).

Each test restates a test of /root/reference/test (file:line cited) on top of `oracle/`.
The reference holds no golden vectors; every expectation below is the closed-form
expression from the reference's test source.
"""
import math

import numpy as np
import pytest

import oracle as fo

rtol_fft = 1e-12          # test/runtests.jl:21
rtol_timesteppers = 1e-12  # test/runtests.jl:24


def isapprox(a, b, rtol=None, atol=0.0):
    """Julia `isapprox` for arrays: norm(a-b) <= max(atol, rtol*max(norm(a), norm(b)))."""
    a, b = np.asarray(a), np.asarray(b)
    if rtol is None:
        rtol = math.sqrt(np.finfo(np.result_type(a, b, np.float32)).eps) if atol == 0 else 0.0
    return np.linalg.norm((a - b).ravel()) <= max(atol, rtol * max(np.linalg.norm(a.ravel()), np.linalg.norm(b.ravel())))


# ------------------------------------------------------------------ grids: test/test_grid.jl, runtests.jl:46-139
nx, Lx, ny, Ly, nz, Lz = 6, 2 * np.pi, 8, 4 * np.pi, 10, 3.0


def grids():
    return (fo.OneDGrid(nx=nx, Lx=Lx), fo.TwoDGrid(nx=nx, Lx=Lx, ny=ny, Ly=Ly),
            fo.ThreeDGrid(nx=nx, Lx=Lx, ny=ny, Ly=Ly, nz=nz, Lz=Lz))


def test_grid_spacing_and_wavenumbers():
    g1, g2, g3 = grids()
    for g in (g1, g2, g3):
        assert g.nx == nx
        assert isapprox(np.diff(g.x), g.dx * np.ones(nx - 1))           # testdx  test_grid.jl:6-11
        assert isapprox(g.x[-1] - g.x[0], g.Lx - g.dx)                   # testx   :27
        assert isapprox(g.k.ravel()[1], 2 * np.pi / g.Lx)                # testdk  :31
        k = g.k.ravel()
        mid = nx // 2
        assert isapprox(k[1:mid], -k[mid + 1:][::-1])                    # testk   :35-48
        assert isapprox(np.concatenate([k[:g.nkr - 1], [abs(k[g.nkr - 1])]]), g.kr.ravel())  # testkr :48
    for g in (g2, g3):
        assert isapprox(np.diff(g.y), g.dy * np.ones(ny - 1))
        assert isapprox(g.l.ravel()[1], 2 * np.pi / g.Ly)
        l = g.l.ravel()
        assert isapprox(l[1:ny // 2], -l[ny // 2 + 1:][::-1])
    assert isapprox(g3.m.ravel()[1], 2 * np.pi / g3.Lz)
    assert isapprox(np.diff(g3.z), g3.dz * np.ones(nz - 1))
    # repr() pins (runtests.jl:118-122): dx, domain end points
    assert repr(float(g1.dx)) == "1.0471975511965976"
    assert repr(float(g1.x[0])) == "-3.141592653589793" and repr(float(g1.x[-1])) == "2.094395102393195"
    assert repr(float(g2.dy)) == "1.5707963267948966" and repr(float(g2.y[-1])) == "4.71238898038469"
    assert repr(float(g3.dz)) == "0.3" and repr(float(g3.z[0])) == "-1.5" and repr(float(g3.z[-1])) == "1.2"


def test_dealias_rule():
    """testdealias, test_grid.jl:80-133: everything with |k| >= round(max(kr)*2/3) is zeroed (a box)."""
    g1, g2, g3 = grids()
    fh = np.ones(g1.nkr, dtype=complex)
    assert fo.dealias(fh, g1) is None
    kmax = round(float(g1.kr.max()) * 2 / 3)
    fh[g1.kr < kmax] = 0
    assert np.abs(fh).sum() == 0

    fh = np.ones((g2.nkr, g2.nl), dtype=complex, order="F")
    fo.dealias(fh, g2)
    kmax, lmax = round(float(g2.kr.max()) * 2 / 3), round(float(np.abs(g2.l).max()) * 2 / 3)
    keep = (g2.kr < kmax) & (g2.l < lmax) & (g2.l >= -lmax)
    assert np.all(fh[keep] == 1)
    fh[keep] = 0
    assert np.abs(fh).sum() == 0

    fh = np.ones((g3.nkr, g3.nl, g3.nm), dtype=complex, order="F")
    fo.dealias(fh, g3)
    mmax = round(float(np.abs(g3.m).max()) * 2 / 3)
    kmax, lmax = round(float(g3.kr.max()) * 2 / 3), round(float(np.abs(g3.l).max()) * 2 / 3)
    keep = (g3.kr < kmax) & (g3.l < lmax) & (g3.l >= -lmax) & (g3.m < mmax) & (g3.m >= -mmax)
    fh[keep] = 0
    assert np.abs(fh).sum() == 0


def test_no_dealias_when_fraction_zero():
    """testnodealias, test_grid.jl:135-145."""
    g = fo.TwoDGrid(nx=nx, Lx=Lx, ny=ny, Ly=Ly, aliased_fraction=0)
    fh = np.ones((g.nkr, g.nl), dtype=complex, order="F")
    assert fo.dealias(fh, g) is None
    assert np.all(fh == 1)


@pytest.mark.parametrize("a", [0, 1 / 3, 1 / 2, 1 / 4])
def test_aliased_fraction(a):
    """test_aliased_fraction, test_grid.jl:203-227 on 16 x 32 x 34."""
    n1, n2, n3 = 16, 32, 34
    g1 = fo.OneDGrid(nx=n1, Lx=Lx, aliased_fraction=a)
    g2 = fo.TwoDGrid(nx=n1, Lx=Lx, ny=n2, Ly=Lx, aliased_fraction=a)
    g3 = fo.ThreeDGrid(nx=n1, Lx=Lx, ny=n2, Ly=Lx, nz=n3, Lz=Lx, aliased_fraction=a)
    lo = lambda n: math.floor((1 - a) / 2 * n) + 1
    hi = lambda n: math.ceil((1 + a) / 2 * n)
    kral = None if a == 0 else (lo(n1), n1 // 2 + 1)
    kal = None if a == 0 else (lo(n1), hi(n1))
    lal = None if a == 0 else (lo(n2), hi(n2))
    mal = None if a == 0 else (lo(n3), hi(n3))
    for g in (g1, g2, g3):
        assert g.kralias == kral and g.kalias == kal
    assert g2.lalias == lal and g3.lalias == lal and g3.malias == mal
    with pytest.raises(ValueError):
        fo.getaliasedwavenumbers(16, 9, 1.0)


def test_alias_ranges_at_bench_sizes():
    """nk=8192, a=1/3 -> 2731:5462, kralias 2731:4097; nk=6 -> 3:4."""
    assert fo.getaliasedwavenumbers(8192, 4097, 1 / 3) == ((2731, 5462), (2731, 4097))
    assert fo.getaliasedwavenumbers(6, 4, 1 / 3) == ((3, 4), (3, 4))


def test_makefilter():
    """testmakefilter, test_grid.jl:165-173: ==1 below K=0.65, <1e-12 above 0.999."""
    for g in grids():
        f = fo.makefilter(g)
        K = fo.fforacle._nondimK(g, True)
        assert np.all(f[K < 0.65] == 1)
        assert np.all(np.abs(f[K > 0.999]) <= 1e-12)
        assert f.dtype == np.float64


def test_domain_error_for_odd_sizes():
    """runtests.jl:133-138."""
    with pytest.raises(fo.DomainError):
        fo.OneDGrid(nx=5, Lx=1)
    with pytest.raises(fo.DomainError):
        fo.TwoDGrid(nx=5, Lx=1, ny=4, Ly=2)
    with pytest.raises(fo.DomainError):
        fo.TwoDGrid(nx=4, Lx=1, ny=5, Ly=2)
    for n in [(5, 4, 6), (4, 5, 6), (4, 6, 5)]:
        with pytest.raises(fo.DomainError):
            fo.ThreeDGrid(nx=n[0], Lx=1, ny=n[1], Ly=2, nz=n[2], Lz=3)


def test_typed_grids_float32():
    """testtyped*grid, test_grid.jl:147-163."""
    g = fo.ThreeDGrid(nx=nx, Lx=Lx, ny=ny, Ly=Ly, nz=nz, Lz=Lz, T=np.float32)
    for v in (g.dx, g.dy, g.dz, g.Lx, g.Ly, g.Lz, g.x[0], g.y[0], g.z[0]):
        assert np.asarray(v).dtype == np.float32
    assert g.k.dtype == np.float32 and g.Krsq.dtype == np.float32


# ------------------------------------------------------------------ FFT: test/createffttestfunctions.jl, test_fft.jl, test_ifft.jl
def test_fft_1d_cosmx():
    g = fo.OneDGrid(nx=32, Lx=2 * np.pi)
    m, phi = 5, np.pi / 3
    k0 = g.k[1]
    f1 = np.cos(m * k0 * g.x + phi)
    f1h = g.fftplan * f1.astype(complex)
    f1hr = g.rfftplan * f1
    f1hr_mul = np.zeros(g.nkr, dtype=complex)
    g.rfftplan.mul(f1hr_mul, f1)
    th = np.zeros(g.nk, dtype=complex)
    thr = np.zeros(g.nkr, dtype=complex)
    for i in range(g.nk):
        if abs(g.k[i]) == m * k0:
            th[i] = -np.exp(np.sign(g.k[i]) * 1j * phi) * g.nx / 2
    for i in range(g.nkr):
        if abs(g.k[i]) == m * k0:
            thr[i] = -np.exp(np.sign(g.kr[i]) * 1j * phi) * g.nx / 2
    assert isapprox(f1h, th, rtol=rtol_fft)
    assert isapprox(f1hr, thr, rtol=rtol_fft)
    assert isapprox(f1hr_mul, thr, rtol=rtol_fft)
    # test_ifft.jl:1-24
    assert isapprox(f1, g.fftplan.solve(f1h).real, rtol=rtol_fft)
    f1b = np.zeros(g.nx)
    g.rfftplan.ldiv(f1b, f1hr.copy())
    assert isapprox(f1, f1b, rtol=rtol_fft)


def test_fft_2d():
    g = fo.TwoDGrid(nx=32, Lx=2 * np.pi, ny=64, Ly=3 * np.pi)
    x, y = g.x.reshape(-1, 1), g.y.reshape(1, -1)
    m, n = 5, 2
    k0, l0 = g.k[1, 0], g.l[0, 1]
    f1 = np.asfortranarray(np.cos(m * k0 * x) * np.cos(n * l0 * y))
    f2 = np.asfortranarray(np.sin(m * k0 * x + n * l0 * y))
    K, Lw, Kr = np.broadcast_to(g.k, (g.nk, g.nl)), np.broadcast_to(g.l, (g.nk, g.nl)), np.broadcast_to(g.kr, (g.nkr, g.nl))
    Lr = np.broadcast_to(g.l, (g.nkr, g.nl))
    f1h_th = np.where((np.abs(K) == m * k0) & (np.abs(Lw) == n * l0), -g.nx * g.ny / 4, 0).astype(complex)
    f2h_th = -1j * (np.where((K == m * k0) & (Lw == n * l0), -g.nx * g.ny / 2, 0) + np.where((K == -m * k0) & (Lw == -n * l0), g.nx * g.ny / 2, 0))
    f1hr_th = np.where((np.abs(Kr) == m * k0) & (np.abs(Lr) == n * l0), -g.nx * g.ny / 4, 0).astype(complex)
    f2hr_th = -1j * np.where((Kr == m * k0) & (Lr == n * l0), -g.nx * g.ny / 2, 0)
    assert isapprox(g.fftplan * f1.astype(complex), f1h_th, rtol=rtol_fft)
    assert isapprox(g.fftplan * f2.astype(complex), f2h_th, rtol=rtol_fft)
    f1hr = np.zeros((g.nkr, g.nl), dtype=complex, order="F")
    f2hr = np.zeros((g.nkr, g.nl), dtype=complex, order="F")
    g.rfftplan.mul(f1hr, f1)
    g.rfftplan.mul(f2hr, f2)
    assert isapprox(f1hr, f1hr_th, rtol=rtol_fft)
    assert isapprox(f2hr, f2hr_th, rtol=rtol_fft)
    for f, fh in ((f1, f1hr), (f2, f2hr)):
        fb = np.zeros_like(f)
        g.rfftplan.ldiv(fb, fh.copy())
        assert isapprox(f, fb, rtol=rtol_fft)


def test_fft_3d_32x30x16():
    g = fo.ThreeDGrid(nx=32, Lx=2 * np.pi, ny=30, Ly=3 * np.pi, nz=16, Lz=4.0)
    x, y, z = g.x.reshape(-1, 1, 1), g.y.reshape(1, -1, 1), g.z.reshape(1, 1, -1)
    mx, my, mz = 5, 2, 3
    k0, l0, m0 = g.k[1, 0, 0], g.l[0, 1, 0], g.m[0, 0, 1]
    f1 = np.asfortranarray(np.cos(mx * k0 * x) * np.cos(my * l0 * y) * np.cos(mz * m0 * z))
    f2 = np.asfortranarray(np.sin(mx * k0 * x + my * l0 * y + mz * m0 * z))
    sh = (g.nkr, g.nl, g.nm)
    Kr, Lr, Mr = np.broadcast_to(g.kr, sh), np.broadcast_to(g.l, sh), np.broadcast_to(g.m, sh)
    N3 = g.nx * g.ny * g.nz
    f1hr_th = np.where((np.abs(Kr) == mx * k0) & (np.abs(Lr) == my * l0) & (np.abs(Mr) == mz * m0), N3 / 8, 0).astype(complex)
    f2hr_th = -1j * np.where((Kr == mx * k0) & (Lr == my * l0) & (Mr == mz * m0), N3 / 2, 0)
    f1hr = np.zeros(sh, dtype=complex, order="F")
    f2hr = np.zeros(sh, dtype=complex, order="F")
    g.rfftplan.mul(f1hr, f1)
    g.rfftplan.mul(f2hr, f2)
    # atol guards the zero entries' rounding noise exactly as norm-based isapprox does
    assert isapprox(f1hr, f1hr_th, rtol=rtol_fft)
    assert isapprox(f2hr, f2hr_th, rtol=rtol_fft)
    shc = (g.nk, g.nl, g.nm)
    K, Lw, M = np.broadcast_to(g.k, shc), np.broadcast_to(g.l, shc), np.broadcast_to(g.m, shc)
    f1h_th = np.where((np.abs(K) == mx * k0) & (np.abs(Lw) == my * l0) & (np.abs(M) == mz * m0), N3 / 8, 0).astype(complex)
    assert isapprox(g.fftplan * f1.astype(complex), f1h_th, rtol=rtol_fft)
    for f, fh in ((f1, f1hr), (f2, f2hr)):
        fb = np.zeros_like(f)
        g.rfftplan.ldiv(fb, fh.copy())
        assert isapprox(f, fb, rtol=rtol_fft)


# ------------------------------------------------------------------ time steppers: test/test_timesteppers.jl
def gaussian_solution(x, t, c0=0.01, sigma=0.2, kappa=1e-2):
    return c0 * sigma / np.sqrt(sigma ** 2 + 2 * kappa * t) * np.exp(-x ** 2 / (2 * (sigma ** 2 + 2 * kappa * t)))


KAPPA = 1e-2
DT = 1e-9 * 1 / KAPPA


@pytest.mark.parametrize("stepper", fo.STEPPERS)
@pytest.mark.parametrize("varying", [False, True])
def test_diffusion_stepforward(stepper, varying):
    """constantdiffusiontest_stepforward / varyingdiffusiontest_stepforward, test_timesteppers.jl:39-59:
    nx=128, 1000 steps, rtol = step*1e-12 against the analytic Gaussian."""
    nsteps = 1000
    kappa = KAPPA * np.ones(128) if varying else KAPPA
    prob = fo.Diffusion.Problem(nx=128, Lx=2 * np.pi, kappa=kappa, dt=DT, stepper=stepper)
    c0 = gaussian_solution(prob.grid.x, 0)
    cf = gaussian_solution(prob.grid.x, nsteps * prob.clock.dt)
    fo.Diffusion.set_c(prob, c0)
    fo.stepforward(prob, nsteps)
    fo.Diffusion.updatevars(prob)
    assert prob.clock.step == nsteps
    assert isapprox(cf, prob.vars.c, rtol=prob.clock.step * rtol_timesteppers)


@pytest.mark.parametrize("stepper", [s for s in fo.STEPPERS if fo.isexplicit(s)])
def test_diffusion_step_until(stepper):
    """constantdiffusiontest_step_until, test_timesteppers.jl:61-70."""
    t_final = 1000 * DT + 1e-6 / np.pi
    prob = fo.Diffusion.Problem(nx=128, Lx=2 * np.pi, kappa=KAPPA, dt=DT, stepper=stepper)
    c0 = gaussian_solution(prob.grid.x, 0)
    cf = gaussian_solution(prob.grid.x, t_final)
    fo.Diffusion.set_c(prob, c0)
    fo.step_until(prob, t_final)
    fo.Diffusion.updatevars(prob)
    assert prob.clock.step == 1004  # floor((1000 dt + 1e-6/pi)/dt) = 1003 full steps + 1 partial
    assert abs(prob.clock.t - t_final) < 1e-15
    assert isapprox(cf, prob.vars.c, rtol=prob.clock.step * rtol_timesteppers)


@pytest.mark.parametrize("stepper", ["ETDRK4", "FilteredETDRK4"])
def test_step_until_throws_for_etdrk4(stepper):
    """runtests.jl:222-226."""
    prob = fo.Diffusion.Problem(nx=16, kappa=KAPPA, dt=DT, stepper=stepper)
    with pytest.raises(RuntimeError):
        fo.step_until(prob, 1.0)


@pytest.mark.parametrize("stepper", fo.STEPPERS)
def test_instantiate_problem(stepper):
    """test_instantiate_problem.jl:1-21: nx=4 constructs; filter kwargs innerK=0, outerK=1/16 give filter[3] < 1e-16."""
    prob = fo.Diffusion.Problem(nx=4, stepper=stepper)
    assert isinstance(prob, fo.Problem)
    if stepper.startswith("Filtered"):
        dummy = fo.Diffusion.Problem(nx=16, stepper=stepper)
        real = fo.Problem(dummy.eqn, stepper, 1.0, dummy.grid, dummy.vars, dummy.params, innerK=0.0, outerK=1 / 16)
        assert real.timestepper.filter[2] < 1e-16


def test_etd_coefficients_limits():
    """getetdcoeffs (timesteppers.jl:689-721): L -> 0 limits zeta = dt/2, alpha = dt/6, beta = dt/6, gamma = dt/6; scalar L."""
    dt = 0.1
    z, a, b, c = fo.getetdcoeffs(dt, 0)
    assert np.ndim(z) == 0
    assert abs(z - dt / 2) < 1e-15 and abs(a - dt / 6) < 1e-15 and abs(b - dt / 6) < 1e-15 and abs(c - dt / 6) < 1e-15
    L = -np.linspace(0, 50, 7)
    z, a, b, c = fo.getetdcoeffs(dt, L)
    zz = dt * L
    with np.errstate(all="ignore"):
        zeta_exact = np.where(zz == 0, dt / 2, dt * (np.exp(zz / 2) - 1) / zz)
    assert z.dtype == np.float64 and np.allclose(z, zeta_exact, rtol=1e-13)
    zc = fo.getetdcoeffs(dt, L.astype(complex))[0]
    assert zc.dtype == np.complex128


def test_diagnostic_decay_rk4():
    """test_diagnostics.jl:11-31 analogue: nx=6, kappa=1, 100 RK4 steps; the k=1 coefficient decays like exp(-kappa k^2 t)."""
    prob = fo.Diffusion.Problem(nx=6, Lx=2 * np.pi, kappa=1.0, dt=1e-3, stepper="RK4")
    c0 = np.cos(prob.grid.x)
    fo.Diffusion.set_c(prob, c0)
    a0 = prob.sol[1]
    fo.stepforward(prob, 100)
    assert np.isclose(prob.sol[1], a0 * np.exp(-1.0 * prob.clock.t), rtol=1e-10)


def test_jacobian_kats():
    """test/runtests.jl:243-280 + test/test_utils.jl:99-103: J(a, a) = 0, J(sin1, sin2) and J(exp1, exp2) against the analytic
    expressions on TwoDGrid(nx=64, Lx=2pi, ny=128, Ly=3pi), atol = nx*ny*10*eps"""
    nx, ny, Lx, Ly = 64, 128, 2 * np.pi, 3 * np.pi
    g = fo.TwoDGrid(nx=nx, Lx=Lx, ny=ny, Ly=Ly)
    x, y = np.asarray(g.x).reshape(-1, 1), np.asarray(g.y).reshape(1, -1)
    k0, l0 = 2 * np.pi / Lx, 2 * np.pi / Ly
    k1, l1, k2, l2 = 2 * k0, 6 * l0, 3 * k0, -3 * l0
    s1, s2 = np.asfortranarray(np.sin(k1 * x + l1 * y)), np.asfortranarray(np.sin(k2 * x + l2 * y))
    e1, e2 = np.asfortranarray(np.exp(1j * (k1 * x + l1 * y))), np.asfortranarray(np.exp(1j * (k2 * x + l2 * y)))
    atol = nx * ny * 10 * np.finfo(np.float64).eps
    assert np.linalg.norm(fo.jacobian(s1, s1, g)) <= atol
    assert np.linalg.norm(fo.jacobian(s1, s2, g) - (k1 * l2 - k2 * l1) * np.cos(k1 * x + l1 * y) * np.cos(k2 * x + l2 * y)) <= atol
    assert np.linalg.norm(fo.jacobian(e1, e2, g) - (k2 * l1 - k1 * l2) * np.exp(1j * ((k1 + k2) * x + (l1 + l2) * y))) <= atol
