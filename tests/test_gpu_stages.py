"""Every fused stepper stage kernel against the reference's broadcast expressions (src/timesteppers.jl) evaluated by
NumPy in the same precision, for scalar / dense-real / dense-complex coefficients and Float32 / Float64 states.
Pure elementwise work: Float64 must agree to rounding (<= 4e-16 relative L2), Float32 to <= 2e-7."""
import ctypes as C

import numpy as np
import pytest

from util import relerr

pytestmark = pytest.mark.gpu
N = 1000  # ragged on purpose (not a multiple of the block size)


@pytest.fixture(scope="module")
def ff():
    import fourierflows_jl_b200 as ff
    assert ff.have_device()
    return ff


def rnd(rng, T, n=N, cplx=True):
    a = rng.standard_normal(n)
    if cplx:
        a = a + 1j * rng.standard_normal(n)
        return a.astype(np.complex64 if T == np.float32 else np.complex128)
    return a.astype(T)


def coef_cases(rng, T, CT):
    """(host value usable in NumPy broadcasting, kind name)"""
    cC = np.complex64 if CT == np.float32 else np.complex128
    return [(CT(-0.37), "scalar"), (rng.standard_normal(N).astype(CT), "real"), ((rng.standard_normal(N) + 1j * rng.standard_normal(N)).astype(cC), "complex")]


def mk(ff, v, CT):
    from fourierflows_jl_b200.problem import make_coef
    if np.ndim(v) == 0:
        c = make_coef(complex(v), CT)
        return c, None
    d = ff.DevArray.from_numpy(v)
    return make_coef(d, CT), d


TOL = {np.float64: 4e-16, np.float32: 2e-7}
CASES = [(np.float64, np.float64), (np.float32, np.float32), (np.float32, np.float64)]


@pytest.mark.parametrize("T,CT", CASES)
@pytest.mark.parametrize("filtered", [False, True])
def test_forward_euler(ff, T, CT, filtered):
    rng = np.random.default_rng(0)
    L = ff._lib
    dt = T(0.013)
    for Lh, _ in coef_cases(rng, T, T):
        sol, Nn = rnd(rng, T), rnd(rng, T)
        filt = np.abs(rnd(rng, T, cplx=False)) if filtered else None
        ref = (filt * (sol + dt * (Nn + Lh * sol))) if filtered else (sol + dt * (Lh * sol + Nn))
        dsol, dN = ff.DevArray.from_numpy(sol), ff.DevArray.from_numpy(Nn)
        c, keep = mk(ff, Lh, T)
        df = ff.DevArray.from_numpy(filt) if filtered else None
        L.call("ffb_stage_fe", dsol.ptr, dN.ptr, C.byref(c), float(dt), df.ptr if df else None, ff.array.ffb_dtype(T), N)
        assert relerr(dsol.to_numpy(), ref.astype(sol.dtype)) <= TOL[T]


@pytest.mark.parametrize("T,CT", CASES[:2])
def test_rk4_stages(ff, T, CT):
    rng = np.random.default_rng(1)
    L = ff._lib
    dt = T(0.02)
    dty = ff.array.ffb_dtype(T)
    for Lh, _ in coef_cases(rng, T, T):
        sol = rnd(rng, T)
        Ns = [rnd(rng, T) for _ in range(4)]
        filt = np.abs(rnd(rng, T, cplx=False))
        # reference sequence (timesteppers.jl:237-261, 279) with the N_i standing in for calcN! results
        r1 = (Ns[0] + Lh * sol).astype(sol.dtype)
        s1 = (sol + (dt / 2) * r1).astype(sol.dtype)
        r2 = (Ns[1] + Lh * s1).astype(sol.dtype)
        s2 = (sol + (dt / 2) * r2).astype(sol.dtype)
        r3 = (Ns[2] + Lh * s2).astype(sol.dtype)
        s3 = (sol + dt * r3).astype(sol.dtype)
        r4 = (Ns[3] + Lh * s3).astype(sol.dtype)
        new = (sol + (dt / 6) * (r1 + 2 * r2 + 2 * r3 + r4)).astype(sol.dtype)
        newf = (new * filt).astype(sol.dtype)
        c, keep = mk(ff, Lh, T)
        dsol, dsol1 = ff.DevArray.from_numpy(sol), ff.DevArray.zeros(sol.dtype, (N,))
        dr = [ff.DevArray.from_numpy(n) for n in Ns]
        L.call("ffb_stage_rk4_substep", dsol1.ptr, dr[0].ptr, dsol.ptr, dsol.ptr, C.byref(c), float(T(dt / 2)), dty, N)
        assert relerr(dr[0].to_numpy(), r1) <= TOL[T] and relerr(dsol1.to_numpy(), s1) <= TOL[T]
        L.call("ffb_stage_rk4_substep", dsol1.ptr, dr[1].ptr, dsol1.ptr, dsol.ptr, C.byref(c), float(T(dt / 2)), dty, N)
        assert relerr(dsol1.to_numpy(), s2) <= TOL[T]
        L.call("ffb_stage_rk4_substep", dsol1.ptr, dr[2].ptr, dsol1.ptr, dsol.ptr, C.byref(c), float(dt), dty, N)
        assert relerr(dsol1.to_numpy(), s3) <= TOL[T]
        dsol_f = dsol.copy()
        L.call("ffb_stage_rk4_final", dsol.ptr, dr[0].ptr, dr[1].ptr, dr[2].ptr, dr[3].ptr, dsol1.ptr, C.byref(c), float(dt), None, 1, dty, N)
        assert relerr(dsol.to_numpy(), new) <= 2 * TOL[T] and relerr(dr[3].to_numpy(), r4) <= TOL[T]
        dr3b = ff.DevArray.from_numpy(Ns[3])
        df = ff.DevArray.from_numpy(filt)
        L.call("ffb_stage_rk4_final", dsol_f.ptr, dr[0].ptr, dr[1].ptr, dr[2].ptr, dr3b.ptr, dsol1.ptr, C.byref(c), float(dt), df.ptr, 0, dty, N)
        assert relerr(dsol_f.to_numpy(), newf) <= 2 * TOL[T]
        assert np.array_equal(dr3b.to_numpy(), Ns[3]), "store_rhs4 = 0 must leave RHS4 untouched"


@pytest.mark.parametrize("T,CT", CASES[:2])
def test_lsrk54_stage(ff, T, CT):
    rng = np.random.default_rng(2)
    L = ff._lib
    dt, A, B = T(0.02), T(-0.41789047449985195), T(0.3792103129996273)
    for Lh, _ in coef_cases(rng, T, T):
        sol, S2, rhs = rnd(rng, T), rnd(rng, T), rnd(rng, T)
        filt = np.abs(rnd(rng, T, cplx=False))
        for first, use_f in ((0, False), (1, False), (0, True)):
            r = (rhs + Lh * sol).astype(sol.dtype)
            s2 = (A * (S2 * (0 if first else 1)) + dt * r).astype(sol.dtype)
            out = (sol + B * s2).astype(sol.dtype)
            if use_f:
                out = (out * filt).astype(sol.dtype)
            c, keep = mk(ff, Lh, T)
            dsol, dS2, drhs = ff.DevArray.from_numpy(sol), ff.DevArray.from_numpy(S2), ff.DevArray.from_numpy(rhs)
            df = ff.DevArray.from_numpy(filt)
            L.call("ffb_stage_lsrk54", dsol.ptr, dS2.ptr, drhs.ptr, C.byref(c), float(A), float(B), float(dt), first, df.ptr if use_f else None,
                   ff.array.ffb_dtype(T), N)
            assert relerr(dS2.to_numpy(), s2) <= TOL[T] and relerr(dsol.to_numpy(), out) <= 2 * TOL[T]


@pytest.mark.parametrize("T,CT", CASES)
def test_etdrk4_stages(ff, T, CT):
    """timesteppers.jl:501-516,552 including Float64 coefficients on a Float32 state (getetdcoeffs quirk, :692,710-715)"""
    rng = np.random.default_rng(3)
    L = ff._lib
    dty = ff.array.ffb_dtype(T)
    cT = np.complex64 if T == np.float32 else np.complex128
    for kind in ("scalar", "real", "complex"):
        def gen():
            if kind == "scalar":
                return CT(rng.standard_normal())
            if kind == "real":
                return rng.standard_normal(N).astype(CT)
            return (rng.standard_normal(N) + 1j * rng.standard_normal(N)).astype(np.complex64 if CT == np.float32 else np.complex128)
        E, E2, ze, al, be, ga = (gen() for _ in range(6))
        sol, sol1 = rnd(rng, T), rnd(rng, T)
        N1, N2, N3, N4 = (rnd(rng, T) for _ in range(4))
        filt = np.abs(rnd(rng, T, cplx=False))
        cs = [mk(ff, v, CT) for v in (E, E2, ze, al, be, ga)]
        for c, _ in cs:
            c.dtype = ff.array.ffb_dtype(CT)
        cE, cE2, cz, ca, cb, cg = (c for c, _ in cs)
        ref12 = (E2 * sol + ze * N1).astype(cT)
        ref3 = (E2 * sol1 + ze * (2 * N3 - N1)).astype(cT)
        refu = (E * sol + al * N1 + 2 * be * (N2 + N3) + ga * N4).astype(cT)
        reff = (refu * filt).astype(cT)
        dsol, dsol1, dout = ff.DevArray.from_numpy(sol), ff.DevArray.from_numpy(sol1), ff.DevArray.zeros(cT, (N,))
        d1, d2, d3, d4 = (ff.DevArray.from_numpy(v) for v in (N1, N2, N3, N4))
        L.call("ffb_stage_etdrk4_substep12", dout.ptr, C.byref(cE2), dsol.ptr, C.byref(cz), d1.ptr, dty, N)
        assert relerr(dout.to_numpy(), ref12) <= 2 * TOL[T]
        L.call("ffb_stage_etdrk4_substep3", dout.ptr, C.byref(cE2), dsol1.ptr, C.byref(cz), d1.ptr, d3.ptr, dty, N)
        assert relerr(dout.to_numpy(), ref3) <= 2 * TOL[T]
        ds = dsol.copy()
        L.call("ffb_stage_etdrk4_update", ds.ptr, C.byref(cE), C.byref(ca), C.byref(cb), C.byref(cg), d1.ptr, d2.ptr, d3.ptr, d4.ptr, None, dty, N)
        assert relerr(ds.to_numpy(), refu) <= 2 * TOL[T]
        df = ff.DevArray.from_numpy(filt)
        ds = dsol.copy()
        L.call("ffb_stage_etdrk4_update", ds.ptr, C.byref(cE), C.byref(ca), C.byref(cb), C.byref(cg), d1.ptr, d2.ptr, d3.ptr, d4.ptr, df.ptr, dty, N)
        assert relerr(ds.to_numpy(), reff) <= 2 * TOL[T]


@pytest.mark.parametrize("T", [np.float64, np.float32])
def test_ab3_stage(ff, T):
    """timesteppers.jl:628-636: three Euler steps (clock.step < 3), Float64 constants"""
    rng = np.random.default_rng(4)
    L = ff._lib
    dt = T(0.02)
    for Lh, _ in coef_cases(rng, T, T):
        sol, rhs, m1, m2 = (rnd(rng, T) for _ in range(4))
        r = (rhs + Lh * sol).astype(sol.dtype)
        for step in (0, 2, 3, 7):
            if step < 3:
                ref = (sol + dt * r).astype(sol.dtype)
            else:
                ref = (sol + dt * (np.float64(23 / 12) * r - np.float64(16 / 12) * m1 + np.float64(5 / 12) * m2)).astype(sol.dtype)
            c, keep = mk(ff, Lh, T)
            dsol, drhs, dm1, dm2 = (ff.DevArray.from_numpy(v) for v in (sol, rhs, m1, m2))
            L.call("ffb_stage_ab3", dsol.ptr, drhs.ptr, dm1.ptr, dm2.ptr, C.byref(c), float(dt), step, None, ff.array.ffb_dtype(T), N)
            assert relerr(drhs.to_numpy(), r) <= TOL[T]
            assert relerr(dsol.to_numpy(), ref) <= 2 * TOL[T], step


@pytest.mark.parametrize("T,CT", CASES)
def test_etd_coefficients_match_oracle(ff, T, CT):
    """getexpLs / getetdcoeffs (timesteppers.jl:673-721): device kernel vs the Complex{Float64} host formula"""
    import oracle as fo
    rng = np.random.default_rng(5)
    Tf = np.dtype(T).type
    dt = Tf(0.01)
    Lr = -np.abs(rng.standard_normal(257) * 300).astype(T)
    Lr[0] = 0
    Lc = (Lr + 1j * rng.standard_normal(257).astype(T) * 10).astype(np.complex64 if T == np.float32 else np.complex128)
    for Lh in (Lr, Lc, 0, -2.5):
        z, a, b, g = fo.getetdcoeffs(dt, Lh)
        if np.ndim(Lh) == 0:  # Float64 / Int scalar L: `dt * L` is a Float64 product in Julia even for a Float32 dt
            E, E2 = fo.getexpLs(np.float64(dt), np.float64(Lh))
            z, a, b, g = fo.getetdcoeffs(np.float64(dt), np.float64(Lh))
        else:
            E, E2 = fo.getexpLs(dt, Lh)
        Ld = ff.DevArray.from_numpy(Lh) if np.ndim(Lh) else Lh
        gz, ga, gb, gg, gE, gE2 = ff.getetdcoeffs_and_expLs(dt, Ld, np.complex64 if T == np.float32 else np.complex128, 257, coef_dtype=CT)
        host = lambda v: v.to_numpy() if isinstance(v, ff.DevArray) else v
        # The reference's r = 1 contour passes close to the removable singularity when |dt*L| ~ 1, so alpha, beta and
        # gamma carry ~1e-11 relative cancellation noise on BOTH sides (zeta is well conditioned).
        tolz = 1e-14 if CT == np.float64 else 3e-7
        tolc = (1e-10 if np.iscomplexobj(Lh) else 1e-12) if CT == np.float64 else 3e-7
        assert relerr(host(gz), np.asarray(z)) <= tolz
        for got, ref in ((ga, a), (gb, b), (gg, g)):
            assert relerr(host(got), np.asarray(ref)) <= tolc
        tole = 4e-16 if (T == np.float64 or np.ndim(Lh) == 0) else 2e-7
        assert relerr(host(gE), np.asarray(E)) <= tole and relerr(host(gE2), np.asarray(E2)) <= tole
        if np.ndim(Lh) and not np.iscomplexobj(Lh):
            assert host(gz).dtype == np.dtype(CT), "real L gives real coefficients stored as coef_dtype"
