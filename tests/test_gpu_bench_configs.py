"""Parity on the configurations bench.py actually times: the FUSED C-driven C3 problem at the
four-step sizes against the CPU oracle, the cuFFT cross-check (cuFFT is the reference's GPU FFT backend,
/root/reference/src/domains.jl:4-5), and the user-callback calcN! seam (`ffb_calcN_fn`, docs/src/problem.md:96-105).

Tolerances (BASELINE.json north_star): rel-L2 <= 1e-12 per step / per transform in Float64, <= 1e-5 in Float32."""
import os

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


@pytest.mark.parametrize("n,nsteps", [(4096, 2), (8192, 1)])
def test_fused_c3_at_bench_size_vs_oracle(ff, n, nsteps):
    """bench.py's N = 1 workload exactly as timed: CProblem(fused=1), 2-D vorticity ETDRK4 Float64, random-phase IC seed
    1234, at 4096^2 (the four-step threshold) for 2 steps and at the full 8192^2 for 1 step; <= 1e-12 per step."""
    nu, dt = 1e-4, 1e-3
    fo.set_fft_workers(os.cpu_count() or 1)
    cp = ff.CProblem((n, n), 2 * np.pi, stepper="ETDRK4", dt=dt, calcN="vorticity2d", nu=nu, fused=1)
    oprob = fo.TwoDNavierStokes.Problem(nx=n, nu=nu, dt=dt, stepper="ETDRK4")
    z0 = fo.random_phase_field((n, n), 2 * np.pi, 64.0, slope=-1.0, seed=1234)
    cp.set_physical(z0)
    oprob.grid.rfftplan.mul(oprob.sol, z0)
    del z0
    assert relerr(cp.sol.to_numpy(), oprob.sol) <= 1e-13
    for s in range(nsteps):
        cp.stepforward(1)
        fo.stepforward(oprob, 1)
        assert relerr(cp.sol.to_numpy(), oprob.sol) <= (s + 1) * 1e-12
    cp.close()


def _torch_rfftn(a):
    """cuFFT through torch on the column-major array `a` (x fastest): the C-ordered transpose view has x as its last axis"""
    import torch
    t = torch.from_numpy(np.ascontiguousarray(a.T)).cuda()
    out = torch.fft.rfftn(t, dim=tuple(range(t.ndim)))
    return np.asfortranarray(out.cpu().numpy().T)


def _torch_irfftn(ah, shape):
    import torch
    t = torch.from_numpy(np.ascontiguousarray(ah.T)).cuda()
    out = torch.fft.irfftn(t, s=tuple(reversed(shape)), dim=tuple(range(t.ndim)))
    return np.asfortranarray(out.cpu().numpy().T)


@pytest.mark.parametrize("T,tol", [(np.float64, 1e-12), (np.float32, 1e-5)], ids=["f64", "f32"])
@pytest.mark.parametrize("shape", [(32,), (32, 64), (32, 30, 16), (4096, 4096), (8192, 8192), (512, 512, 512)],
                         ids=lambda s: "x".join(map(str, s)))
def test_rfft_irfft_match_cufft(ff, shape, T, tol):
    """ffb_fft_forward / ffb_fft_inverse against cuFFT (torch.fft on CUDA) on the reference's KAT shapes (test_fft.jl: 32,
    32x64, 32x30x16) and on the benchmark sizes; Hermitian-consistent spectra for the inverse (cuFFT's c2r, like this
    library's, ignores the inconsistent parts -- covered separately in test_gpu_fft.py)."""
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("torch sees no CUDA device")
    rng = np.random.default_rng(77)
    a = np.asfortranarray(rng.standard_normal(shape).astype(T))
    plan = ff.Plan(shape, T, ff._lib.FFB_R2C)
    ah = plan * ff.DevArray.from_numpy(a)
    ref = _torch_rfftn(a)
    got = ah.to_numpy()
    assert got.shape == ref.shape
    assert relerr(got, ref) <= tol
    back = plan.solve(ah).to_numpy()
    refb = _torch_irfftn(ref, shape)
    assert relerr(back, refb) <= tol and relerr(back, a) <= tol
    torch.cuda.empty_cache()


def test_user_callback_calcN_matches_builtin_bitwise(ff):
    """B3 form (ii) with a user callback (`ffb_calcN_fn`; reference contract docs/src/problem.md:96-105: calcN!(N, sol, t, ...)
    fully overwrites N): array-kappa diffusion (src/diffusion.jl:135-143) written from library ops inside a ctypes callback
    must reproduce the built-in FFB_CALCN_DIFFUSION bit for bit, for every explicit stepper family."""
    L = ff._lib
    nx = 128
    g = ff.OneDGrid(ff.GPU(), nx=nx, Lx=2 * np.pi)
    x = np.asarray(fo.OneDGrid(nx=nx, Lx=2 * np.pi).x)
    kap_h = 1e-2 * (1.0 + 0.3 * np.cos(x))
    kap = ff.DevArray.from_numpy(kap_h)
    c0 = 0.01 * np.exp(-x ** 2 / (2 * 0.2 ** 2))
    for stepper in ("RK4", "FilteredLSRK54", "ETDRK4", "AB3", "ForwardEuler"):
        sh = ff.DevArray((g.nkr,), np.complex128)
        ph = ff.DevArray((nx,), np.float64)
        calls, times = [0], []

        def calcN(N, sol, t, user):
            try:
                Nd = ff.DevArray((g.nkr,), np.complex128, ptr=N)
                sd = ff.DevArray((g.nkr,), np.complex128, ptr=sol)
                ff.spectral_mul(sh, sd, g, coef=1j, px=1)          # cxh = im * kr * sol
                g.rfftplan.ldiv(ph, sh)                             # cx = irfft(cxh)
                ff.mul_real(ph, ph, kap)                            # cx *= kappa
                g.rfftplan.mul(sh, ph)                              # cxh = rfft(cx)
                ff.spectral_mul(Nd, sh, g, coef=1j, px=1)           # N = im * kr * cxh
                calls[0] += 1
                times.append(t)
                return 0
            except Exception:  # noqa: BLE001  (no exception may cross the C ABI)
                return L.FFB_EINVAL

        kw = dict(stepper=stepper, dt=1e-4, scalar_zero_L=True, aliased_fraction=0)
        cb = ff.CProblem((nx,), 2 * np.pi, calcN="callback", callback=calcN, **kw)
        bi = ff.CProblem((nx,), 2 * np.pi, calcN="diffusion", kappa=kap, **kw)
        cb.set_physical(c0)
        bi.set_physical(c0)
        cb.stepforward(7)
        bi.stepforward(7)
        per_step = {"RK4": 4, "FilteredLSRK54": 5, "ETDRK4": 4, "AB3": 1, "ForwardEuler": 1}[stepper]
        assert calls[0] == 7 * per_step
        assert times[0] == 0.0 and all(b >= a for a, b in zip(times, times[1:]))
        a, b = cb.sol.to_numpy(), bi.sol.to_numpy()
        assert np.array_equal(a, b), f"{stepper}: callback calcN! differs from the built-in one"
        # and both agree with the oracle
        op = fo.Diffusion.Problem(nx=nx, Lx=2 * np.pi, kappa=kap_h, dt=1e-4, stepper=stepper)
        fo.Diffusion.set_c(op, c0)
        fo.stepforward(op, 7)
        assert relerr(a, op.sol) <= 7e-12
        cb.close()
        bi.close()


def test_callback_error_propagates(ff):
    """a failing user calcN! aborts ffb_step with the callback's status (nothing is swallowed)"""
    L = ff._lib
    cb = ff.CProblem((64,), 2 * np.pi, calcN="callback", callback=lambda N, sol, t, user: L.FFB_EINVAL, stepper="RK4", dt=1e-3,
                     scalar_zero_L=True)
    with pytest.raises(ff.FFBError):
        cb.stepforward(1)
    cb.close()
