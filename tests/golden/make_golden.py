"""Generates tests/golden/golden_v1.npz: small seeded inputs and the CPU oracle's outputs for them.

The reference (FourierFlows.jl) is pure Julia and cannot be run in this image, and its test-suite holds no golden vectors
(only closed-form known-answer tests, restated in tests/test_oracle_reference_kats.py).  These fixtures therefore freeze
the ORACLE's outputs -- they guard against drift of the oracle (CPU test) and give the CUDA path a fixed target that does
not depend on the oracle code of the day (GPU test).  Regenerate with `python tests/golden/make_golden.py` from the repo
root; tests/test_golden.py compares with rtol 1e-13 (oracle) / 1e-12 per step (device, north_star)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle as fo  # noqa: E402

KAPPA = 1e-2


def gaussian(x, t=0.0, c0=0.01, sigma=0.2, kappa=KAPPA):
    return c0 * sigma / np.sqrt(sigma ** 2 + 2 * kappa * t) * np.exp(-x ** 2 / (2 * (sigma ** 2 + 2 * kappa * t)))


def build():
    g = {}
    rng = np.random.default_rng(20260117)
    # --- transforms: ragged 3-D r2c (32 x 30 x 16) and a 2-D c2c (16 x 30)
    x = np.asfortranarray(rng.standard_normal((32, 30, 16)))
    g["rfft3d_in"] = x
    g["rfft3d_out"] = fo.RfftPlan(x.shape, np.float64) * x
    z = np.asfortranarray(rng.standard_normal((16, 30)) + 1j * rng.standard_normal((16, 30)))
    g["fft2d_in"] = z
    g["fft2d_out"] = fo.FftPlan(z.shape, np.float64) * z
    # --- grid products
    grid = fo.TwoDGrid(nx=32, Lx=2 * np.pi, ny=24, Ly=3.0, aliased_fraction=1 / 3)
    g["grid2d_kr"], g["grid2d_l"], g["grid2d_Krsq"] = grid.kr.copy(), grid.l.copy(), grid.Krsq.copy()
    g["grid2d_filter"] = fo.makefilter(grid)
    fh = np.asfortranarray(np.ones((grid.nkr, grid.nl), dtype=np.complex128))
    fo.dealias(fh, grid)
    g["grid2d_dealias_mask"] = fh.real.copy()
    # --- ETD coefficients (complex128 contour mean, timesteppers.jl:689-721) for a real L spanning |dt L| over 1
    L = -np.linspace(0.0, 400.0, 33)
    for name, arr in zip(("expLdt", "exp12Ldt"), fo.getexpLs(0.01, L)):
        g["etd_" + name] = np.asarray(arr)
    for name, arr in zip(("zeta", "alpha", "beta", "gamma"), fo.getetdcoeffs(0.01, L)):
        g["etd_" + name] = np.asarray(arr)
    g["etd_L"] = L
    # --- Diffusion, all ten steppers, 100 steps from the Gaussian of the reference's tests
    for stepper in fo.STEPPERS:
        prob = fo.Diffusion.Problem(nx=128, Lx=2 * np.pi, kappa=KAPPA, dt=1e-9 / KAPPA, stepper=stepper)
        fo.Diffusion.set_c(prob, gaussian(prob.grid.x))
        fo.stepforward(prob, 100)
        g["diffusion_" + stepper] = prob.sol.copy()
    # --- 2-D vorticity (config C3 equation) 64 x 48, 5 steps
    z0 = fo.random_phase_field((64, 48), 2 * np.pi, 8.0, slope=-1, seed=77)
    g["vort_ic"] = z0
    for stepper in ("ETDRK4", "FilteredRK4", "AB3"):
        prob = fo.TwoDNavierStokes.Problem(nx=64, ny=48, nu=1e-3, dt=2e-3, stepper=stepper)
        prob.grid.rfftplan.mul(prob.sol, z0)
        fo.stepforward(prob, 5)
        g["vort_" + stepper] = prob.sol.copy()
    # --- 3-D Burgers-like (config C4 / C5 equation) 16^3, 3 steps, both precisions of the state
    for T, tag in ((np.float64, "f64"), (np.float32, "f32")):
        c0 = fo.random_phase_field((16, 16, 16), 2 * np.pi, 4.0, slope=0, seed=78, T=T)
        g["burgers_ic_" + tag] = c0
        prob = fo.Burgers3D.Problem(nx=16, kappa=1e-3, dt=1e-3, stepper="ETDRK4", T=T)
        prob.grid.rfftplan.mul(prob.sol, c0)
        fo.stepforward(prob, 3)
        g["burgers_ETDRK4_" + tag] = prob.sol.copy()
    return g


if __name__ == "__main__":
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_v1.npz")
    g = build()
    np.savez_compressed(out, **g)
    print(out, os.path.getsize(out), "bytes;", len(g), "arrays")
