import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import fourierflows_jl_b200 as ff
import oracle as fo
from util import relerr

for nx, ny in ((64, 64), (128, 96)):
    for stepper in ("ForwardEuler", "ETDRK4"):
        nu, dt = 1e-3, 2e-3
        cp = ff.CProblem((nx, ny), 2 * np.pi, stepper=stepper, dt=dt, calcN="vorticity2d", nu=nu)
        prob = ff.TwoDNavierStokes.Problem(ff.GPU(), nx=nx, ny=ny, nu=nu, dt=dt, stepper=stepper)
        oprob = fo.TwoDNavierStokes.Problem(nx=nx, ny=ny, nu=nu, dt=dt, stepper=stepper)
        z0 = fo.random_phase_field((nx, ny), 2 * np.pi, 8.0, slope=-1, seed=1234)
        cp.set_physical(z0)
        prob.grid.rfftplan.mul(prob.sol, ff.DevArray.from_numpy(z0))
        oprob.grid.rfftplan.mul(oprob.sol, z0)
        for s in range(3):
            cp.stepforward(1); ff.stepforward(prob, 1); fo.stepforward(oprob, 1)
            a, b, c = cp.sol.to_numpy(), prob.sol.to_numpy(), oprob.sol
            d = np.abs(a - c)
            i = np.unravel_index(np.argmax(d), d.shape)
            print(f"{nx}x{ny} {stepper} step {s+1}: C-vs-oracle {relerr(a, c):.2e} API-vs-oracle {relerr(b, c):.2e} C-vs-API {relerr(a, b):.2e}; max diff at {i} val {c[i]:.3e} diff {d[i]:.2e}")

# ETD coefficients
for T, CT in ((np.float64, np.float64), (np.float32, np.float64)):
    rng = np.random.default_rng(5)
    Tf = np.dtype(T).type
    dt = Tf(0.01)
    Lr = -np.abs(rng.standard_normal(257) * 300).astype(T); Lr[0] = 0
    Lc = (Lr + 1j * rng.standard_normal(257).astype(T) * 10).astype(np.complex64 if T == np.float32 else np.complex128)
    for Lh in (Lr, Lc, 0, -2.5):
        z, a, b, g = fo.getetdcoeffs(dt, Lh)
        Ld = ff.DevArray.from_numpy(Lh) if np.ndim(Lh) else Lh
        gz, ga, gb, gg, gE, gE2 = ff.getetdcoeffs_and_expLs(dt, Ld, np.complex64 if T == np.float32 else np.complex128, 257, coef_dtype=CT)
        host = lambda v: v.to_numpy() if isinstance(v, ff.DevArray) else v
        print(np.dtype(T).name, "L kind", type(Lh).__name__, getattr(Lh, 'dtype', None), [f"{relerr(host(x), np.asarray(y)):.2e}" for x, y in ((gz, z), (ga, a), (gb, b), (gg, g))])
        if np.ndim(Lh):
            d = np.abs(host(gg) - g); i = np.argmax(d / np.abs(g)); print("   worst gamma rel", (d / np.abs(g))[i], "at L*dt =", (dt * Lh)[i], g[i], host(gg)[i])
