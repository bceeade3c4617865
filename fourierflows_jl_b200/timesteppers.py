"""Time steppers -- host mirror of /root/reference/src/timesteppers.jl.

The control flow (which `calcN!` is called with which arrays at which stage time) is the reference's; every group of
`@.` broadcasts between two `calcN!` calls is ONE fused device kernel behind the C ABI (seam B3).
"""
from __future__ import annotations

import ctypes as C
import math
from fractions import Fraction

import numpy as np

from . import _lib as L
from .array import DevArray, devzeros, ffb_dtype, fltype
from .domains import makefilter
from .problem import make_coef

fullyexplicitsteppers = ["ForwardEuler", "RK4", "AB3", "LSRK54", "FilteredForwardEuler", "FilteredRK4", "FilteredAB3",
                         "FilteredLSRK54"]  # src/timesteppers.jl:37-46
STEPPERS = ["ForwardEuler", "RK4", "LSRK54", "ETDRK4", "AB3", "FilteredForwardEuler", "FilteredRK4", "FilteredLSRK54",
            "FilteredETDRK4", "FilteredAB3"]


def isexplicit(stepper) -> bool:
    return str(stepper) in fullyexplicitsteppers


class AbstractTimeStepper:
    filter = None

    def _n(self, eqn):
        n = 1
        for s in eqn.dims:
            n *= s
        return n


def _fptr(ts):
    return ts.filter.ptr if ts.filter is not None else None


# ------------------------------------------------------------------ Forward Euler (src/timesteppers.jl:98-150)
class ForwardEulerTimeStepper(AbstractTimeStepper):
    def __init__(self, equation, dev):
        (self.N,) = devzeros(dev, equation.T, equation.dims, 1)

    def stepforward(self, sol, clock, equation, vars, params, grid):
        equation.calcN(self.N, sol, clock.t, clock, vars, params, grid)
        Lc = make_coef(equation.L, fltype(equation.T))
        L.call("ffb_stage_fe", sol.ptr, self.N.ptr, C.byref(Lc), float(clock.dt), _fptr(self), ffb_dtype(equation.T), self._n(equation))
        clock.t = clock.T(clock.t + clock.dt)
        clock.step += 1


class FilteredForwardEulerTimeStepper(ForwardEulerTimeStepper):
    def __init__(self, equation, dev, **filterkwargs):
        super().__init__(equation, dev)
        self.filter = makefilter(equation, **filterkwargs)


# ------------------------------------------------------------------ RK4 (:180-285)
class RK4TimeStepper(AbstractTimeStepper):
    def __init__(self, equation, dev):
        self.sol1, self.RHS1, self.RHS2, self.RHS3, self.RHS4 = devzeros(dev, equation.T, equation.dims, 5)

    def stepforward(self, sol, clock, eq, vars, params, grid):
        t, dt = clock.t, clock.dt
        T = clock.T
        n, dty = self._n(eq), ffb_dtype(eq.T)
        Lc = make_coef(eq.L, fltype(eq.T))
        half = T(dt / 2)
        # RK4substeps! (:237-258); each addlinearterm! is fused with the following substepsol!
        eq.calcN(self.RHS1, sol, t, clock, vars, params, grid)
        L.call("ffb_stage_rk4_substep", self.sol1.ptr, self.RHS1.ptr, sol.ptr, sol.ptr, C.byref(Lc), float(half), dty, n)
        eq.calcN(self.RHS2, self.sol1, T(t + half), clock, vars, params, grid)
        L.call("ffb_stage_rk4_substep", self.sol1.ptr, self.RHS2.ptr, self.sol1.ptr, sol.ptr, C.byref(Lc), float(half), dty, n)
        eq.calcN(self.RHS3, self.sol1, T(t + half), clock, vars, params, grid)
        L.call("ffb_stage_rk4_substep", self.sol1.ptr, self.RHS3.ptr, self.sol1.ptr, sol.ptr, C.byref(Lc), float(dt), dty, n)
        eq.calcN(self.RHS4, self.sol1, T(t + dt), clock, vars, params, grid)
        # addlinearterm!(RHS4) + RK4update! (:255,261) [+ `sol *= filter` :279]
        L.call("ffb_stage_rk4_final", sol.ptr, self.RHS1.ptr, self.RHS2.ptr, self.RHS3.ptr, self.RHS4.ptr, self.sol1.ptr,
               C.byref(Lc), float(dt), _fptr(self), 1, dty, n)
        clock.t = T(clock.t + clock.dt)
        clock.step += 1


class FilteredRK4TimeStepper(RK4TimeStepper):
    def __init__(self, equation, dev, **filterkwargs):
        super().__init__(equation, dev)
        self.filter = makefilter(equation, **filterkwargs)


# ------------------------------------------------------------------ LSRK54 (:318-414)
_A = [Fraction(0), Fraction(-567301805773, 1357537059087), Fraction(-2404267990393, 2016746695238),
      Fraction(-3550918686646, 2091501179385), Fraction(-1275806237668, 842570457699)]
_B = [Fraction(1432997174477, 9575080441755), Fraction(5161836677717, 13612068292357),
      Fraction(1720146321549, 2090206949498), Fraction(3134564353537, 4481467310338),
      Fraction(2277821191437, 14882151754819)]
_Cc = [Fraction(0), Fraction(1432997174477, 9575080441755), Fraction(2526269341429, 6820363962896),
       Fraction(2006345519317, 3224310063776), Fraction(2802321613138, 2924317926251)]


class LSRK54TimeStepper(AbstractTimeStepper):
    def __init__(self, equation, dev):
        self.S2, self.RHS = devzeros(dev, equation.T, equation.dims, 2)
        Tf = fltype(equation.T).type
        self.A = tuple(Tf(a.numerator / a.denominator) for a in _A)
        self.B = tuple(Tf(b.numerator / b.denominator) for b in _B)
        self.C = tuple(Tf(c.numerator / c.denominator) for c in _Cc)

    def stepforward(self, sol, clock, eq, vars, params, grid):
        t, dt, T = clock.t, clock.dt, clock.T
        n, dty = self._n(eq), ffb_dtype(eq.T)
        Lc = make_coef(eq.L, fltype(eq.T))
        for i in range(5):  # LSRK54update! (:383-395); `@. S2 = 0` is folded into the first stage
            eq.calcN(self.RHS, sol, T(t + T(self.C[i] * dt)), clock, vars, params, grid)
            filt = _fptr(self) if i == 4 else None
            L.call("ffb_stage_lsrk54", sol.ptr, self.S2.ptr, self.RHS.ptr, C.byref(Lc), float(self.A[i]), float(self.B[i]),
                   float(dt), 1 if i == 0 else 0, filt, dty, n)
        clock.t = T(clock.t + clock.dt)
        clock.step += 1


class FilteredLSRK54TimeStepper(LSRK54TimeStepper):
    def __init__(self, equation, dev, **filterkwargs):
        super().__init__(equation, dev)
        self.filter = makefilter(equation, **filterkwargs)


# ------------------------------------------------------------------ ETDRK4 (:434-558)
def getetdcoeffs_and_expLs(dt, L_, T, n, coef_dtype=np.float64):
    """`getexpLs` (:673-678) + `getetdcoeffs` (:689-721) on the device.  Returns (zeta, alpha, beta, gamma, expLdt,
    exphLdt) as scalars (scalar L) or DevArrays of `coef_dtype` reals / complex pairs."""
    Tf = fltype(T)
    Lc = make_coef(L_, Tf)
    cd = np.dtype(coef_dtype)
    if Lc.kind == L.FFB_COEF_SCALAR:
        hs = (C.c_double * 12)()
        L.call("ffb_etd_coeffs", float(dt), C.byref(Lc), ffb_dtype(Tf), ffb_dtype(cd), 1, None, None, None, None, None, None, hs)
        vals = [complex(hs[2 * i], hs[2 * i + 1]) for i in range(6)]
        if Lc.im == 0.0:
            vals = [v.real for v in vals]
        E, E2, z, a, b, g = vals
        return z, a, b, g, E, E2
    cplx = Lc.kind == L.FFB_COEF_COMPLEX
    et = (np.complex128 if cd == np.float64 else np.complex64) if cplx else cd
    arrs = [DevArray(L_.shape, et) for _ in range(6)]
    L.call("ffb_etd_coeffs", float(dt), C.byref(Lc), ffb_dtype(Tf), ffb_dtype(cd), n, *[a.ptr for a in arrs], None)
    E, E2, z, a, b, g = arrs
    return z, a, b, g, E, E2


class ETDRK4TimeStepper(AbstractTimeStepper):
    def __init__(self, equation, dt, dev, coef_dtype=np.float64):
        Tf = fltype(equation.T)
        dt = Tf.type(dt)  # ensure dt is correct type (:457)
        n = self._n(equation)
        self.zeta, self.alpha, self.beta, self.gamma, self.expLdt, self.exphLdt = getetdcoeffs_and_expLs(
            dt, equation.L, equation.T, n, coef_dtype)
        self.sol1, self.sol2, self.N1, self.N2, self.N3, self.N4 = devzeros(dev, equation.T, equation.dims, 6)
        self._cd = np.dtype(coef_dtype)

    def _coef(self, v):
        c = make_coef(v, self._cd)
        if c.kind == L.FFB_COEF_SCALAR:
            c.dtype = ffb_dtype(self._cd)
        return c

    def stepforward(self, sol, clock, eq, vars, params, grid):
        T = clock.T
        n, dty = self._n(eq), ffb_dtype(eq.T)
        cE, cE2, cz, ca, cb, cg = (self._coef(v) for v in (self.expLdt, self.exphLdt, self.zeta, self.alpha, self.beta, self.gamma))
        # ETDRK4substeps! (:518-537)
        eq.calcN(self.N1, sol, clock.t, clock, vars, params, grid)
        L.call("ffb_stage_etdrk4_substep12", self.sol1.ptr, C.byref(cE2), sol.ptr, C.byref(cz), self.N1.ptr, dty, n)
        t2 = T(clock.t + T(clock.dt / 2))
        eq.calcN(self.N2, self.sol1, t2, clock, vars, params, grid)
        L.call("ffb_stage_etdrk4_substep12", self.sol2.ptr, C.byref(cE2), sol.ptr, C.byref(cz), self.N2.ptr, dty, n)
        eq.calcN(self.N3, self.sol2, t2, clock, vars, params, grid)
        L.call("ffb_stage_etdrk4_substep3", self.sol2.ptr, C.byref(cE2), self.sol1.ptr, C.byref(cz), self.N1.ptr, self.N3.ptr, dty, n)
        t3 = T(clock.t + clock.dt)
        eq.calcN(self.N4, self.sol2, t3, clock, vars, params, grid)
        # ETDRK4update! (:502) [+ `sol *= filter` :552]
        L.call("ffb_stage_etdrk4_update", sol.ptr, C.byref(cE), C.byref(ca), C.byref(cb), C.byref(cg), self.N1.ptr, self.N2.ptr,
               self.N3.ptr, self.N4.ptr, _fptr(self), dty, n)
        clock.t = T(clock.t + clock.dt)
        clock.step += 1


class FilteredETDRK4TimeStepper(ETDRK4TimeStepper):
    def __init__(self, equation, dt, dev, coef_dtype=np.float64, **filterkwargs):
        super().__init__(equation, dt, dev, coef_dtype)
        self.filter = makefilter(equation, **filterkwargs)


# ------------------------------------------------------------------ AB3 (:565-667)
class AB3TimeStepper(AbstractTimeStepper):
    def __init__(self, equation, dev):
        self.RHS, self.RHSm1, self.RHSm2 = devzeros(dev, equation.T, equation.dims, 3)

    def stepforward(self, sol, clock, eq, vars, params, grid):
        n, dty = self._n(eq), ffb_dtype(eq.T)
        Lc = make_coef(eq.L, fltype(eq.T))
        eq.calcN(self.RHS, sol, clock.t, clock, vars, params, grid)
        # addlinearterm! + AB3update! (three Euler steps while clock.step < 3, :629) [+ filter]
        L.call("ffb_stage_ab3", sol.ptr, self.RHS.ptr, self.RHSm1.ptr, self.RHSm2.ptr, C.byref(Lc), float(clock.dt), clock.step,
               _fptr(self), dty, n)
        clock.t = clock.T(clock.t + clock.dt)
        clock.step += 1
        # `RHS_2 = RHS_1; RHS_1 = RHS` (:647-648) as a pointer rotation
        self.RHS, self.RHSm1, self.RHSm2 = self.RHSm2, self.RHS, self.RHSm1


class FilteredAB3TimeStepper(AB3TimeStepper):
    def __init__(self, equation, dev, **filterkwargs):
        super().__init__(equation, dev)
        self.filter = makefilter(equation, **filterkwargs)


_CLASSES = {c.__name__: c for c in (ForwardEulerTimeStepper, FilteredForwardEulerTimeStepper, RK4TimeStepper,
                                    FilteredRK4TimeStepper, LSRK54TimeStepper, FilteredLSRK54TimeStepper, ETDRK4TimeStepper,
                                    FilteredETDRK4TimeStepper, AB3TimeStepper, FilteredAB3TimeStepper)}


def TimeStepper(stepper, equation, dt=None, dev=None, **kw):
    """`TimeStepper(stepper, equation, dt, dev; kw...)` (:57-69): explicit steppers take (eqn, dev), ETD ones (eqn, dt, dev)."""
    name = f"{stepper}TimeStepper"
    if name not in _CLASSES:
        raise L.FFBError(L.FFB_EINVAL, f"UndefVarError: {name} not defined")
    from .array import GPU
    dev = GPU() if dev is None else dev
    if isexplicit(stepper):
        return _CLASSES[name](equation, dev, **kw)
    return _CLASSES[name](equation, dt, dev, **kw)


def stepforward(prob, *args):
    """`stepforward!(prob)`, `stepforward!(prob, nsteps)`, `stepforward!(prob, diags, nsteps)` (:6-35)."""
    from .diagnostics import increment
    if len(args) == 0:
        diags, nsteps = None, 1
    elif len(args) == 1:
        diags, nsteps = None, args[0]
    else:
        diags, nsteps = args
    for _ in range(int(nsteps)):
        prob.timestepper.stepforward(prob.sol, prob.clock, prob.eqn, prob.vars, prob.params, prob.grid)
        if diags is not None:
            increment(diags)
    return None


def step_until(prob, stop_time):
    """`step_until!(prob, stop_time)` (:734-760), bug-compatible with `t_remaining = time_interval - prob.clock.t` (:752)."""
    if isinstance(prob.timestepper, ETDRK4TimeStepper):
        raise L.FFBError(L.FFB_ESTEPPER, "step_until! requires fully explicit time stepper; does not work with ETDRK4")
    clock = prob.clock
    if not stop_time > clock.t:
        raise L.FFBError(L.FFB_EINVAL, "stop_time must be greater than prob.clock.t")
    dt = clock.dt
    time_interval = stop_time - clock.t
    nsteps = math.floor(time_interval / dt)
    stepforward(prob, nsteps)
    t_remaining = time_interval - clock.t
    clock.dt = clock.T(t_remaining)
    stepforward(prob)
    clock.dt = dt
    return None
