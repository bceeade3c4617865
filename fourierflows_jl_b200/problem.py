"""`Equation`, `Clock`, `Problem` -- host mirror of /root/reference/src/problem.jl."""
from __future__ import annotations

import ctypes as C
from typing import Callable, Optional

import numpy as np

from . import _lib as L
from .array import DevArray, cxtype, ffb_dtype, zeros


def make_coef(Lop, T) -> L.ffb_coef:
    """Describe `equation.L` (or an ETD coefficient) for the C ABI: scalar, dense real or dense complex."""
    c = L.ffb_coef()
    if isinstance(Lop, DevArray):
        c.ptr = Lop.ptr
        c.kind = L.FFB_COEF_COMPLEX if Lop.dtype.kind == "c" else L.FFB_COEF_REAL
        c.dtype = ffb_dtype(Lop.dtype)
        c.re = c.im = 0.0
    else:
        v = complex(Lop)
        c.ptr = None
        c.kind = L.FFB_COEF_SCALAR
        c.dtype = ffb_dtype(T)
        c.re, c.im = v.real, v.imag
    return c


class Equation:
    """`Equation(L, calcN!, grid; dims=supersize(L), T=nothing)` (src/problem.jl:11-34).

    `L`: scalar or `DevArray` (real or complex) broadcast-compatible with `sol`; `calcN(N, sol, t, clock, vars, params,
    grid)` must fully overwrite `N` (docs/src/problem.md:96-105)."""

    def __init__(self, L_, calcN: Callable, grid, dims: Optional[tuple] = None, T=None):
        self.L = L_
        self.calcN = calcN
        self.grid = grid
        if dims is None:
            if not isinstance(L_, DevArray):
                raise L.FFBError(L.FFB_EINVAL, "scalar L needs explicit dims (supersize of a number is ())")
            dims = L_.shape
        self.dims = tuple(dims)
        self.T = cxtype(grid.T) if T is None else np.dtype(T)


class Clock:
    """`Clock{T}(dt, t, step)` (src/problem.jl:43-50): dt and t are stored in the grid's float type."""

    def __init__(self, T, dt, t=0, step=0):
        self.T = np.dtype(T).type
        self.dt = self.T(dt)
        self.t = self.T(t)
        self.step = int(step)

    def __repr__(self):
        return f"Clock(dt={self.dt}, step={self.step}, t={self.t})"


class EmptyVars:
    pass


class EmptyParams:
    pass


class Problem:
    """`Problem(eqn, stepper, dt, grid, vars=EmptyVars, params=EmptyParams; stepperkwargs...)` (src/problem.jl:99-111)."""

    def __init__(self, eqn: Equation, stepper: str, dt, grid, vars=EmptyVars, params=EmptyParams, **stepperkwargs):
        from .timesteppers import TimeStepper

        dev = grid.device
        self.clock = Clock(grid.T, dt, 0, 0)
        self.timestepper = TimeStepper(stepper, eqn, dt, dev, **stepperkwargs)
        self.sol = zeros(dev, eqn.T, eqn.dims)
        self.eqn, self.grid, self.vars, self.params = eqn, grid, vars, params

    def __repr__(self):
        return f"Problem(grid on {self.grid.device}, timestepper {type(self.timestepper).__name__})"
