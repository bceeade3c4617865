"""`FourierFlows.Diffusion` test-bed module -- host mirror of /root/reference/src/diffusion.jl (1-D diffusion,
constant or spatially varying diffusivity)."""
from __future__ import annotations

import numpy as np

from . import problem as P
from .array import GPU, DevArray, cxtype, devzeros, zeros
from .domains import OneDGrid
from .utils import axpby, mul_real, spectral_mul


class Params:
    """src/diffusion.jl:68-74."""

    def __init__(self, grid, kappa):
        if np.ndim(kappa) == 0 and not isinstance(kappa, DevArray):
            self.kappa = kappa
        else:
            self.kappa = kappa if isinstance(kappa, DevArray) else DevArray.from_numpy(np.asarray(kappa, dtype=grid.T))


class Vars:
    """src/diffusion.jl:99-122."""

    def __init__(self, grid):
        self.c, self.cx = devzeros(grid.device, grid.T, (grid.nx,), 2)
        self.ch, self.cxh = devzeros(grid.device, cxtype(grid.T), (grid.nkr,), 2)


def calcN_const(N, sol, t, clock, vars, params, grid):
    """`@. N = 0` (src/diffusion.jl:129-133)."""
    N.fill_zero()


def calcN_array(N, sol, t, clock, vars, params, grid):
    """src/diffusion.jl:135-143."""
    spectral_mul(vars.cxh, sol, grid, coef=1j, px=1)        # @. vars.cxh = im * grid.kr * sol
    grid.rfftplan.ldiv(vars.cx, vars.cxh)                   # ldiv!(vars.cx, grid.rfftplan, vars.cxh)
    mul_real(vars.cx, vars.cx, params.kappa)                # @. vars.cx *= params.κ
    grid.rfftplan.mul(vars.cxh, vars.cx)                    # mul!(vars.cxh, grid.rfftplan, vars.cx)
    spectral_mul(N, vars.cxh, grid, coef=1j, px=1)          # @. N = im * grid.kr * vars.cxh


def Equation(grid, params):
    """src/diffusion.jl:82-90."""
    if isinstance(params.kappa, DevArray):
        return P.Equation(0, calcN_array, grid, dims=(grid.nkr,), T=cxtype(grid.T))
    Lop = zeros(grid.device, grid.T, (grid.nkr,))
    # @. L = - params.κ * grid.kr^2  ==  (-κ) * (kr*kr): kr^2 = Krsq of the 1-D grid
    axpby(Lop, -params.kappa, grid._dense("Krsq"))
    return P.Equation(Lop, calcN_const, grid)


def Problem(dev=None, *, nx=128, Lx=2 * np.pi, kappa=0, dt=0.01, stepper="RK4", aliased_fraction=0, T=np.float64):
    """`Diffusion.Problem(dev; nx, Lx, κ, dt, stepper, aliased_fraction, T)` (src/diffusion.jl:44-59)."""
    dev = GPU() if dev is None else dev
    grid = OneDGrid(dev, nx=nx, Lx=Lx, aliased_fraction=aliased_fraction, T=T)
    params = Params(grid, kappa)
    vars = Vars(grid)
    equation = Equation(grid, params)
    return P.Problem(equation, stepper, dt, grid, vars, params)


def updatevars(prob_or_vars, grid=None, sol=None):
    """`updatevars!(vars, grid, sol)` / `updatevars!(prob)` (src/diffusion.jl:150-160)."""
    if grid is None:
        vars, grid, sol = prob_or_vars.vars, prob_or_vars.grid, prob_or_vars.sol
    else:
        vars = prob_or_vars
    vars.ch.copy_from(sol)                                   # @. vars.ch = sol
    spectral_mul(vars.cxh, sol, grid, coef=1j, px=1)         # @. vars.cxh = im * grid.kr * sol
    grid.rfftplan.ldiv(vars.c, vars.ch)                      # the transform preserves its input: no deepcopy needed
    grid.rfftplan.ldiv(vars.cx, vars.cxh)


def set_c(prob, c):
    """`set_c!(prob, c)` (src/diffusion.jl:167-176)."""
    if isinstance(c, DevArray):
        prob.vars.c.copy_from(c)
    else:
        prob.vars.c.copy_from_host(np.asarray(c, dtype=prob.grid.T))
    prob.grid.rfftplan.mul(prob.sol, prob.vars.c)
    updatevars(prob)
