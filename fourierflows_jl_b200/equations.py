"""User-level equations written against the public API, exactly as a FourierFlows.jl user would write them
(`Params`/`Vars`/`Equation(L, calcN!)`), used by the parity tests and as examples of the elementwise vocabulary.

* `TwoDNavierStokes`: the 2-D vorticity equation of the child package the reference points to (README.md:81-83;
  GeophysicalFlows `TwoDNavierStokes.calcN_advection!`, restated in SURVEY 8d C3).
* `Burgers3D`: the builder-defined 3-D test equation of SURVEY 8d C4/C5: N = -1/2 im kr rfft(irfft(sol)^2).
"""
from __future__ import annotations

import numpy as np

from . import problem as P
from .array import GPU, DevArray, cxtype, devzeros
from .domains import ThreeDGrid, TwoDGrid, dealias
from .utils import axpby, mul_real, spectral_mul


class TwoDNavierStokes:
    class Params:
        def __init__(self, nu):
            self.nu = nu

    class Vars:
        def __init__(self, grid):
            self.zeta, self.u, self.v = devzeros(grid.device, grid.T, (grid.nx, grid.ny), 3)
            self.zetah, self.uh, self.vh = devzeros(grid.device, cxtype(grid.T), (grid.nkr, grid.nl), 3)

    @staticmethod
    def calcN(N, sol, t, clock, vars, params, grid):
        spectral_mul(vars.uh, sol, grid, coef=1j, py=1, w=grid.invKrsq)     # @. vars.uh =  im * grid.l  * grid.invKrsq * sol
        spectral_mul(vars.vh, sol, grid, coef=-1j, px=1, w=grid.invKrsq)    # @. vars.vh = -im * grid.kr * grid.invKrsq * sol
        grid.rfftplan.ldiv(vars.u, vars.uh)
        grid.rfftplan.ldiv(vars.v, vars.vh)
        grid.rfftplan.ldiv(vars.zeta, sol)                                  # zetah = sol; the transform preserves its input
        mul_real(vars.u, vars.u, vars.zeta)                                 # @. uζ *= vars.ζ
        mul_real(vars.v, vars.v, vars.zeta)
        grid.rfftplan.mul(vars.uh, vars.u)
        grid.rfftplan.mul(vars.vh, vars.v)
        spectral_mul(N, vars.uh, grid, coef=-1j, px=1)                      # @. N = - im * grid.kr * vars.uh - im * grid.l * vars.vh
        spectral_mul(N, vars.vh, grid, coef=-1j, py=1, accumulate=True)
        dealias(N, grid)

    @staticmethod
    def Problem(dev=None, *, nx=256, Lx=2 * np.pi, ny=None, Ly=None, nu=0.0, dt=0.01, stepper="RK4", aliased_fraction=1 / 3,
                T=np.float64, **stepperkwargs):
        dev = GPU() if dev is None else dev
        grid = TwoDGrid(dev, nx=nx, Lx=Lx, ny=ny, Ly=Ly, aliased_fraction=aliased_fraction, T=T)
        params = TwoDNavierStokes.Params(nu)
        vars = TwoDNavierStokes.Vars(grid)
        Lop = DevArray((grid.nkr, grid.nl), grid.T)
        axpby(Lop, -nu, grid.Krsq)                                          # L = @. -ν * grid.Krsq
        eqn = P.Equation(Lop, TwoDNavierStokes.calcN, grid)
        return P.Problem(eqn, stepper, dt, grid, vars, params, **stepperkwargs)


class Burgers3D:
    class Params:
        def __init__(self, kappa):
            self.kappa = kappa

    class Vars:
        def __init__(self, grid):
            (self.c,) = devzeros(grid.device, grid.T, grid.shape, 1)
            (self.ch,) = devzeros(grid.device, cxtype(grid.T), (grid.nkr,) + tuple(grid.shape[1:]), 1)

    @staticmethod
    def calcN(N, sol, t, clock, vars, params, grid):
        grid.rfftplan.ldiv(vars.c, sol)
        mul_real(vars.c, vars.c, vars.c)                                    # @. c = c^2
        grid.rfftplan.mul(vars.ch, vars.c)
        spectral_mul(N, vars.ch, grid, coef=-0.5j, px=1, dealias=True)      # @. N = -0.5im * grid.kr * ch ; dealias!(N, grid)

    @staticmethod
    def Problem(dev=None, *, nx=64, Lx=2 * np.pi, ny=None, nz=None, kappa=1e-3, dt=1e-3, stepper="FilteredRK4",
                aliased_fraction=1 / 3, T=np.float64, **stepperkwargs):
        dev = GPU() if dev is None else dev
        grid = ThreeDGrid(dev, nx=nx, Lx=Lx, ny=ny, nz=nz, aliased_fraction=aliased_fraction, T=T)
        params = Burgers3D.Params(kappa)
        vars = Burgers3D.Vars(grid)
        Lop = DevArray((grid.nkr, grid.nl, grid.nm), grid.T)
        axpby(Lop, -kappa, grid.Krsq)
        eqn = P.Equation(Lop, Burgers3D.calcN, grid)
        return P.Problem(eqn, stepper, dt, grid, vars, params, **stepperkwargs)
