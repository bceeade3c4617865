"""NumPy/SciPy restatement of the FourierFlows.jl hot path (TEST INFRASTRUCTURE -- see oracle/__init__.py).

Every function cites the reference file:line it follows (paths relative to /root/reference).
Arrays are Fortran-ordered with the *same logical shape as in Julia* (x / kx fastest), so that
`a.ravel(order="K")` is byte-identical to the dense column-major buffer the C ABI sees.

FFT backend: `scipy.fft` (pocketfft) stands in for FFTW.jl (third-party, not in the reference tree;
`Project.toml:13,24` gives only the compat range "1").  Conventions pinned by the reference's own
known-answer tests (`test/test_fft.jl`, `test/test_ifft.jl`): forward unnormalised with sign -1,
inverse scaled by 1/N, half spectrum along the first dimension.
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass, field
from fractions import Fraction
from typing import Callable, Optional

import numpy as np
import scipy.fft as sfft

__all__ = [
    "DomainError", "fftfreq", "rfftfreq", "getaliasedwavenumbers", "OneDGrid", "TwoDGrid", "ThreeDGrid",
    "dealias", "makefilter_K", "makefilter", "Equation", "Clock", "Problem", "TimeStepper", "stepforward",
    "step_until", "getetdcoeffs", "getexpLs", "STEPPERS", "isexplicit", "cxtype", "fltype",
    "Diffusion", "RfftPlan", "FftPlan", "set_fft_workers", "zeros", "LSRK54_A", "LSRK54_B", "LSRK54_C",
    "TwoDNavierStokes", "Burgers3D", "random_phase_field", "jacobianh", "jacobian",
]

_WORKERS = os.cpu_count() or 1


def set_fft_workers(n: int) -> None:
    """Analogue of `FFTW.set_num_threads(nthreads)` (src/domains.jl:85,206,347)."""
    global _WORKERS
    _WORKERS = max(1, int(n))


class DomainError(ValueError):
    """Julia `DomainError` raised for odd grid sizes (src/domains.jl:66,179,316)."""


# ----------------------------------------------------------------------------- types (src/utils.jl:18-28)
def cxtype(T):
    T = np.dtype(T)
    return T if T.kind == "c" else np.dtype(np.complex64 if T == np.float32 else np.complex128)


def fltype(T):
    T = np.dtype(T)
    return T if T.kind == "f" else np.dtype(np.float32 if T == np.complex64 else np.float64)


def zeros(T, dims):
    """`zeros(dev, T, dims)` (src/utils.jl:79)."""
    return np.zeros(dims, dtype=T, order="F")


# ----------------------------------------------------------------------------- wavenumbers
def fftfreq(n: int, fs: float) -> np.ndarray:
    """AbstractFFTs `fftfreq(n, fs)` = Frequencies(((n-1)>>1)+1, n, fs/n); element i (1-based) is
    `(i-1 - (i <= n_nonneg ? 0 : n)) * multiplier` (third-party; restated)."""
    n_nonneg = ((n - 1) >> 1) + 1
    mult = fs / n
    i = np.arange(n, dtype=np.int64)
    return (i - np.where(i < n_nonneg, 0, n)).astype(np.float64) * mult


def rfftfreq(n: int, fs: float) -> np.ndarray:
    """AbstractFFTs `rfftfreq(n, fs)` = Frequencies((n>>1)+1, (n>>1)+1, fs/n)."""
    m = (n >> 1) + 1
    return np.arange(m, dtype=np.float64) * (fs / n)


def getaliasedwavenumbers(nk: int, nkr: int, aliased_fraction: float):
    """src/domains.jl:408-421.  Returns 1-based inclusive (iL, iR) ranges like Julia's `iL:iR`,
    or (None, None) when `aliased_fraction == 0`."""
    L = (1 - aliased_fraction) / 2
    R = (1 + aliased_fraction) / 2
    iL = math.floor(L * nk) + 1
    iR = math.ceil(R * nk)
    if not aliased_fraction < 1:
        raise ValueError("`aliased_fraction` must be less than 1")
    if aliased_fraction > 0:
        return (iL, iR), (iL, nkr)
    return None, None


def _slice(r):
    """1-based inclusive Julia range -> Python slice."""
    return slice(r[0] - 1, r[1])


# ----------------------------------------------------------------------------- FFT plans
class RfftPlan:
    """`grid.rfftplan` (src/domains.jl:87,208,349): r2c over *all* dims of a real (nx[,ny[,nz]]) array.
    `mul(out, a)` = `mul!(out, plan, a)`: forward, unnormalised.  `ldiv(out, ah)` = `ldiv!(out, plan, ah)`:
    inverse scaled by 1/N."""

    def __init__(self, shape, T):
        self.shape = tuple(int(s) for s in shape)
        self.T = np.dtype(T)
        nd = len(self.shape)
        self.axes = tuple(range(nd - 1, -1, -1))  # last entry = axis 0 = the halved (real) axis

    def mul(self, out, a):
        out[...] = sfft.rfftn(a, axes=self.axes, workers=_WORKERS)
        return out

    def ldiv(self, out, ah):
        s = tuple(self.shape[ax] for ax in self.axes)
        out[...] = sfft.irfftn(ah, s=s, axes=self.axes, workers=_WORKERS)
        return out

    def __mul__(self, a):  # plan * a
        return np.asfortranarray(sfft.rfftn(a, axes=self.axes, workers=_WORKERS))

    def solve(self, ah):  # plan \ ah
        s = tuple(self.shape[ax] for ax in self.axes)
        return np.asfortranarray(sfft.irfftn(ah, s=s, axes=self.axes, workers=_WORKERS))


class FftPlan:
    """`grid.fftplan` (src/domains.jl:86,207,348): c2c over all dims."""

    def __init__(self, shape, T):
        self.shape = tuple(int(s) for s in shape)
        self.T = np.dtype(T)

    def mul(self, out, a):
        out[...] = sfft.fftn(a, workers=_WORKERS)
        return out

    def ldiv(self, out, ah):
        out[...] = sfft.ifftn(ah, workers=_WORKERS)
        return out

    def __mul__(self, a):
        return np.asfortranarray(sfft.fftn(a, workers=_WORKERS))

    def solve(self, ah):
        return np.asfortranarray(sfft.ifftn(ah, workers=_WORKERS))


# ----------------------------------------------------------------------------- grids
def _range(x0, dx, n, T):
    """`range(T(x0), step=T(dx), length=nx)` (src/domains.jl:74).  Julia builds a `StepRangeLen` on
    `TwicePrecision`, i.e. element i = T(x0) + (i-1)*T(dx) evaluated in extended precision and rounded once
    (pinned by the `repr(grid)` test, test/runtests.jl:118-122: z[end] == 1.2 for nz=10, Lz=3)."""
    T = np.dtype(T).type
    wide = np.longdouble
    return (wide(T(x0)) + np.arange(n).astype(wide) * wide(T(dx))).astype(T)


class OneDGrid:
    """src/domains.jl:61-101."""

    ndim = 1

    def __init__(self, nx, Lx, x0=None, nthreads=None, T=np.float64, aliased_fraction=1 / 3, plans=True):
        if nx % 2 != 0:
            raise DomainError("nx must be even")
        T = np.dtype(T)
        x0 = -Lx / 2 if x0 is None else x0
        dx = Lx / nx
        self.T = T
        self.nx, self.nk, self.nkr = nx, nx, nx // 2 + 1
        self.dx, self.Lx = T.type(dx), T.type(Lx)
        self.x = _range(x0, dx, nx, T)
        self.k = fftfreq(nx, 2 * np.pi / Lx * nx).astype(T)
        self.kr = rfftfreq(nx, 2 * np.pi / Lx * nx).astype(T)
        with np.errstate(divide="ignore"):
            self.invksq = (1 / (self.k * self.k)).astype(T)
            self.invkrsq = (1 / (self.kr * self.kr)).astype(T)
        self.invksq[0] = 0
        self.invkrsq[0] = 0
        if nthreads is not None:
            set_fft_workers(nthreads)
        self.fftplan = FftPlan((nx,), T) if plans else None
        self.rfftplan = RfftPlan((nx,), T) if plans else None
        self.aliased_fraction = T.type(aliased_fraction)
        self.kalias, self.kralias = getaliasedwavenumbers(self.nk, self.nkr, aliased_fraction)

    @property
    def shape(self):
        return (self.nx,)


class TwoDGrid:
    """src/domains.jl:175-223."""

    ndim = 2

    def __init__(self, nx, Lx, ny=None, Ly=None, x0=None, y0=None, nthreads=None, T=np.float64,
                 aliased_fraction=1 / 3, plans=True, dense=True):
        ny = nx if ny is None else ny
        Ly = Lx if Ly is None else Ly
        if nx % 2 != 0 or ny % 2 != 0:
            raise DomainError("nx and ny must be even")
        T = np.dtype(T)
        x0 = -Lx / 2 if x0 is None else x0
        y0 = -Ly / 2 if y0 is None else y0
        dx, dy = Lx / nx, Ly / ny
        self.T = T
        self.nx, self.ny, self.nk, self.nl, self.nkr = nx, ny, nx, ny, nx // 2 + 1
        self.dx, self.dy, self.Lx, self.Ly = T.type(dx), T.type(dy), T.type(Lx), T.type(Ly)
        self.x, self.y = _range(x0, dx, nx, T), _range(y0, dy, ny, T)
        self.k = fftfreq(nx, 2 * np.pi / Lx * nx).astype(T).reshape(nx, 1)
        self.l = fftfreq(ny, 2 * np.pi / Ly * ny).astype(T).reshape(1, ny)
        self.kr = rfftfreq(nx, 2 * np.pi / Lx * nx).astype(T).reshape(self.nkr, 1)
        if dense:
            with np.errstate(divide="ignore"):
                self.Ksq = np.asfortranarray(self.k * self.k + self.l * self.l)
                self.invKsq = np.asfortranarray((1 / self.Ksq).astype(T))
                self.invKsq[0, 0] = 0
                self.Krsq = np.asfortranarray(self.kr * self.kr + self.l * self.l)
                self.invKrsq = np.asfortranarray((1 / self.Krsq).astype(T))
                self.invKrsq[0, 0] = 0
        if nthreads is not None:
            set_fft_workers(nthreads)
        self.fftplan = FftPlan((nx, ny), T) if plans else None
        self.rfftplan = RfftPlan((nx, ny), T) if plans else None
        self.aliased_fraction = T.type(aliased_fraction)
        self.kalias, self.kralias = getaliasedwavenumbers(self.nk, self.nkr, aliased_fraction)
        self.lalias, _ = getaliasedwavenumbers(self.nl, self.nl, aliased_fraction)

    @property
    def shape(self):
        return (self.nx, self.ny)


class ThreeDGrid:
    """src/domains.jl:311-366."""

    ndim = 3

    def __init__(self, nx, Lx, ny=None, Ly=None, nz=None, Lz=None, x0=None, y0=None, z0=None, nthreads=None,
                 T=np.float64, aliased_fraction=1 / 3, plans=True, dense=True):
        ny = nx if ny is None else ny
        Ly = Lx if Ly is None else Ly
        nz = nx if nz is None else nz
        Lz = Lx if Lz is None else Lz
        if nx % 2 != 0 or ny % 2 != 0 or nz % 2 != 0:
            raise DomainError("nx, ny, and nz must be even")
        T = np.dtype(T)
        x0 = -Lx / 2 if x0 is None else x0
        y0 = -Ly / 2 if y0 is None else y0
        z0 = -Lz / 2 if z0 is None else z0
        dx, dy, dz = Lx / nx, Ly / ny, Lz / nz
        self.T = T
        self.nx, self.ny, self.nz = nx, ny, nz
        self.nk, self.nl, self.nm, self.nkr = nx, ny, nz, nx // 2 + 1
        self.dx, self.dy, self.dz = T.type(dx), T.type(dy), T.type(dz)
        self.Lx, self.Ly, self.Lz = T.type(Lx), T.type(Ly), T.type(Lz)
        self.x, self.y, self.z = _range(x0, dx, nx, T), _range(y0, dy, ny, T), _range(z0, dz, nz, T)
        self.k = fftfreq(nx, 2 * np.pi / Lx * nx).astype(T).reshape(nx, 1, 1)
        self.l = fftfreq(ny, 2 * np.pi / Ly * ny).astype(T).reshape(1, ny, 1)
        self.m = fftfreq(nz, 2 * np.pi / Lz * nz).astype(T).reshape(1, 1, nz)
        self.kr = rfftfreq(nx, 2 * np.pi / Lx * nx).astype(T).reshape(self.nkr, 1, 1)
        if dense:
            with np.errstate(divide="ignore"):
                self.Ksq = np.asfortranarray(self.k * self.k + self.l * self.l + self.m * self.m)
                self.invKsq = np.asfortranarray((1 / self.Ksq).astype(T))
                self.invKsq[0, 0, 0] = 0
                self.Krsq = np.asfortranarray(self.kr * self.kr + self.l * self.l + self.m * self.m)
                self.invKrsq = np.asfortranarray((1 / self.Krsq).astype(T))
                self.invKrsq[0, 0, 0] = 0
        if nthreads is not None:
            set_fft_workers(nthreads)
        self.fftplan = FftPlan((nx, ny, nz), T) if plans else None
        self.rfftplan = RfftPlan((nx, ny, nz), T) if plans else None
        self.aliased_fraction = T.type(aliased_fraction)
        self.kalias, self.kralias = getaliasedwavenumbers(self.nk, self.nkr, aliased_fraction)
        self.lalias, _ = getaliasedwavenumbers(self.nl, self.nl // 2 + 1, aliased_fraction)
        self.malias, _ = getaliasedwavenumbers(self.nm, self.nm // 2 + 1, aliased_fraction)

    @property
    def shape(self):
        return (self.nx, self.ny, self.nz)


# ----------------------------------------------------------------------------- dealias (src/domains.jl:428-476)
def dealias(fh: np.ndarray, grid) -> None:
    """`dealias!(fh, grid)`: zero `fh[kalias,:,...]`, `fh[:,lalias,:,...]`, `fh[:,:,malias,:]`; no-op when the grid
    was built with `aliased_fraction = 0` (src/domains.jl:434).  Returns None like the reference."""
    if grid.kalias is None:
        return None
    kalias = grid.kralias if fh.shape[0] == grid.nkr else grid.kalias  # :437,450,464
    fh[_slice(kalias), ...] = 0
    if grid.ndim >= 2:
        fh[:, _slice(grid.lalias), ...] = 0
    if grid.ndim >= 3:
        fh[:, :, _slice(grid.malias), ...] = 0
    return None


# ----------------------------------------------------------------------------- filter (src/domains.jl:506-546)
def makefilter_K(K: np.ndarray, order=4, innerK=2 / 3, outerK=1, tol=1e-15) -> np.ndarray:
    """`makefilter(K::Array; ...)` (src/domains.jl:506-516).  `decay` is a Float64 scalar, so the exponent is
    evaluated in Float64 even for Float32 `K`, then converted back to `typeof(K)`."""
    TK = K.dtype
    decay = -math.log(tol) / (outerK - innerK) ** order
    Kd = K.astype(np.float64)
    d = Kd - innerK
    p = d.copy()
    for _ in range(int(order) - 1):  # literal integer power
        p = p * d
    filt = np.exp(-decay * p)
    filt[Kd < innerK] = 1
    return filt.astype(TK)


def _nondimK(g, realvars: bool) -> np.ndarray:
    """src/domains.jl:520-538: nondimensional wavenumber in grid precision T."""
    T = g.T.type
    pi = T(np.pi)
    if g.ndim == 1:
        return (g.kr * g.dx / pi) if realvars else np.abs(g.k * g.dx / pi)
    kx = g.kr if realvars else g.k
    if g.ndim == 2:
        a, b = kx * g.dx / pi, g.l * g.dy / pi
        return np.sqrt(a * a + b * b)
    a, b, c = kx * g.dx / pi, g.l * g.dy / pi, g.m * g.dz / pi
    return np.sqrt(a * a + b * b + c * c)


def makefilter(g, T=None, sz=None, realvars=None, **kwargs) -> np.ndarray:
    """`makefilter(g; realvars, kw...)` and `makefilter(g, T, sz; kw...) = ones(T, sz) .* makefilter(g; realvars=sz[1]==g.nkr)`
    (src/domains.jl:520-541)."""
    if sz is not None:
        realvars = sz[0] == g.nkr
    elif realvars is None:
        realvars = True
    f = makefilter_K(np.asarray(_nondimK(g, realvars)), **kwargs)
    if sz is None:
        return np.asfortranarray(f)
    T = g.T if T is None else np.dtype(T)
    fshape = f.shape + (1,) * (len(sz) - f.ndim)
    return np.asfortranarray(np.ones(sz, dtype=T, order="F") * f.reshape(fshape).astype(T))


# ----------------------------------------------------------------------------- problem.jl
@dataclass
class Equation:
    """src/problem.jl:11-34.  `L`: scalar or array (real or complex); `dims` defaults to `size(L)`;
    `T` defaults to `cxtype(G)`."""
    L: object
    calcN: Callable
    grid: object
    dims: Optional[tuple] = None
    T: Optional[np.dtype] = None

    def __post_init__(self):
        if self.dims is None:
            self.dims = np.shape(self.L)
        self.T = cxtype(self.grid.T) if self.T is None else np.dtype(self.T)


@dataclass
class Clock:
    """src/problem.jl:43-50: `dt` and `t` are stored in the grid float type T."""
    dt: float
    t: float
    step: int


@dataclass
class _Stepper:
    name: str
    filtered: bool
    arrays: dict = field(default_factory=dict)
    filter: Optional[np.ndarray] = None

    def __getattr__(self, k):
        arrs = object.__getattribute__(self, "arrays")
        if k in arrs:
            return arrs[k]
        raise AttributeError(k)


STEPPERS = ["ForwardEuler", "RK4", "LSRK54", "ETDRK4", "AB3", "FilteredForwardEuler", "FilteredRK4",
            "FilteredLSRK54", "FilteredETDRK4", "FilteredAB3"]  # test/runtests.jl:26-37

_EXPLICIT = {"ForwardEuler", "RK4", "AB3", "LSRK54", "FilteredForwardEuler", "FilteredRK4", "FilteredAB3",
             "FilteredLSRK54"}  # src/timesteppers.jl:37-46


def isexplicit(stepper: str) -> bool:
    return stepper in _EXPLICIT


# LSRK54 coefficients as exact rationals converted to T (src/timesteppers.jl:335-350)
LSRK54_A = [Fraction(0), Fraction(-567301805773, 1357537059087), Fraction(-2404267990393, 2016746695238),
            Fraction(-3550918686646, 2091501179385), Fraction(-1275806237668, 842570457699)]
LSRK54_B = [Fraction(1432997174477, 9575080441755), Fraction(5161836677717, 13612068292357),
            Fraction(1720146321549, 2090206949498), Fraction(3134564353537, 4481467310338),
            Fraction(2277821191437, 14882151754819)]
LSRK54_C = [Fraction(0), Fraction(1432997174477, 9575080441755), Fraction(2526269341429, 6820363962896),
            Fraction(2006345519317, 3224310063776), Fraction(2802321613138, 2924317926251)]


def _rat(fr: Fraction, T):
    """Julia `T[num//den]` with T = Complex{Tf}: the rational is converted to the float type Tf."""
    return fltype(T).type(fr.numerator / fr.denominator)


def getexpLs(dt, L):
    """src/timesteppers.jl:673-678 (evaluated in the precision of `dt` and `L`)."""
    return np.exp(dt * L), np.exp(dt * L / 2)


def getetdcoeffs(dt, L, ncirc=32, rcirc=1):
    """src/timesteppers.jl:689-721.  32-point contour mean in Complex{Float64}; `real.()` when `L` is real.
    `dt` keeps its own float type (Float32 `dt` times Float32 `L` is a Float32 product, then promoted)."""
    Larr = np.asarray(L)
    circ = (rcirc * np.exp(2j * np.pi / ncirc * (np.arange(ncirc) + 0.5))).astype(np.complex128)
    dtL = np.asarray(dt * Larr)  # in the precision of dt*L, as in `dt * L .+ circ`
    if dtL.size > (1 << 20):
        return _getetdcoeffs_chunked(dt, dtL, circ, np.iscomplexobj(Larr))
    zc = dtL[..., None].astype(np.complex128) + circ.reshape((1,) * dtL.ndim + (ncirc,))
    ez, ez2 = np.exp(zc), np.exp(zc / 2)
    zc2 = zc * zc
    zc3 = zc2 * zc
    zeta_c = (ez2 - 1) / zc
    alpha_c = (-4 - zc + ez * (4 - 3 * zc + zc2)) / zc3
    beta_c = (2 + zc + ez * (-2 + zc)) / zc3
    gamma_c = (-4 - 3 * zc - zc2 + ez * (4 - zc)) / zc3
    out = []
    for c in (zeta_c, alpha_c, beta_c, gamma_c):
        v = dt * np.mean(c, axis=-1)
        if not np.iscomplexobj(Larr):
            v = v.real
        out.append(np.asfortranarray(v) if np.ndim(v) else v[()])
    return tuple(out)


def _getetdcoeffs_chunked(dt, dtL, circ, cplx):
    """Same arithmetic as `getetdcoeffs`, element by element, for arrays whose (size x 32) contour temporaries would not fit
    host memory (8192^2: 17 GB each): 1 Mi-element chunks on a thread pool (NumPy releases the GIL), the contour mean
    accumulated point by point."""
    from concurrent.futures import ThreadPoolExecutor
    flat = np.ravel(dtL, order="F")
    outs = [np.empty(flat.size, dtype=np.complex128 if cplx else np.float64) for _ in range(4)]
    n = len(circ)

    def work(lo):
        z0 = flat[lo:lo + (1 << 20)].astype(np.complex128)
        acc = [np.zeros(z0.size, dtype=np.complex128) for _ in range(4)]
        for j in range(n):
            zc = z0 + circ[j]
            ez, ez2 = np.exp(zc), np.exp(zc / 2)
            zc2 = zc * zc
            zc3 = zc2 * zc
            acc[0] += (ez2 - 1) / zc
            acc[1] += (-4 - zc + ez * (4 - 3 * zc + zc2)) / zc3
            acc[2] += (2 + zc + ez * (-2 + zc)) / zc3
            acc[3] += (-4 - 3 * zc - zc2 + ez * (4 - zc)) / zc3
        for o, a in zip(outs, acc):
            v = dt * (a / n)
            o[lo:lo + z0.size] = v if cplx else v.real
    with ThreadPoolExecutor(max(1, _WORKERS)) as ex:
        list(ex.map(work, range(0, flat.size, 1 << 20)))
    return tuple(np.asfortranarray(o.reshape(dtL.shape, order="F")) for o in outs)


def TimeStepper(stepper: str, eqn: Equation, dt=None, **filterkwargs) -> _Stepper:
    """`TimeStepper(stepper, equation, dt, dev; kw...)` (src/timesteppers.jl:57-69) and the ten constructors."""
    if stepper not in STEPPERS:
        raise ValueError(f"unknown stepper {stepper}")
    filtered = stepper.startswith("Filtered")
    base = stepper[len("Filtered"):] if filtered else stepper
    z = lambda: zeros(eqn.T, eqn.dims)
    ts = _Stepper(base, filtered)
    if base == "ForwardEuler":
        ts.arrays = dict(N=z())
    elif base == "RK4":
        ts.arrays = dict(sol1=z(), RHS1=z(), RHS2=z(), RHS3=z(), RHS4=z())
    elif base == "LSRK54":
        ts.arrays = dict(S2=z(), RHS=z(), A=[_rat(a, eqn.T) for a in LSRK54_A], B=[_rat(b, eqn.T) for b in LSRK54_B],
                         C=[_rat(c, eqn.T) for c in LSRK54_C])
    elif base == "ETDRK4":
        dtT = fltype(eqn.T).type(dt)  # src/timesteppers.jl:457
        expLdt, exphLdt = getexpLs(dtT, eqn.L)
        zeta, alpha, beta, gamma = getetdcoeffs(dtT, eqn.L)
        ts.arrays = dict(zeta=zeta, alpha=alpha, beta=beta, gamma=gamma, expLdt=expLdt, exphLdt=exphLdt,
                         sol1=z(), sol2=z(), N1=z(), N2=z(), N3=z(), N4=z())
    elif base == "AB3":
        ts.arrays = dict(RHS=z(), RHSm1=z(), RHSm2=z())
    if filtered:
        ts.filter = makefilter(eqn.grid, fltype(eqn.T), eqn.dims, **filterkwargs)  # src/domains.jl:541
    return ts


class Problem:
    """src/problem.jl:99-111."""

    def __init__(self, eqn: Equation, stepper: str, dt, grid, vars=None, params=None, **stepperkwargs):
        T = grid.T.type
        self.clock = Clock(T(dt), T(0), 0)
        self.timestepper = TimeStepper(stepper, eqn, dt, **stepperkwargs)
        self.sol = zeros(eqn.T, eqn.dims)
        self.eqn, self.grid, self.vars, self.params = eqn, grid, vars, params
        self.stepper_name = stepper


# ----------------------------------------------------------------------------- time stepping
_ab3h1, _ab3h2, _ab3h3 = np.float64(23 / 12), np.float64(16 / 12), np.float64(5 / 12)  # src/timesteppers.jl:565-567 (Float64 constants)


def _store(dst, val):
    dst[...] = val  # rounds to dst's dtype like a Julia broadcast assignment


def _rk4substeps(sol, clock, ts, eq, v, p, g, t, dt):
    """src/timesteppers.jl:237-258."""
    L = eq.L
    eq.calcN(ts.RHS1, sol, t, clock, v, p, g)
    _store(ts.RHS1, ts.RHS1 + L * sol)
    _store(ts.sol1, sol + (dt / 2) * ts.RHS1)
    eq.calcN(ts.RHS2, ts.sol1, t + dt / 2, clock, v, p, g)
    _store(ts.RHS2, ts.RHS2 + L * ts.sol1)
    _store(ts.sol1, sol + (dt / 2) * ts.RHS2)
    eq.calcN(ts.RHS3, ts.sol1, t + dt / 2, clock, v, p, g)
    _store(ts.RHS3, ts.RHS3 + L * ts.sol1)
    _store(ts.sol1, sol + dt * ts.RHS3)
    eq.calcN(ts.RHS4, ts.sol1, t + dt, clock, v, p, g)
    _store(ts.RHS4, ts.RHS4 + L * ts.sol1)


def _etdrk4substeps(sol, clock, ts, eq, v, p, g):
    """src/timesteppers.jl:518-537."""
    eq.calcN(ts.N1, sol, clock.t, clock, v, p, g)
    _store(ts.sol1, ts.exphLdt * sol + ts.zeta * ts.N1)
    t2 = clock.t + clock.dt / 2
    eq.calcN(ts.N2, ts.sol1, t2, clock, v, p, g)
    _store(ts.sol2, ts.exphLdt * sol + ts.zeta * ts.N2)
    eq.calcN(ts.N3, ts.sol2, t2, clock, v, p, g)
    _store(ts.sol2, ts.exphLdt * ts.sol1 + ts.zeta * (2 * ts.N3 - ts.N1))
    t3 = clock.t + clock.dt
    eq.calcN(ts.N4, ts.sol2, t3, clock, v, p, g)


def stepforward(prob: Problem, nsteps: int = 1, diags=None) -> None:
    """`stepforward!(prob[, diags], nsteps)` (src/timesteppers.jl:6-35) dispatching on the stepper type."""
    for _ in range(nsteps):
        _step(prob.sol, prob.clock, prob.timestepper, prob.eqn, prob.vars, prob.params, prob.grid)
        if diags is not None:
            for d in (diags if isinstance(diags, (list, tuple)) else [diags]):
                d.increment()


def _step(sol, clock, ts, eq, v, p, g):
    dt, L = clock.dt, eq.L
    name = ts.name
    if name == "ForwardEuler":
        eq.calcN(ts.N, sol, clock.t, clock, v, p, g)
        if ts.filtered:  # src/timesteppers.jl:144 (note `N + L*sol`)
            _store(sol, ts.filter * (sol + dt * (ts.N + L * sol)))
        else:  # :113
            _store(sol, sol + dt * (L * sol + ts.N))
    elif name == "RK4":  # :266-285
        _rk4substeps(sol, clock, ts, eq, v, p, g, clock.t, dt)
        _store(sol, sol + (dt / 6) * (ts.RHS1 + 2 * ts.RHS2 + 2 * ts.RHS3 + ts.RHS4))
        if ts.filtered:
            _store(sol, sol * ts.filter)
    elif name == "LSRK54":  # :383-414
        ts.S2[...] = 0
        t = clock.t
        for i in range(5):
            eq.calcN(ts.RHS, sol, t + ts.C[i] * dt, clock, v, p, g)
            _store(ts.RHS, ts.RHS + L * sol)
            _store(ts.S2, ts.A[i] * ts.S2 + dt * ts.RHS)
            _store(sol, sol + ts.B[i] * ts.S2)
        if ts.filtered:
            _store(sol, sol * ts.filter)
    elif name == "ETDRK4":  # :539-558
        _etdrk4substeps(sol, clock, ts, eq, v, p, g)
        _store(sol, ts.expLdt * sol + ts.alpha * ts.N1 + 2 * ts.beta * (ts.N2 + ts.N3) + ts.gamma * ts.N4)
        if ts.filtered:
            _store(sol, sol * ts.filter)
    elif name == "AB3":  # :628-667
        eq.calcN(ts.RHS, sol, clock.t, clock, v, p, g)
        _store(ts.RHS, ts.RHS + L * sol)
        if clock.step < 3:
            _store(sol, sol + dt * ts.RHS)
        else:
            _store(sol, sol + dt * (_ab3h1 * ts.RHS - _ab3h2 * ts.RHSm1 + _ab3h3 * ts.RHSm2))
        if ts.filtered:
            _store(sol, sol * ts.filter)
    else:
        raise ValueError(name)
    clock.t = type(clock.t)(clock.t + dt)
    clock.step += 1
    if name == "AB3":  # history shifts happen after the clock tick (:644-648)
        ts.RHSm2[...] = ts.RHSm1
        ts.RHSm1[...] = ts.RHS


def step_until(prob: Problem, stop_time) -> None:
    """src/timesteppers.jl:734-760, including the `t_remaining = time_interval - prob.clock.t` quirk (:752)."""
    if prob.timestepper.name == "ETDRK4":
        raise RuntimeError("step_until! requires fully explicit time stepper; does not work with ETDRK4")
    if not stop_time > prob.clock.t:
        raise RuntimeError("stop_time must be greater than prob.clock.t")
    dt = prob.clock.dt
    time_interval = stop_time - prob.clock.t
    nsteps = math.floor(time_interval / dt)
    stepforward(prob, nsteps)
    t_remaining = time_interval - prob.clock.t
    prob.clock.dt = type(dt)(t_remaining)
    stepforward(prob, 1)
    prob.clock.dt = dt


# ----------------------------------------------------------------------------- Diffusion testbed (src/diffusion.jl)
class Diffusion:
    """Namespace mirroring `FourierFlows.Diffusion` (src/diffusion.jl:44-176)."""

    @dataclass
    class Params:
        kappa: object

    @dataclass
    class Vars:
        c: np.ndarray
        cx: np.ndarray
        ch: np.ndarray
        cxh: np.ndarray

    @staticmethod
    def calcN_const(N, sol, t, clock, vars, params, grid):  # :129-133
        N[...] = 0

    @staticmethod
    def calcN_array(N, sol, t, clock, vars, params, grid):  # :135-143
        vars.cxh[...] = (1j * grid.kr) * sol
        grid.rfftplan.ldiv(vars.cx, vars.cxh)
        vars.cx[...] = vars.cx * params.kappa
        grid.rfftplan.mul(vars.cxh, vars.cx)
        N[...] = (1j * grid.kr) * vars.cxh

    @staticmethod
    def Problem(nx=128, Lx=2 * np.pi, kappa=0, dt=0.01, stepper="RK4", aliased_fraction=0, T=np.float64):  # :44-59
        grid = OneDGrid(nx=nx, Lx=Lx, aliased_fraction=aliased_fraction, T=T)
        T = grid.T
        if np.ndim(kappa) == 0:
            params = Diffusion.Params(kappa)
            L = zeros(T, (grid.nkr,))
            L[...] = -kappa * grid.kr * grid.kr  # `@. L = - params.κ * grid.kr^2` (:84)
            eqn = Equation(L, Diffusion.calcN_const, grid)
        else:
            params = Diffusion.Params(np.asarray(kappa))
            eqn = Equation(0, Diffusion.calcN_array, grid, dims=(grid.nkr,), T=cxtype(T))  # :89-90
        vars = Diffusion.Vars(zeros(T, (grid.nx,)), zeros(T, (grid.nx,)), zeros(cxtype(T), (grid.nkr,)),
                              zeros(cxtype(T), (grid.nkr,)))
        return Problem(eqn, stepper, dt, grid, vars, params)

    @staticmethod
    def updatevars(prob):  # :150-160
        v, g, sol = prob.vars, prob.grid, prob.sol
        v.ch[...] = sol
        v.cxh[...] = (1j * g.kr) * sol
        g.rfftplan.ldiv(v.c, v.ch.copy())
        g.rfftplan.ldiv(v.cx, v.cxh.copy())

    @staticmethod
    def set_c(prob, c):  # :167-176
        prob.vars.c[...] = c
        prob.grid.rfftplan.mul(prob.sol, prob.vars.c)
        Diffusion.updatevars(prob)


# ----------------------------------------------------------------------------- benchmark equations
class TwoDNavierStokes:
    """2-D vorticity equation `calcN_advection!` of the child package the reference points to (README.md:81-83;
    GeophysicalFlows TwoDNavierStokes, not in the reference tree) restated as in SURVEY 8d C3, followed by
    `dealias!(N, grid)`; L = -nu * Krsq."""

    @dataclass
    class Params:
        nu: float

    @dataclass
    class Vars:
        zeta: np.ndarray
        u: np.ndarray
        v: np.ndarray
        zetah: np.ndarray
        uh: np.ndarray
        vh: np.ndarray

    @staticmethod
    def calcN(N, sol, t, clock, vars, params, grid):
        vars.uh[...] = ((1j * grid.l) * grid.invKrsq) * sol
        vars.vh[...] = ((-1j * grid.kr) * grid.invKrsq) * sol
        vars.zetah[...] = sol
        grid.rfftplan.ldiv(vars.u, vars.uh)
        grid.rfftplan.ldiv(vars.v, vars.vh)
        grid.rfftplan.ldiv(vars.zeta, vars.zetah)
        vars.u[...] = vars.u * vars.zeta
        vars.v[...] = vars.v * vars.zeta
        grid.rfftplan.mul(vars.uh, vars.u)
        grid.rfftplan.mul(vars.vh, vars.v)
        N[...] = (-1j * grid.kr) * vars.uh - (1j * grid.l) * vars.vh
        dealias(N, grid)

    @staticmethod
    def Problem(nx=256, Lx=2 * np.pi, ny=None, Ly=None, nu=0.0, dt=0.01, stepper="RK4", aliased_fraction=1 / 3, T=np.float64,
                **stepperkwargs):
        grid = TwoDGrid(nx=nx, Lx=Lx, ny=ny, Ly=Ly, aliased_fraction=aliased_fraction, T=T)
        T = grid.T
        cT = cxtype(T)
        vars = TwoDNavierStokes.Vars(*(zeros(T, (grid.nx, grid.ny)) for _ in range(3)), *(zeros(cT, (grid.nkr, grid.nl)) for _ in range(3)))
        L = np.asfortranarray((T.type(-nu) * grid.Krsq).astype(T))
        eqn = Equation(L, TwoDNavierStokes.calcN, grid)
        return Problem(eqn, stepper, dt, grid, vars, TwoDNavierStokes.Params(nu), **stepperkwargs)


class Burgers3D:
    """Builder-defined 3-D test equation of SURVEY 8d C4/C5: N = -1/2 im kr rfft(irfft(sol)^2), `dealias!(N, grid)`,
    L = -kappa * Krsq (the reference's Diffusion module is 1-D only)."""

    @dataclass
    class Params:
        kappa: float

    @dataclass
    class Vars:
        c: np.ndarray
        ch: np.ndarray

    @staticmethod
    def calcN(N, sol, t, clock, vars, params, grid):
        grid.rfftplan.ldiv(vars.c, sol)
        vars.c[...] = vars.c * vars.c
        grid.rfftplan.mul(vars.ch, vars.c)
        N[...] = (-0.5j * grid.kr) * vars.ch
        dealias(N, grid)

    @staticmethod
    def Problem(nx=64, Lx=2 * np.pi, ny=None, nz=None, kappa=1e-3, dt=1e-3, stepper="FilteredRK4", aliased_fraction=1 / 3,
                T=np.float64, **stepperkwargs):
        grid = ThreeDGrid(nx=nx, Lx=Lx, ny=ny, nz=nz, aliased_fraction=aliased_fraction, T=T)
        T = grid.T
        vars = Burgers3D.Vars(zeros(T, grid.shape), zeros(cxtype(T), (grid.nkr, grid.nl, grid.nm)))
        L = np.asfortranarray((T.type(-kappa) * grid.Krsq).astype(T))
        eqn = Equation(L, Burgers3D.calcN, grid)
        return Problem(eqn, stepper, dt, grid, vars, Burgers3D.Params(kappa), **stepperkwargs)


def jacobianh(a, b, grid):
    """src/utils.jl:190-205 on a TwoDGrid (real fields through the rfft plan, complex ones through the fft plan)."""
    if not np.iscomplexobj(a):
        bh = grid.rfftplan * b
        bx = grid.rfftplan.solve((1j * grid.kr) * bh)
        by = grid.rfftplan.solve((1j * grid.l) * bh)
        return (1j * grid.kr) * (grid.rfftplan * (a * by)) - (1j * grid.l) * (grid.rfftplan * (a * bx))
    bh = grid.fftplan * b
    bx = grid.fftplan.solve((1j * grid.k) * bh)
    by = grid.fftplan.solve((1j * grid.l) * bh)
    return (1j * grid.k) * (grid.fftplan * (a * by)) - (1j * grid.l) * (grid.fftplan * (a * bx))


def jacobian(a, b, grid):
    """src/utils.jl:212-218."""
    jh = jacobianh(a, b, grid)
    return grid.fftplan.solve(jh) if np.iscomplexobj(a) else grid.rfftplan.solve(jh)


def random_phase_field(shape, Lext, K0, slope=1.0, seed=1234, T=np.float64):
    """Synthetic random-phase real field: |f_hat| ~ K^slope * exp(-(K/K0)^2), phases U[0, 2pi),
    Hermitian-consistent by construction (generated with irfftn on the host), normalised to rms 1."""
    rng = np.random.default_rng(seed)
    nd = len(shape)
    Lext = (Lext,) * nd if np.isscalar(Lext) else tuple(Lext)
    ks = [rfftfreq(shape[0], 2 * np.pi / Lext[0] * shape[0])] + [fftfreq(shape[d], 2 * np.pi / Lext[d] * shape[d]) for d in range(1, nd)]
    K2 = np.zeros((shape[0] // 2 + 1,) + tuple(shape[1:]))
    for d, k in enumerate(ks):
        sh = [1] * nd
        sh[d] = len(k)
        K2 = K2 + (k * k).reshape(sh)
    K = np.sqrt(K2)
    with np.errstate(divide="ignore", invalid="ignore"):
        amp = np.where(K > 0, K ** slope * np.exp(-(K / K0) ** 2), 0.0)
    fh = amp * np.exp(2j * np.pi * rng.random(K.shape))
    axes = tuple(range(nd - 1, -1, -1))
    f = sfft.irfftn(np.asfortranarray(fh), s=tuple(shape[ax] for ax in axes), axes=axes, workers=_WORKERS)
    f = f / np.sqrt(np.mean(f * f))
    return np.asfortranarray(f.astype(T))
