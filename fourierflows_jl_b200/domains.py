"""Grids, FFT plans, `dealias!` and `makefilter` -- host mirror of /root/reference/src/domains.jl.

Grid scalars and index ranges are plain host values (as in the reference); wavenumber vectors, the dense
`Ksq/invKsq/Krsq/invKrsq` arrays, the FFT plans and everything executed per step live on the B200 behind the C ABI.
The dense arrays are created lazily on first access (24 GiB at 1024^3 Float64 in the reference).
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from . import _lib as L
from .array import GPU, DevArray, _require_gpu, cxtype, ffb_dtype, fltype

DomainError = L.DomainError


def getaliasedwavenumbers(nk, nkr, aliased_fraction):
    """src/domains.jl:408-421 (Float64 arithmetic; 1-based inclusive ranges as (lo, hi) tuples or None)."""
    Lf = (1 - aliased_fraction) / 2
    Rf = (1 + aliased_fraction) / 2
    iL = math.floor(Lf * nk) + 1
    iR = math.ceil(Rf * nk)
    if not aliased_fraction < 1:
        raise L.FFBError(L.FFB_EINVAL, "`aliased_fraction` must be less than 1")
    if aliased_fraction > 0:
        return (iL, iR), (iL, nkr)
    return None, None


class Plan:
    """`grid.rfftplan` / `grid.fftplan` (src/domains.jl:86-87,207-208,348-349): AbstractFFTs plan protocol.

    `mul(out, a)` = `mul!(out, plan, a)` forward unnormalised; `ldiv(out, ah)` = `ldiv!(out, plan, ah)` inverse scaled
    by 1/N; `plan * a` and `plan.solve(ah)` (= `plan \\ ah`) allocate."""

    def __init__(self, shape, T, kind, nbatch=1, flags=L.FFB_PLAN_DEFAULT):
        self.shape = tuple(int(s) for s in shape)
        self.T = np.dtype(T)
        self.kind = kind
        self.nbatch = nbatch
        n = (C.c_int64 * 3)(*self.shape, *([1] * (3 - len(self.shape))))
        h = C.c_void_p()
        L.call("ffb_plan_create", C.byref(h), len(self.shape), n, ffb_dtype(self.T), kind, nbatch, flags)
        self._h = h

    @property
    def spectral_shape(self):
        s0 = self.shape[0] // 2 + 1 if self.kind == L.FFB_R2C else self.shape[0]
        return (s0,) + self.shape[1:] + ((self.nbatch,) if self.nbatch > 1 else ())

    @property
    def physical_shape(self):
        return self.shape + ((self.nbatch,) if self.nbatch > 1 else ())

    def describe(self):
        buf = C.create_string_buffer(512)
        L.call("ffb_plan_describe", self._h, buf, 512)
        return buf.value.decode()

    def mul(self, out: DevArray, a: DevArray):
        L.call("ffb_fft_forward", self._h, a.ptr, out.ptr)
        return out

    def ldiv(self, out: DevArray, ah: DevArray):
        L.call("ffb_fft_inverse", self._h, ah.ptr, out.ptr)
        return out

    # fused forms (ffb_fft_forward_ex / ffb_fft_inverse_ex): spectral multiplies, products and dealias folded into the passes
    @staticmethod
    def _fuse(coef=1.0, kx=None, l=None, m=None, w=None, acc=None, acoef=0.0, akx=None, al=None, am=None, alias=None, mul=None, square=False):
        f = L.ffb_fuse()
        c, a = complex(coef), complex(acoef)
        f.cr, f.ci, f.ar, f.ai = c.real, c.imag, a.real, a.imag
        ptr = lambda x: x.ptr if x is not None else None
        f.kx, f.l, f.m, f.w, f.acc, f.akx, f.al, f.am, f.mul = (ptr(v) for v in (kx, l, m, w, acc, akx, al, am, mul))
        f.dealias = 1 if alias is not None else 0
        al3 = list(alias) + [None] * (3 - len(alias)) if alias is not None else [None] * 3
        f.alias_lo = (C.c_int32 * 3)(*[(r[0] if r else 0) for r in al3])
        f.alias_hi = (C.c_int32 * 3)(*[(r[1] if r else 0) for r in al3])
        f.square_input = 1 if square else 0
        return f

    def ldiv_ex(self, out: DevArray, ah: DevArray, coef=1.0, kx=None, l=None, m=None, w=None, mul=None):
        """`out = irfft((coef * kx * l * m * w) .* ah) .* mul` in one pass chain."""
        f = self._fuse(coef=coef, kx=kx, l=l, m=m, w=w, mul=mul)
        L.call("ffb_fft_inverse_ex", self._h, ah.ptr, out.ptr, C.byref(f))
        return out

    def ldiv_multi(self, outs, ah: DevArray, variants):
        """`outs[v] = irfft((coef_v * kx_v * l_v * m_v * w_v) .* ah) .* mul_v` for every keyword dict of `variants`, in order, sharing
        the read of `ah` (ffb_fft_inverse_multi)."""
        n = len(outs)
        if n != len(variants):
            raise ValueError("one keyword dict per output")
        fuses = (L.ffb_fuse * n)(*[self._fuse(**kw) for kw in variants])
        ptrs = (C.c_void_p * n)(*[o.ptr for o in outs])
        L.call("ffb_fft_inverse_multi", self._h, ah.ptr, n, ptrs, fuses)
        return outs

    def mul_ex(self, out: DevArray, a: DevArray, coef=1.0, kx=None, l=None, m=None, w=None, acc=None, acoef=0.0, akx=None, al=None, am=None,
               alias=None, square=False):
        """`out = dealias!((coef * kx * l * m * w) .* rfft(a) + (acoef * akx * al * am) .* acc)`; `alias` = per-dimension
        1-based ranges (kralias, lalias[, malias]) or None; `square`: transform `a.^2`."""
        f = self._fuse(coef=coef, kx=kx, l=l, m=m, w=w, acc=acc, acoef=acoef, akx=akx, al=al, am=am, alias=alias, square=square)
        L.call("ffb_fft_forward_ex", self._h, a.ptr, out.ptr, C.byref(f))
        return out

    def __mul__(self, a: DevArray):
        return self.mul(DevArray(self.spectral_shape, cxtype(self.T)), a)

    def solve(self, ah: DevArray):
        dt = self.T if self.kind == L.FFB_R2C else cxtype(self.T)
        return self.ldiv(DevArray(self.physical_shape, dt), ah)

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                L.load().ffb_plan_destroy(self._h)
                self._h = None
        except Exception:
            pass


def mul_(out, plan, a):
    """`mul!(out, plan, a)`."""
    return plan.mul(out, a)


def ldiv_(out, plan, ah):
    """`ldiv!(out, plan, ah)`."""
    return plan.ldiv(out, ah)


def _wavenumbers(n, Lext, T, real_half):
    out = DevArray((n // 2 + 1 if real_half else n,), T)
    L.call("ffb_wavenumbers", out.ptr, n, float(Lext), ffb_dtype(T), 1 if real_half else 0)
    return out


def _range(x0, dx, n, T):
    """`range(T(x0), step=T(dx), length=n)` (src/domains.jl:74): StepRangeLen on TwicePrecision, rounded once."""
    T = np.dtype(T).type
    wide = np.longdouble
    return (wide(T(x0)) + np.arange(n).astype(wide) * wide(T(dx))).astype(T)


class AbstractGrid:
    ndim = 0

    def make_desc(self, dims, kx_alias=None):
        """ffb_desc for an array of extents `dims` = (n0[, n1[, n2]][, nfields]) on this grid."""
        d = L.ffb_desc()
        d.ndim = self.ndim
        dd = list(dims[: self.ndim]) + [1] * (3 - self.ndim)
        nf = 1
        for s in dims[self.ndim:]:
            nf *= s
        dd.append(nf)
        d.dims = (C.c_int64 * 4)(*dd)
        d.dtype = ffb_dtype(self.T)
        al = [kx_alias, getattr(self, "lalias", None) if self.ndim >= 2 else None,
              getattr(self, "malias", None) if self.ndim >= 3 else None]
        lo = [(a[0] if a else 0) for a in al]
        hi = [(a[1] if a else 0) for a in al]
        d.alias_lo = (C.c_int32 * 3)(*lo)
        d.alias_hi = (C.c_int32 * 3)(*hi)
        return d

    def _dense(self, name):
        """Lazily materialised dense `Ksq/invKsq/Krsq/invKrsq` (src/domains.jl:197-203,338-344)."""
        cache = self.__dict__.setdefault("_dense_cache", {})
        if name in cache:
            return cache[name]
        real_half = name in ("Krsq", "invKrsq", "invkrsq")
        kx = self.kr if real_half else self.k
        dims = (kx.shape[0],) + tuple(self.shape[1:])
        desc = self.make_desc(dims)
        ksq = DevArray(dims, self.T)
        inv = DevArray(dims, self.T)
        L.call("ffb_ksq", ksq.ptr, inv.ptr, kx.ptr, self.l.ptr if self.ndim >= 2 else None,
               self.m.ptr if self.ndim >= 3 else None, C.byref(desc))
        if real_half:
            cache["Krsq"], cache["invKrsq"], cache["invkrsq"] = ksq, inv, inv
        else:
            cache["Ksq"], cache["invKsq"], cache["invksq"] = ksq, inv, inv
        return cache[name]

    Ksq = property(lambda self: self._dense("Ksq"))
    invKsq = property(lambda self: self._dense("invKsq"))
    Krsq = property(lambda self: self._dense("Krsq"))
    invKrsq = property(lambda self: self._dense("invKrsq"))

    def __repr__(self):
        return f"{type(self).__name__}(T={self.T}, shape={self.shape}, aliased_fraction={self.aliased_fraction})"


class OneDGrid(AbstractGrid):
    """`OneDGrid(dev; nx, Lx, x0=-Lx/2, T=Float64, aliased_fraction=1/3)` (src/domains.jl:61-101)."""

    ndim = 1

    def __init__(self, dev=None, *, nx, Lx, x0=None, nthreads=None, effort=None, T=np.float64, aliased_fraction=1 / 3):
        dev = GPU() if dev is None else dev
        _require_gpu(dev)
        if nx % 2 != 0:
            raise DomainError("nx must be even")
        T = np.dtype(T)
        x0 = -Lx / 2 if x0 is None else x0
        self.device, self.T = dev, T
        self.nx, self.nk, self.nkr = nx, nx, nx // 2 + 1
        self.dx, self.Lx = T.type(Lx / nx), T.type(Lx)
        self.x = _range(x0, Lx / nx, nx, T)
        self.k = _wavenumbers(nx, Lx, T, False)
        self.kr = _wavenumbers(nx, Lx, T, True)
        self.fftplan = Plan((nx,), T, L.FFB_C2C)
        self.rfftplan = Plan((nx,), T, L.FFB_R2C)
        self.aliased_fraction = T.type(aliased_fraction)
        self.kalias, self.kralias = getaliasedwavenumbers(self.nk, self.nkr, aliased_fraction)

    shape = property(lambda self: (self.nx,))
    invksq = property(lambda self: self._dense("invksq"))
    invkrsq = property(lambda self: self._dense("invkrsq"))


class TwoDGrid(AbstractGrid):
    """`TwoDGrid(dev; nx, Lx, ny=nx, Ly=Lx, ...)` (src/domains.jl:175-223)."""

    ndim = 2

    def __init__(self, dev=None, *, nx, Lx, ny=None, Ly=None, x0=None, y0=None, nthreads=None, effort=None, T=np.float64,
                 aliased_fraction=1 / 3):
        dev = GPU() if dev is None else dev
        _require_gpu(dev)
        ny = nx if ny is None else ny
        Ly = Lx if Ly is None else Ly
        if nx % 2 != 0 or ny % 2 != 0:
            raise DomainError("nx and ny must be even")
        T = np.dtype(T)
        x0 = -Lx / 2 if x0 is None else x0
        y0 = -Ly / 2 if y0 is None else y0
        self.device, self.T = dev, T
        self.nx, self.ny, self.nk, self.nl, self.nkr = nx, ny, nx, ny, nx // 2 + 1
        self.dx, self.dy, self.Lx, self.Ly = T.type(Lx / nx), T.type(Ly / ny), T.type(Lx), T.type(Ly)
        self.x, self.y = _range(x0, Lx / nx, nx, T), _range(y0, Ly / ny, ny, T)
        self.k = _wavenumbers(nx, Lx, T, False)
        self.l = _wavenumbers(ny, Ly, T, False)
        self.kr = _wavenumbers(nx, Lx, T, True)
        self.fftplan = Plan((nx, ny), T, L.FFB_C2C)
        self.rfftplan = Plan((nx, ny), T, L.FFB_R2C)
        self.aliased_fraction = T.type(aliased_fraction)
        self.kalias, self.kralias = getaliasedwavenumbers(self.nk, self.nkr, aliased_fraction)
        self.lalias, _ = getaliasedwavenumbers(self.nl, self.nl, aliased_fraction)

    shape = property(lambda self: (self.nx, self.ny))


class ThreeDGrid(AbstractGrid):
    """`ThreeDGrid(dev; nx, Lx, ny=nx, Ly=Lx, nz=nx, Lz=Lx, ...)` (src/domains.jl:311-366)."""

    ndim = 3

    def __init__(self, dev=None, *, nx, Lx, ny=None, Ly=None, nz=None, Lz=None, x0=None, y0=None, z0=None, nthreads=None,
                 effort=None, T=np.float64, aliased_fraction=1 / 3):
        dev = GPU() if dev is None else dev
        _require_gpu(dev)
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
        self.device, self.T = dev, T
        self.nx, self.ny, self.nz = nx, ny, nz
        self.nk, self.nl, self.nm, self.nkr = nx, ny, nz, nx // 2 + 1
        self.dx, self.dy, self.dz = T.type(Lx / nx), T.type(Ly / ny), T.type(Lz / nz)
        self.Lx, self.Ly, self.Lz = T.type(Lx), T.type(Ly), T.type(Lz)
        self.x, self.y, self.z = _range(x0, Lx / nx, nx, T), _range(y0, Ly / ny, ny, T), _range(z0, Lz / nz, nz, T)
        self.k = _wavenumbers(nx, Lx, T, False)
        self.l = _wavenumbers(ny, Ly, T, False)
        self.m = _wavenumbers(nz, Lz, T, False)
        self.kr = _wavenumbers(nx, Lx, T, True)
        self.fftplan = Plan((nx, ny, nz), T, L.FFB_C2C)
        self.rfftplan = Plan((nx, ny, nz), T, L.FFB_R2C)
        self.aliased_fraction = T.type(aliased_fraction)
        self.kalias, self.kralias = getaliasedwavenumbers(self.nk, self.nkr, aliased_fraction)
        self.lalias, _ = getaliasedwavenumbers(self.nl, self.nl // 2 + 1, aliased_fraction)
        self.malias, _ = getaliasedwavenumbers(self.nm, self.nm // 2 + 1, aliased_fraction)

    shape = property(lambda self: (self.nx, self.ny, self.nz))


def gridpoints(grid):
    """`gridpoints(grid)` (src/domains.jl:379-398): meshgrid arrays on the device."""
    if grid.ndim == 1:
        return DevArray.from_numpy(grid.x)
    if grid.ndim == 2:
        X, Y = np.meshgrid(grid.x, grid.y, indexing="ij")
        return DevArray.from_numpy(X), DevArray.from_numpy(Y)
    X, Y, Z = np.meshgrid(grid.x, grid.y, grid.z, indexing="ij")
    return DevArray.from_numpy(X), DevArray.from_numpy(Y), DevArray.from_numpy(Z)


def dealias(fh: DevArray, grid) -> None:
    """`dealias!(fh, grid)` (src/domains.jl:428-476).  No-op for grids built with `aliased_fraction = 0` (:434);
    `kralias` vs `kalias` chosen from `size(fh, 1) == grid.nkr` (:437,450,464); returns None."""
    if grid.kalias is None:
        return None
    kal = grid.kralias if fh.shape[0] == grid.nkr else grid.kalias
    desc = grid.make_desc(fh.shape, kx_alias=kal)
    L.call("ffb_dealias", fh.ptr, C.byref(desc))
    return None


def makefilter(arg, T=None, sz=None, *, realvars=None, order=4, innerK=2 / 3, outerK=1, tol=1e-15):
    """`makefilter(grid; realvars, kw...)`, `makefilter(grid, T, sz; kw...)` and `makefilter(equation; kw...)`
    (src/domains.jl:506-546).  The dense filter of `sz` is evaluated on the device."""
    if hasattr(arg, "grid") and hasattr(arg, "dims"):  # makefilter(equation)
        return makefilter(arg.grid, fltype(arg.T), arg.dims, order=order, innerK=innerK, outerK=outerK, tol=tol)
    g = arg
    if sz is not None:
        realvars = sz[0] == g.nkr
    elif realvars is None:
        realvars = True
    kx = g.kr if realvars else g.k
    if sz is None:
        sz = (kx.shape[0],) + tuple(g.shape[1:])
    T = g.T if T is None else np.dtype(T)
    if T != g.T:
        raise L.FFBError(L.FFB_EUNSUPPORTED, "filter type must match the grid's float type")
    desc = g.make_desc(sz)
    out = DevArray(sz, T)
    L.call("ffb_make_filter", out.ptr, kx.ptr, g.l.ptr if g.ndim >= 2 else None, g.m.ptr if g.ndim >= 3 else None,
           float(g.dx), float(g.dy) if g.ndim >= 2 else 0.0, float(g.dz) if g.ndim >= 3 else 0.0,
           float(order), float(innerK), float(outerK), float(tol), C.byref(desc))
    return out
