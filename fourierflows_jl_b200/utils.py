"""Elementwise vocabulary for user `calcN!` code and the Parseval sums of /root/reference/src/utils.jl:113-183.

Julia `@.` broadcasts on device arrays are lowered to this closed set of fused kernels; anything else must be
expressed through them (there is no CPU fallback)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L
from .array import DevArray, cxtype, ffb_dtype, fltype


def axpby(out: DevArray, a, x: DevArray, b=0.0, y: DevArray = None):
    """`@. out = a*x + b*y` (real or complex arrays of one type; y optional)."""
    L.call("ffb_ew_axpby", out.ptr, float(a), x.ptr, float(b), y.ptr if y is not None else None,
           1 if out.dtype.kind == "c" else 0, ffb_dtype(out.dtype), out.size)
    return out


def mul_real(out: DevArray, x: DevArray, y: DevArray):
    """`@. out = x * y` for real physical-space arrays (`@. vars.cx *= params.κ`, src/diffusion.jl:138)."""
    L.call("ffb_ew_mul_real", out.ptr, x.ptr, y.ptr, ffb_dtype(out.dtype), out.size)
    return out


def spectral_mul(out: DevArray, inp: DevArray, grid, coef=1.0, px=0, py=0, pz=0, w: DevArray = None, accumulate=False,
                 dealias=False):
    """`@. out (+)= coef * kx^px * l^py * m^pz * w * inp` with `coef` a complex scalar (e.g. `im * grid.kr * sol`,
    src/diffusion.jl:136), optionally followed by `dealias!(out, grid)`.  kx = kr when `size(inp,1) == grid.nkr`."""
    half = inp.shape[0] == grid.nkr
    kx = grid.kr if half else grid.k
    kal = None
    if dealias and grid.kalias is not None:
        kal = grid.kralias if half else grid.kalias
    desc = grid.make_desc(inp.shape, kx_alias=kal)
    c = complex(coef)
    L.call("ffb_ew_spectral_mul", out.ptr, inp.ptr, c.real, c.imag, kx.ptr, px, grid.l.ptr if grid.ndim >= 2 else None, py,
           grid.m.ptr if grid.ndim >= 3 else None, pz, w.ptr if w is not None else None, 1 if accumulate else 0,
           1 if (dealias and grid.kalias is not None) else 0, C.byref(desc))
    return out


def _psum(uh: DevArray, grid, abs2: bool):
    half = uh.shape[0] == grid.nkr
    desc = grid.make_desc(uh.shape)
    r = C.c_double(0.0)
    L.call("ffb_parseval_sum", C.byref(r), uh.ptr, 1 if abs2 else 0, 1 if half else 0, C.byref(desc))
    return r.value


def parsevalsum2(uh: DevArray, grid):
    """`parsevalsum2(uh, grid)` (src/utils.jl:113-139; 1-D and 2-D grids like the reference)."""
    s = _psum(uh, grid, True)
    if grid.ndim == 1:
        return s * float(grid.Lx) / grid.nx ** 2
    if grid.ndim == 2:
        return s * float(grid.Lx) * float(grid.Ly) / (grid.nx ** 2 * grid.ny ** 2)
    raise L.FFBError(L.FFB_EUNSUPPORTED, "parsevalsum2 is defined for OneDGrid and TwoDGrid (src/utils.jl:113-139)")


def parsevalsum(uh: DevArray, grid):
    """`parsevalsum(uh, grid)` (src/utils.jl:157-183)."""
    s = _psum(uh, grid, False)
    if grid.ndim == 1:
        return s * float(grid.Lx) / grid.nx ** 2
    if grid.ndim == 2:
        return s * float(grid.Lx) * float(grid.Ly) / (grid.nx ** 2 * grid.ny ** 2)
    raise L.FFBError(L.FFB_EUNSUPPORTED, "parsevalsum is defined for OneDGrid and TwoDGrid (src/utils.jl:157-183)")


def mul(out: DevArray, x: DevArray, y: DevArray):
    """`@. out = x * y` for same-shape real / complex arrays (a real factor scales both parts of a complex one)."""
    L.call("ffb_ew_mul", out.ptr, x.ptr, 1 if x.dtype.kind == "c" else 0, y.ptr, 1 if y.dtype.kind == "c" else 0, ffb_dtype(out.dtype), out.size)
    return out


def jacobianh(a: DevArray, b: DevArray, grid):
    """`jacobianh(a, b, grid)` (src/utils.jl:190-205): Fourier transform of the Jacobian J(a, b) = d(a b_y)/dx - d(a b_x)/dy on a
    TwoDGrid.  Real fields: five transforms with every multiply folded into a pass (`ffb_jacobianh`); complex fields: the
    reference's expression on the c2c plan."""
    import numpy as np
    if grid.ndim != 2:
        raise L.FFBError(L.FFB_EUNSUPPORTED, "jacobianh is defined for TwoDGrid (src/utils.jl:190)")
    if a.dtype.kind != "c":
        cT = np.complex64 if a.dtype == np.float32 else np.complex128
        out, sh = DevArray((grid.nkr, grid.nl), cT), DevArray((grid.nkr, grid.nl), cT)
        p1, p2 = DevArray(a.shape, a.dtype), DevArray(a.shape, a.dtype)
        L.call("ffb_jacobianh", grid.rfftplan._h, out.ptr, a.ptr, b.ptr, grid.kr.ptr, grid.l.ptr, sh.ptr, p1.ptr, p2.ptr)
        return out
    plan = grid.fftplan
    bh = plan * b
    t = DevArray(bh.shape, bh.dtype)
    bx = plan.solve(spectral_mul(t, bh, grid, coef=1j, px=1))
    by = plan.solve(spectral_mul(t, bh, grid, coef=1j, py=1))
    aby, abx = mul(DevArray(a.shape, a.dtype), a, by), mul(DevArray(a.shape, a.dtype), a, bx)
    out = spectral_mul(DevArray(bh.shape, bh.dtype), plan * aby, grid, coef=1j, px=1)
    return spectral_mul(out, plan * abx, grid, coef=-1j, py=1, accumulate=True)


def jacobian(a: DevArray, b: DevArray, grid):
    """`jacobian(a, b, grid)` (src/utils.jl:212-218)."""
    jh = jacobianh(a, b, grid)
    return (grid.rfftplan if a.dtype.kind != "c" else grid.fftplan).solve(jh)


# ---------------------------------------------------------------- `fft / ifft / rfft / irfft` (re-exported FFTW names, src/FourierFlows.jl:72)
# Allocating whole-array transforms; the plan of a (shape, type, kind) is kept for the next call (a few, least recently used dropped:
# a plan owns twiddle tables and scratch).
_FREE_PLANS: dict = {}
_FREE_PLANS_MAX = 4


def _free_plan(shape, T, kind):
    from .domains import Plan
    key = (tuple(int(v) for v in shape), np.dtype(T).str, int(kind))
    plan = _FREE_PLANS.pop(key, None)
    if plan is None:
        plan = Plan(key[0], T, kind)
    _FREE_PLANS[key] = plan                     # most recently used last
    while len(_FREE_PLANS) > _FREE_PLANS_MAX:
        _FREE_PLANS.pop(next(iter(_FREE_PLANS)))
    return plan


def rfft(a: DevArray) -> DevArray:
    """`rfft(a)`: real (nx, ny, nz) -> complex (nx/2+1, ny, nz), unnormalised, all dimensions"""
    if np.dtype(a.dtype).kind != "f":
        raise TypeError("rfft needs a real array")
    return _free_plan(a.shape, a.dtype, L.FFB_R2C) * a


def irfft(ah: DevArray, nx: int) -> DevArray:
    """`irfft(ah, nx)`: complex (nx/2+1, ny, nz) -> real (nx, ny, nz), scaled by 1/(nx*ny*nz); `nx` = first physical dimension"""
    if np.dtype(ah.dtype).kind != "c" or ah.shape[0] != int(nx) // 2 + 1:
        raise ValueError("irfft(ah, nx): ah must be complex with size(ah, 1) == nx/2 + 1")
    return _free_plan((int(nx),) + tuple(ah.shape[1:]), fltype(ah.dtype), L.FFB_R2C).solve(ah)


def fft(a: DevArray) -> DevArray:
    """`fft(a)` of a complex array, all dimensions, unnormalised"""
    if np.dtype(a.dtype).kind != "c":
        raise TypeError("fft needs a complex array (convert a real field first, or use rfft)")
    return _free_plan(a.shape, fltype(a.dtype), L.FFB_C2C) * a


def ifft(ah: DevArray) -> DevArray:
    """`ifft(ah)` of a complex array, all dimensions, scaled by 1/N"""
    if np.dtype(ah.dtype).kind != "c":
        raise TypeError("ifft needs a complex array")
    return _free_plan(ah.shape, fltype(ah.dtype), L.FFB_C2C).solve(ah)
