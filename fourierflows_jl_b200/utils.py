"""Elementwise vocabulary for user `calcN!` code and the Parseval sums of /root/reference/src/utils.jl:113-183.

Julia `@.` broadcasts on device arrays are lowered to this closed set of fused kernels; anything else must be
expressed through them (there is no CPU fallback)."""
from __future__ import annotations

import ctypes as C

from . import _lib as L
from .array import DevArray, ffb_dtype


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
