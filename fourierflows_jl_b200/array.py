"""Device arrays: the B1 seam (`zeros(GPU(), T, dims)` src/utils.jl:80, `device_array(GPU())` src/utils.jl:330,
upload `device_array(dev){T}(host)` src/domains.jl:77, download `Array(x)` src/output.jl:79).

A `DevArray` owns one dense column-major device buffer allocated by `ffb_malloc`; the shape is in Julia order
(x fastest).  Host interchange uses Fortran-ordered NumPy arrays, so bytes match Julia's `Array` exactly.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L


class Device:
    pass


class GPU(Device):
    """`GPU()` of src/FourierFlows.jl:94-101."""

    def __repr__(self):
        return "GPU"


class CPU(Device):
    """Exists only so that `OneDGrid(CPU(); ...)` fails loudly: this library has no CPU path."""

    def __repr__(self):
        return "CPU"


def _require_gpu(dev):
    if not isinstance(dev, GPU):
        raise L.FFBError(L.FFB_EUNSUPPORTED, "fourierflows.jl_b200 runs on GPU() only; there is no CPU fallback")


def ffb_dtype(T) -> int:
    T = np.dtype(T)
    if T in (np.dtype(np.float64), np.dtype(np.complex128)):
        return L.FFB_F64
    if T in (np.dtype(np.float32), np.dtype(np.complex64)):
        return L.FFB_F32
    raise TypeError(f"unsupported element type {T}")


def cxtype(T):
    T = np.dtype(T)
    return T if T.kind == "c" else np.dtype(np.complex64 if T == np.float32 else np.complex128)


def fltype(T):
    T = np.dtype(T)
    return T if T.kind == "f" else np.dtype(np.float32 if T == np.complex64 else np.float64)


class DevArray:
    __slots__ = ("ptr", "shape", "dtype", "_owner", "_owns", "__weakref__")

    def __init__(self, shape, dtype, ptr=None, owner=None):
        self.shape = tuple(int(s) for s in (shape if isinstance(shape, (tuple, list)) else (shape,)))
        self.dtype = np.dtype(dtype)
        self._owner = owner  # keeps the allocation a view points into alive
        self._owns = False
        if ptr is None:
            p = C.c_void_p()
            L.call("ffb_malloc", C.byref(p), self.nbytes)
            self.ptr = p.value
            self._owns = True
        else:
            self.ptr = ptr  # borrowed pointer (a view, or a buffer owned by a library handle)

    # --- construction
    @classmethod
    def zeros(cls, dtype, shape):
        a = cls(shape, dtype)
        L.call("ffb_memset_zero", a.ptr, a.nbytes)
        return a

    @classmethod
    def from_numpy(cls, host):
        host = np.asarray(host)
        a = cls(host.shape, host.dtype)
        a.copy_from_host(host)
        return a

    # --- properties
    @property
    def size(self):
        n = 1
        for s in self.shape:
            n *= s
        return n

    @property
    def nbytes(self):
        return self.size * self.dtype.itemsize

    @property
    def ndim(self):
        return len(self.shape)

    # --- transfers
    def copy_from_host(self, host):
        h = np.asfortranarray(host, dtype=self.dtype)
        if h.shape != self.shape:
            raise ValueError(f"shape mismatch {h.shape} vs {self.shape}")
        L.call("ffb_h2d", self.ptr, h.ctypes.data, self.nbytes)
        L.call("ffb_sync")  # the pageable source must stay alive until the copy has been staged
        return self

    def to_numpy(self):
        out = np.empty(self.shape, dtype=self.dtype, order="F")
        L.call("ffb_d2h", out.ctypes.data, self.ptr, self.nbytes)
        return out

    def copy_from(self, other: "DevArray"):
        if other.nbytes != self.nbytes:
            raise ValueError("size mismatch")
        L.call("ffb_d2d", self.ptr, other.ptr, self.nbytes)
        return self

    def copy(self):
        return DevArray(self.shape, self.dtype).copy_from(self)

    def fill_zero(self):
        L.call("ffb_memset_zero", self.ptr, self.nbytes)
        return self

    def view(self, shape=None, dtype=None, offset_elems=0):
        """Reinterpreting view (no copy); keeps the owner alive."""
        dtype = self.dtype if dtype is None else np.dtype(dtype)
        shape = self.shape if shape is None else shape
        return DevArray(shape, dtype, ptr=self.ptr + offset_elems * self.dtype.itemsize, owner=self)

    def __del__(self):
        try:
            if getattr(self, "_owns", False) and getattr(self, "ptr", None):
                L.load().ffb_free(self.ptr)
                self.ptr = None
        except Exception:
            pass

    def __repr__(self):
        return f"DevArray(shape={self.shape}, dtype={self.dtype})"


def zeros(dev, T, dims):
    """`zeros(dev, T, dims)` (src/utils.jl:79-80)."""
    _require_gpu(dev)
    return DevArray.zeros(T, dims)


def device_array(dev):
    """`device_array(dev)` (src/utils.jl:329-330): returns the constructor that uploads a host array."""
    _require_gpu(dev)
    return DevArray.from_numpy


def devzeros(dev, T, dims, count):
    """`@devzeros dev T dims a b c...` (src/utils.jl:89-94)."""
    return tuple(zeros(dev, T, dims) for _ in range(count))
