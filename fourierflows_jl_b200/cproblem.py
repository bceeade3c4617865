"""C-driven problems (`ffb_problem_*`): the whole `stepforward!` loop, including the built-in `calcN!` of the benchmark
equations, runs inside libfourierflows_b200.so with no host language in the loop (SURVEY 8b, B3 form ii)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L
from .array import DevArray, cxtype, ffb_dtype

_STEPPER = {"ForwardEuler": L.FFB_FORWARD_EULER, "RK4": L.FFB_RK4, "LSRK54": L.FFB_LSRK54, "ETDRK4": L.FFB_ETDRK4, "AB3": L.FFB_AB3}
_CALCN = {"zero": L.FFB_CALCN_ZERO, "diffusion": L.FFB_CALCN_DIFFUSION, "vorticity2d": L.FFB_CALCN_VORTICITY2D,
          "burgers3d": L.FFB_CALCN_BURGERS3D, "callback": L.FFB_CALCN_CALLBACK}


class CProblem:
    def __init__(self, n, Lext, stepper="ETDRK4", dt=1e-3, calcN="vorticity2d", nu=0.0, T=np.float64, aliased_fraction=1 / 3,
                 coef_dtype=None, kappa: DevArray = None, scalar_zero_L=False, callback=None, filter_kwargs=None, fused=0, dist=None):
        n = tuple(int(v) for v in (n if isinstance(n, (tuple, list)) else (n,)))
        Lext = tuple(float(v) for v in (Lext if isinstance(Lext, (tuple, list)) else (Lext,) * len(n)))
        self.T = np.dtype(T)
        self.n = n
        self.nkr = n[0] // 2 + 1
        self.dist = dist
        P = dist.nranks if dist is not None else 1
        if dist is not None and len(n) == 2:
            self.spectral_shape = (n[0] // 2 // P + 1, n[1])      # kx block + the Nyquist / padding column (dist.spectral_slab_2d)
            self.physical_shape = (n[0], n[1] // P)               # y-slab
        elif dist is not None:
            self.spectral_shape = (self.nkr, n[1] // P, n[2])     # y-slab
            self.physical_shape = (n[0], n[1], n[2] // P)         # z-slab
        else:
            self.spectral_shape = (self.nkr,) + n[1:]
            self.physical_shape = n
        cfg = L.ffb_problem_config()
        cfg.ndim = len(n)
        cfg.n = (C.c_int64 * 3)(*n, *([1] * (3 - len(n))))
        cfg.L = (C.c_double * 3)(*Lext, *([1.0] * (3 - len(n))))
        cfg.dtype = ffb_dtype(self.T)
        cfg.aliased_fraction = float(aliased_fraction)
        filtered = stepper.startswith("Filtered")
        cfg.stepper = _STEPPER[stepper[len("Filtered"):] if filtered else stepper]
        cfg.filtered = 1 if filtered else 0
        fk = filter_kwargs or {}
        # each keyword has its own "use the reference default" sentinel in ffb_problem_config (<= 0; innerK: < 0)
        cfg.filter_order, cfg.filter_innerK = float(fk.get("order", 0.0)), float(fk.get("innerK", -1.0))
        cfg.filter_outerK, cfg.filter_tol = float(fk.get("outerK", 0.0)), float(fk.get("tol", 0.0))
        cfg.dt = float(dt)
        cfg.calcN = _CALCN[calcN]
        self._cb = L.CALCN_FN(callback) if callback is not None else L.CALCN_FN()
        cfg.callback = self._cb
        cfg.user = None
        cfg.nu = float(nu)
        cfg.scalar_zero_L = 1 if scalar_zero_L else 0
        self._kappa = kappa
        cfg.kappa = kappa.ptr if kappa is not None else None
        cfg.coef_dtype = ffb_dtype(np.float64 if coef_dtype is None else coef_dtype)
        cfg.fused = int(fused)
        self._fused = bool(fused)
        cfg.dist = dist._h if dist is not None else None
        h = C.c_void_p()
        L.call("ffb_problem_create", C.byref(h), C.byref(cfg))
        self._h = h
        p, cnt = C.c_void_p(), C.c_int64()
        L.call("ffb_problem_sol", self._h, C.byref(p), C.byref(cnt))
        # non-owning view WITHOUT a back-reference: a CProblem <-> sol cycle would leave multi-GB problems to the cyclic GC;
        # the view is valid while the problem lives, close() invalidates it
        self.sol = DevArray(self.spectral_shape, cxtype(self.T), ptr=p.value)

    @property
    def clock(self):
        t, step, dt = C.c_double(), C.c_int64(), C.c_double()
        L.call("ffb_problem_clock", self._h, C.byref(t), C.byref(step), C.byref(dt))
        return t.value, step.value, dt.value

    def enable_p2p(self, mode: str = "peer-store"):
        """slab-decomposed problems: exchange through peer memory, see `DistPlan.enable_p2p` (collective call on every rank)"""
        from .dist import enable_p2p
        ph = C.c_void_p()
        L.call("ffb_problem_plan", self._h, C.byref(ph))
        self.exchange = enable_p2p(ph, self.dist, mode, self.T, self.n, prefer="peer-store" if self._fused else None)
        return self

    def device_bytes(self):
        b = C.c_size_t()
        L.call("ffb_problem_bytes", self._h, C.byref(b))
        return b.value

    def set_physical(self, field):
        f = np.asfortranarray(field, dtype=self.T)
        assert f.shape == self.physical_shape
        L.call("ffb_problem_set_physical", self._h, f.ctypes.data)
        L.call("ffb_sync")

    def get_physical(self):
        out = np.empty(self.physical_shape, dtype=self.T, order="F")
        L.call("ffb_problem_get_physical", self._h, out.ctypes.data)
        return out

    def stepforward(self, nsteps=1):
        L.call("ffb_step", self._h, int(nsteps))

    def step_until(self, stop_time):
        L.call("ffb_step_until", self._h, float(stop_time))

    def pipeline(self, depth=3):
        """host-buffer pipeline of this problem (`ffb_pipeline_*`): see `HostPipeline`"""
        return HostPipeline(self, depth)

    def close(self):
        if getattr(self, "_h", None):
            L.load().ffb_problem_destroy(self._h)
            self._h = None
            if getattr(self, "sol", None) is not None:
                self.sol.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class PinnedBuffer:
    """Page-locked host array (`ffb_host_alloc_pinned`) viewed as a Fortran-ordered NumPy array."""

    def __init__(self, shape, dtype):
        self.shape, self.dtype = tuple(shape), np.dtype(dtype)
        self.nbytes = int(np.prod(self.shape)) * self.dtype.itemsize
        p = C.c_void_p()
        L.call("ffb_host_alloc_pinned", C.byref(p), self.nbytes)
        self.ptr = p.value
        self.array = np.frombuffer((C.c_char * self.nbytes).from_address(self.ptr), dtype=self.dtype).reshape(self.shape, order="F")

    def close(self):
        if getattr(self, "ptr", None):
            self.array = None
            L.load().ffb_host_free_pinned(C.c_void_p(self.ptr))
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class HostPipeline:
    """Independent spectral states in pinned host memory, stepped on the GPU with the copies of neighbouring submissions
    overlapping the steps of the current one (no reference counterpart; the blocking form is `sol .= ...; stepforward!; Array(sol)`).

        pipe = prob.pipeline(depth=3)
        t = pipe.submit(inp, out, nsteps=1)     # PinnedBuffer in / out; returns at once unless `depth` submissions are in flight
        pipe.wait(t)                            # out.array now holds the stepped state
    """

    def __init__(self, prob: CProblem, depth=3):
        h = C.c_void_p()
        L.call("ffb_pipeline_create", C.byref(h), prob._h, int(depth))
        self._h, self.prob, self.depth = h, prob, int(depth)

    def submit(self, inp: PinnedBuffer, out: PinnedBuffer, nsteps=1) -> int:
        if inp.nbytes != self.prob.sol.nbytes or out.nbytes != self.prob.sol.nbytes:
            raise ValueError("host buffers must have the size of prob.sol")
        t = C.c_int(-1)
        L.call("ffb_pipeline_submit", self._h, C.c_void_p(inp.ptr), C.c_void_p(out.ptr), int(nsteps), C.byref(t))
        return t.value

    def wait(self, ticket: int) -> None:
        L.call("ffb_pipeline_wait", self._h, int(ticket))

    def wait_all(self) -> None:
        for t in range(self.depth):
            self.wait(t)

    def close(self):
        if getattr(self, "_h", None):
            L.load().ffb_pipeline_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
