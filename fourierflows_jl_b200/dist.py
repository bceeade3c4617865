"""Multi-GPU slab decomposition: host-side logic.  One process per GPU; `torch.distributed` (or any other
launcher) only moves the 128-byte NCCL unique id between ranks -- the exchange itself is NCCL inside the library.

Partitioning of a 3-D grid over P ranks:
  physical (nx, ny, nz)      -> rank r holds z in [r*nz/P, (r+1)*nz/P)          shape (nx, ny, nz/P)
  spectral (nx/2+1, ny, nz)  -> rank r holds y (l) in [r*ny/P, (r+1)*ny/P)      shape (nx/2+1, ny/P, nz)
Stages, dealias and filter are pointwise in spectral space (no communication); only the FFT exchanges data.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L
from .array import DevArray, cxtype, ffb_dtype


def slab_range(n: int, nranks: int, rank: int):
    """0-based half-open index range of `rank`'s slab along a dimension of extent n (n % nranks == 0)."""
    if n % nranks != 0:
        raise L.FFBError(L.FFB_EUNSUPPORTED, f"extent {n} is not divisible by the number of ranks {nranks}")
    w = n // nranks
    return rank * w, (rank + 1) * w


def local_alias_range(alias, n: int, nranks: int, rank: int):
    """Intersection of a 1-based inclusive alias range (`grid.lalias`) with this rank's slab, in local 1-based
    indices; None when the slab holds no aliased index."""
    if alias is None:
        return None
    lo0, hi0 = slab_range(n, nranks, rank)
    lo, hi = max(alias[0], lo0 + 1), min(alias[1], hi0)
    if lo > hi:
        return None
    return lo - lo0, hi - lo0


def physical_slab(a: np.ndarray, nranks: int, rank: int) -> np.ndarray:
    lo, hi = slab_range(a.shape[-1], nranks, rank)
    return np.asfortranarray(a[..., lo:hi])


def spectral_slab(ah: np.ndarray, nranks: int, rank: int) -> np.ndarray:
    lo, hi = slab_range(ah.shape[1], nranks, rank)
    return np.asfortranarray(ah[:, lo:hi, ...])


def spectral_slab_2d(ah: np.ndarray, nranks: int, rank: int) -> np.ndarray:
    """Local spectral slab of a 2-D grid: the half spectrum (nx/2 + 1, ny) is cut along kx into blocks of kb = nx/(2P) wavenumbers;
    every rank holds (kb + 1, ny): its block plus one extra column -- the Nyquist wavenumber on the last rank, zero padding elsewhere."""
    nkr = ah.shape[0]
    if (nkr - 1) % nranks != 0:
        raise L.FFBError(L.FFB_EUNSUPPORTED, f"nx/2 = {nkr - 1} is not divisible by the number of ranks {nranks}")
    kb = (nkr - 1) // nranks
    out = np.zeros((kb + 1,) + ah.shape[1:], dtype=ah.dtype, order="F")
    out[:kb] = ah[rank * kb:(rank + 1) * kb]
    if rank == nranks - 1:
        out[kb] = ah[nkr - 1]
    return out


def gather_spectral_2d(slabs, nranks: int) -> np.ndarray:
    """inverse of `spectral_slab_2d`: the full half spectrum from the ranks' slabs (list indexed by rank)"""
    kb = slabs[0].shape[0] - 1
    return np.asfortranarray(np.concatenate([s[:kb] for s in slabs] + [slabs[-1][kb:kb + 1]], axis=0))


def physical_slab_2d(a: np.ndarray, nranks: int, rank: int) -> np.ndarray:
    lo, hi = slab_range(a.shape[1], nranks, rank)
    return np.asfortranarray(a[:, lo:hi])


def local_kx_alias_2d(kralias, nx: int, nranks: int, rank: int):
    """`grid.kralias` (1-based inclusive iL:nkr) in the local indices of a 2-D spectral slab: the block's part of the range plus the
    extra column (always aliased); None when aliased_fraction = 0"""
    if kralias is None:
        return None
    kb = nx // 2 // nranks
    lo = max(kralias[0], rank * kb + 1)
    return ((lo - rank * kb) if lo <= (rank + 1) * kb else kb + 1, kb + 1)


def exchange_bytes_per_rank(shape, nranks: int, itemsize: int) -> int:
    """Bytes each rank sends per 3-D transform: S/P * (P-1)/P."""
    nkr = shape[0] // 2 + 1
    S = nkr * shape[1] * shape[2] * 2 * itemsize
    return S // nranks * (nranks - 1) // nranks


class Dist:
    """NCCL communicator handle (`ffb_dist`)."""

    def __init__(self, rank: int, nranks: int, unique_id: bytes):
        assert len(unique_id) == 128
        self.rank, self.nranks = rank, nranks
        buf = C.create_string_buffer(unique_id, 128)
        h = C.c_void_p()
        L.call("ffb_dist_init", C.byref(h), rank, nranks, buf)
        self._h = h

    @staticmethod
    def unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        L.call("ffb_dist_unique_id", buf)
        return buf.raw

    @classmethod
    def from_torch(cls):
        """Build the communicator inside an initialised `torch.distributed` process group (torchrun)."""
        import torch.distributed as dist
        rank, world = dist.get_rank(), dist.get_world_size()
        obj = [cls.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(obj, src=0)
        return cls(rank, world, obj[0])

    def alltoall(self, send: DevArray, recv: DevArray):
        L.call("ffb_dist_alltoall", self._h, send.ptr, recv.ptr, send.nbytes // self.nranks)

    def close(self):
        if getattr(self, "_h", None):
            L.load().ffb_dist_destroy(self._h)
            self._h = None


EXCHANGE = {"nccl": 0, "peer-store": 1, "copy-engine": 2}


def enable_p2p(plan_handle, dist: "Dist", mode: str = "peer-store", T=None, shape=None, prefer=None):
    """Exchange the IPC handles of a plan's two receive buffers over torch.distributed, hand the mapped peer pointers to
    the plan (`ffb_plan_dist_set_peers`) and select the exchange (`ffb_plan_dist_set_exchange`):
    "peer-store" = the pass before the exchange stores into the peers' buffers, "copy-engine" = chunked cudaMemcpyAsync pushes,
    "auto" = measure and keep the fastest (needs T and the global shape).  Returns the exchange in use."""
    import torch.distributed as td
    b0, b1, nb = C.c_void_p(), C.c_void_p(), C.c_size_t()
    L.call("ffb_plan_dist_recv_buffers", plan_handle, C.byref(b0), C.byref(b1), C.byref(nb))
    h0, h1 = C.create_string_buffer(64), C.create_string_buffer(64)
    o0, o1 = C.c_size_t(), C.c_size_t()
    L.call("ffb_dist_ipc_export", b0, h0, C.byref(o0))
    L.call("ffb_dist_ipc_export", b1, h1, C.byref(o1))
    gathered = [None] * dist.nranks
    td.all_gather_object(gathered, (h0.raw, o0.value, h1.raw, o1.value))
    P = dist.nranks
    p0, p1 = (C.c_void_p * P)(), (C.c_void_p * P)()
    for r in range(P):
        if r == dist.rank:
            p0[r], p1[r] = b0.value, b1.value
        else:
            q0, q1 = C.c_void_p(), C.c_void_p()
            L.call("ffb_dist_ipc_open", C.create_string_buffer(gathered[r][0], 64), gathered[r][1], C.byref(q0))
            L.call("ffb_dist_ipc_open", C.create_string_buffer(gathered[r][2], 64), gathered[r][3], C.byref(q1))
            p0[r], p1[r] = q0.value, q1.value
    L.call("ffb_plan_dist_set_peers", plan_handle, p0, p1)
    # receive buffers are pooled per process and may have served an earlier plan: every rank drains its own stream, then all
    # ranks meet, so no peer can store into a buffer its owner is still reading for the previous plan
    L.call("ffb_sync")
    td.barrier()
    if mode == "auto":
        return autotune_exchange(plan_handle, dist, T, shape, prefer=prefer)
    L.call("ffb_plan_dist_set_exchange", plan_handle, EXCHANGE[mode])
    return mode


def autotune_exchange(plan_handle, dist: "Dist", T, shape, reps: int = 2, prefer=None, slack: float = 1.25):
    """Plan-time measurement (in the spirit of FFTW_MEASURE): time one forward + inverse transform of the plan's size with each
    exchange the plan supports and keep the fastest.  Collective; every rank takes the same decision (max over ranks)."""
    import time
    import torch.distributed as td
    P = dist.nranks
    x = DevArray.zeros(T, (shape[0], shape[1], shape[2] // P))
    xh = DevArray.zeros(cxtype(T), (shape[0] // 2 + 1, shape[1] // P, shape[2]))
    results = {}
    for mode in ("nccl", "copy-engine", "peer-store"):
        try:
            L.call("ffb_plan_dist_set_exchange", plan_handle, EXCHANGE[mode])
        except L.FFBError:
            continue   # not supported for this size
        for timed in (False, True):
            L.call("ffb_sync")
            td.barrier()
            t0 = time.perf_counter()
            for _ in range(reps):
                L.call("ffb_fft_forward", plan_handle, x.ptr, xh.ptr)
                L.call("ffb_fft_inverse", plan_handle, xh.ptr, x.ptr)
            L.call("ffb_sync")
            dt = time.perf_counter() - t0
        all_dt = [None] * P
        td.all_gather_object(all_dt, dt)
        results[mode] = max(all_dt) / reps
    best = min(results, key=results.get)
    # `prefer`: an exchange the caller gains more from than this plain-transform measurement shows (fused problems: only the
    # peer-store / NCCL exchanges fold calcN!'s elementwise work into the passes and skip the aliased modes) wins unless it is
    # more than `slack` times slower
    if prefer in results and results[prefer] <= slack * results[best]:
        best = prefer
    L.call("ffb_plan_dist_set_exchange", plan_handle, EXCHANGE[best])
    td.barrier()
    AUTOTUNE_LOG.append({"shape": tuple(shape), "dtype": np.dtype(T).name, "ranks": P, "ms": {k: round(1e3 * v, 3) for k, v in results.items()}, "chosen": best, "preferred": prefer})
    return best


AUTOTUNE_LOG = []


class DistPlan:
    """Slab-decomposed `rfftplan` of a 2-D or 3-D grid: `mul(out, a)` / `ldiv(out, ah)` on the local slabs."""

    def __init__(self, shape, T, dist: Dist, nchunks: int = 0):
        self.shape = tuple(int(s) for s in shape)
        self.T = np.dtype(T)
        self.dist = dist
        nd = len(self.shape)
        n = (C.c_int64 * 3)(*self.shape, *([1] * (3 - nd)))
        h = C.c_void_p()
        L.call("ffb_plan_create_dist", C.byref(h), nd, n, ffb_dtype(self.T), dist._h, nchunks)
        self._h = h
        P = dist.nranks
        if nd == 2:   # physical y-slabs <-> spectral kx blocks (+ the Nyquist / padding column)
            self.physical_shape = (self.shape[0], self.shape[1] // P)
            self.spectral_shape = (self.shape[0] // 2 // P + 1, self.shape[1])
        else:
            self.physical_shape = (self.shape[0], self.shape[1], self.shape[2] // P)
            self.spectral_shape = (self.shape[0] // 2 + 1, self.shape[1] // P, self.shape[2])

    def describe(self):
        buf = C.create_string_buffer(512)
        L.call("ffb_plan_describe", self._h, buf, 512)
        return buf.value.decode()

    def enable_p2p(self, mode: str = "peer-store"):
        """Map every peer's receive buffers (CUDA IPC over NVLink) and exchange through them: "peer-store" (the pass before
        the exchange stores straight into them) or "copy-engine" (chunked cudaMemcpyAsync pushes beside the next chunk's pass).
        Collective call: every rank of the torch.distributed group must make it."""
        self.exchange = enable_p2p(self._h, self.dist, mode, self.T, self.shape)
        return self

    def mul(self, out: DevArray, a: DevArray):
        L.call("ffb_fft_forward", self._h, a.ptr, out.ptr)
        return out

    def ldiv(self, out: DevArray, ah: DevArray):
        L.call("ffb_fft_inverse", self._h, ah.ptr, out.ptr)
        return out

    def __mul__(self, a):
        return self.mul(DevArray(self.spectral_shape, cxtype(self.T)), a)

    def solve(self, ah):
        return self.ldiv(DevArray(self.physical_shape, self.T), ah)

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                L.load().ffb_plan_destroy(self._h)
                self._h = None
        except Exception:
            pass
