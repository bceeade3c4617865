"""Asynchronous output path.  The reference's `saveoutput(out)` (src/output.jl:61-79) downloads every field with a
blocking `Array(data)` inside the step loop; here a snapshot ring (`ffb_snapshot_*`) stages the field on the device in stream order
and moves it to pinned host memory on a dedicated copy stream while stepping continues.  `Output` mirrors the reference type
(`Output(prob, filename, fields...)`, `saveoutput(out)`); JLD2 is Julia-only, so snapshots are written as NumPy `.npz` groups keyed
like the reference's `snapshots/<field>/<step>`.  Decomposed fields: every rank snapshots its slab, `gather_to_rank0` assembles them."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _lib as L
from .array import DevArray


class AsyncSnapshot:
    """Ring of (device staging, pinned host) buffer pairs for arrays of at most `nbytes` bytes."""

    def __init__(self, nbytes: int, nbuf: int = 2):
        h = C.c_void_p()
        L.call("ffb_snapshot_create", C.byref(h), int(nbytes), int(nbuf))
        self._h, self.nbytes, self.nbuf = h, int(nbytes), int(nbuf)

    def begin(self, a: DevArray) -> int:
        """enqueue the snapshot of `a` (non-blocking); returns the slot"""
        slot = C.c_int(-1)
        L.call("ffb_snapshot_begin", self._h, a.ptr, a.nbytes, C.byref(slot))
        return slot.value

    def ready(self, slot: int) -> bool:
        r = C.c_int(0)
        L.call("ffb_snapshot_ready", self._h, slot, C.byref(r))
        return bool(r.value)

    def wait(self, slot: int, shape, dtype) -> np.ndarray:
        """block until the slot has landed; returns a Fortran-ordered view of the pinned buffer (valid until `release`)"""
        p = C.c_void_p()
        L.call("ffb_snapshot_wait", self._h, slot, C.byref(p))
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        buf = (C.c_char * n).from_address(p.value)
        return np.frombuffer(buf, dtype=dtype).reshape(shape, order="F")

    def release(self, slot: int) -> None:
        L.call("ffb_snapshot_release", self._h, slot)

    def close(self):
        if getattr(self, "_h", None):
            L.load().ffb_snapshot_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass


def gather_to_rank0(local: np.ndarray, assemble):
    """Gather every rank's host slab on rank 0 (torch.distributed, any backend) and assemble them with `assemble(list_of_slabs)`;
    returns None on the other ranks.  Runs on the writer's side of the snapshot ring, never inside the step loop."""
    import torch.distributed as td
    world, rank = td.get_world_size(), td.get_rank()
    parts = [None] * world if rank == 0 else None
    td.gather_object(np.ascontiguousarray(local), parts, dst=0)
    return assemble(parts) if rank == 0 else None


class Output:
    """`Output(prob, filename, fields)` (src/output.jl:12-50): `fields` maps a name to `prob -> DevArray`."""

    def __init__(self, prob, filename: str, fields: dict, nbuf: int = 2):
        self.prob, self.path, self.fields = prob, filename, dict(fields)
        self._pending = []    # (step, t, name, slot, shape, dtype)
        self._snap = None
        self._nbuf = nbuf
        self._index = 0

    def _ring(self, nbytes):
        if self._snap is None or self._snap.nbytes < nbytes:
            self.flush()
            self._snap = AsyncSnapshot(nbytes, max(self._nbuf, len(self.fields)))
        return self._snap

    def saveoutput(self):
        """`saveoutput(out)` (src/output.jl:61-71): enqueue one snapshot of every field; files are written by `flush` (or by the next
        `saveoutput` that needs the ring slots), so the step loop never waits for PCIe or the disk."""
        clock = self.prob.clock
        step, t = (clock.step, float(clock.t)) if hasattr(clock, "step") else (clock[1], clock[0])
        arrays = {k: f(self.prob) for k, f in self.fields.items()}
        ring = self._ring(max(a.nbytes for a in arrays.values()))
        if len(self._pending) + len(arrays) > ring.nbuf:
            self.flush()
        for name, a in arrays.items():
            self._pending.append((step, t, name, ring.begin(a), a.shape, a.dtype))

    def flush(self):
        """write every pending snapshot (blocks until its copy has landed)"""
        if not self._pending:
            return
        groups = {}
        for step, t, name, slot, shape, dtype in self._pending:
            groups.setdefault(step, {"t": t})[name] = np.array(self._snap.wait(slot, shape, dtype), order="F")
            self._snap.release(slot)
        self._pending = []
        base, _ = os.path.splitext(self.path)
        for step, g in groups.items():
            np.savez(f"{base}_snapshot_{step}.npz", **{f"snapshots/{k}/{step}": v for k, v in g.items()})
            self._index += 1

    def close(self):
        self.flush()
        if self._snap is not None:
            self._snap.close()
            self._snap = None


def saveoutput(out: Output):
    out.saveoutput()
