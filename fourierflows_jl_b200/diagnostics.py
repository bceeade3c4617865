"""`Diagnostic` bookkeeping -- host mirror of /root/reference/src/diagnostics.jl (host-side only; the `calc`
function typically ends in a device reduction such as `parsevalsum2`, the one forced sync of a production loop)."""
from __future__ import annotations


class Diagnostic:
    """`Diagnostic(calc, prob; freq=1, nsteps=100, ndata=ceil(Int, (nsteps+1)/freq))` (src/diagnostics.jl:36-49)."""

    def __init__(self, calc, prob, freq=1, nsteps=100, ndata=None):
        import math
        ndata = math.ceil((nsteps + 1) / freq) if ndata is None else ndata
        self.calc, self.prob, self.freq = calc, prob, freq
        first = calc(prob)
        self.data = [first] + [None] * (ndata - 1)
        self.t = [prob.clock.t] + [None] * (ndata - 1)
        self.steps = [prob.clock.step] + [None] * (ndata - 1)
        self.i = 1  # number of stored entries (Julia's 1-based `diag.i`)

    def extend(self, n=None):
        """`extend!(diag, n)` (:56-69)."""
        n = len(self.t) if n is None else n
        self.data += [None] * n
        self.t += [None] * n
        self.steps += [None] * n

    def update(self):
        """`update!(diag)` (:78-89)."""
        if self.i >= len(self.data):
            self.extend()
        self.data[self.i] = self.calc(self.prob)
        self.t[self.i] = self.prob.clock.t
        self.steps[self.i] = self.prob.clock.step
        self.i += 1

    def increment(self):
        """`increment!(diag)` (:92-96): record when `clock.step % freq == 0`."""
        if self.prob.clock.step % self.freq == 0:
            self.update()

    def __getitem__(self, idx):  # :109-116
        return self.data[: self.i][idx]

    def __len__(self):
        return self.i


def increment(diags):
    """`increment!(diags)` for one Diagnostic or a list (:98-104)."""
    if isinstance(diags, (list, tuple)):
        for d in diags:
            d.increment()
    else:
        diags.increment()
