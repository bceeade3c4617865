"""CPU oracle for the FourierFlows.jl pseudospectral time-stepping hot path.

TEST INFRASTRUCTURE ONLY.  This package is a NumPy/SciPy restatement of the reference's
algorithm (FourierFlows.jl v0.10.7, `/root/reference/src/{domains,timesteppers,problem,diffusion}.jl`).
It may be imported only by `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` /
`--impl reference` legs of `bench.py` -- always as the checker or the timed CPU baseline,
never as part of the product path (`fourierflows_jl_b200/` must not import it).

Parity pinning: the reference is pure Julia and cannot run in this image (no `julia`, no FFTW);
its FFT arithmetic lives in third-party FFTW.jl / cuFFT (no pinned Manifest).  The reference's
tests hold no golden vectors, only closed-form known-answer tests; `tests/test_oracle_*.py`
re-states every one of those against this oracle (FFT single-mode spectra, Gaussian diffusion for
all ten steppers, alias ranges, dealias box, filter plateau/tail, DomainError).  Bit-level parity
with the Julia implementation is therefore UNPINNED; behavioural parity is pinned by those tests.
"""
from .fforacle import *  # noqa: F401,F403
