"""fourierflows_jl_b200 -- B200-native drop-in for the pseudospectral time-stepping hot path of FourierFlows.jl.

Host-side mirror of the reference API (`OneDGrid/TwoDGrid/ThreeDGrid(dev; nx, Lx, ...)`, `Problem(eqn, stepper, dt, grid)`,
`stepforward!`/`step_until!`, `dealias!`, `Diagnostic`, user `calcN!`/`L`) over the C ABI of `libfourierflows_b200.so`
(include/fourierflows_b200.h).  Python + ctypes stands in for Julia + `ccall` (no Julia toolchain in this image; the
Julia wrapper is `julia/FourierFlowsB200.jl`).  Julia's `f!` is spelled `f` here.

`fourierflows.jl_b200` at the repo root is a symlink to this directory (the name the task layout uses is not importable).
"""
from . import _lib
from ._lib import DomainError, FFBError, have_device, launch_count, prof_enable, prof_report
from .array import CPU, GPU, DevArray, cxtype, device_array, devzeros, fltype, zeros
from .domains import (OneDGrid, Plan, ThreeDGrid, TwoDGrid, dealias, getaliasedwavenumbers, gridpoints, ldiv_,
                      makefilter, mul_)
from .problem import Clock, EmptyParams, EmptyVars, Equation, Problem
from .timesteppers import (STEPPERS, TimeStepper, getetdcoeffs_and_expLs, isexplicit, step_until, stepforward)
from .diagnostics import Diagnostic, increment
from .utils import axpby, fft, ifft, irfft, jacobian, jacobianh, mul, mul_real, parsevalsum, parsevalsum2, rfft, spectral_mul
from .output import AsyncSnapshot, Output, gather_to_rank0, saveoutput
from . import diffusion as Diffusion
from .equations import Burgers3D, TwoDNavierStokes
from .cproblem import CProblem, HostPipeline, PinnedBuffer
from .dist import (Dist, DistPlan, exchange_bytes_per_rank, gather_spectral_2d, local_alias_range, local_kx_alias_2d, physical_slab,
                   physical_slab_2d, slab_range, spectral_slab, spectral_slab_2d)

__all__ = [n for n in dir() if not n.startswith("_")]
