"""Tiny driver for ncu captures: a few ETDRK4 steps of the C3 problem (fused calcN)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fourierflows_jl_b200 as ff  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
prob = ff.CProblem((n, n), 2 * np.pi, stepper="ETDRK4", dt=1e-3, calcN="vorticity2d", nu=1e-4, fused=1)
rng = np.random.default_rng(0)
prob.set_physical(np.asfortranarray(rng.standard_normal((n, n))))
prob.stepforward(steps)
ff._lib.call("ffb_sync")
print("done", prob.clock)
