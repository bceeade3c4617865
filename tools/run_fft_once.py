"""Tiny driver for ncu captures: a few forward / inverse 2-D (or 3-D) r2c transforms."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fourierflows_jl_b200 as ff  # noqa: E402
from fourierflows_jl_b200 import _lib as L  # noqa: E402

shape = tuple(int(v) for v in sys.argv[1].split("x")) if len(sys.argv) > 1 else (8192, 8192)
T = np.float32 if (len(sys.argv) > 2 and sys.argv[2] == "f32") else np.float64
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
plan = ff.Plan(shape, T, L.FFB_R2C)
x = ff.DevArray.zeros(T, shape)
xh = ff.DevArray.zeros(ff.cxtype(T), plan.spectral_shape)
for _ in range(reps):
    plan.mul(xh, x)
    plan.ldiv(x, xh)
L.call("ffb_sync")
print("done", plan.describe())
