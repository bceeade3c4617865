"""Tuning sweep (GPU box): per-kernel GB/s of the FFT passes as a function of the lines-per-CTA override."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fourierflows_jl_b200 as ff  # noqa: E402
from fourierflows_jl_b200 import _lib as L  # noqa: E402

cases = [((8192, 8192), np.float64), ((4096, 4096), np.float64), ((2048, 2048, 128), np.float32), ((1024, 1024, 256), np.float64),
         ((512, 512, 512), np.float64), ((8192, 8192), np.float32)]
for shape, T in cases:
    plan = ff.Plan(shape, T, L.FFB_R2C)
    x = ff.DevArray.zeros(T, shape)
    xh = ff.DevArray.zeros(ff.cxtype(T), plan.spectral_shape)
    for var in ("FFB_W_COLS", "FFB_W_ROWS"):
        for w in (0, 1, 2, 4, 8, 16, 32):
            os.environ.pop("FFB_W_COLS", None)
            os.environ.pop("FFB_W_ROWS", None)
            if w:
                os.environ[var] = str(w)
            for _ in range(2):
                plan.mul(xh, x); plan.ldiv(x, xh)
            ff.prof_enable(True)
            for _ in range(5):
                plan.mul(xh, x); plan.ldiv(x, xh)
            rep = ff.prof_report()
            ff.prof_enable(False)
            want = "cols" if var == "FFB_W_COLS" else "rows"
            line = " ".join(f"{r['name'][4:]}={r['bytes'] / r['ms'] / 1e6:.0f}" for r in sorted(rep, key=lambda r: r['name']) if want in r["name"])
            print(f"{shape} {np.dtype(T).name} {var}={w or 'auto'}: {line}", flush=True)
    del plan, x, xh
