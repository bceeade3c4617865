"""Per-kernel GB/s for selected plans under environment overrides (GPU box)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fourierflows_jl_b200 as ff  # noqa: E402
from fourierflows_jl_b200 import _lib as L  # noqa: E402


def run(shape, T, env):
    for k in ("FFB_W_COLS", "FFB_W_ROWS", "FFB_FOURSTEP_MIN"):
        os.environ.pop(k, None)
    os.environ.update({k: str(v) for k, v in env.items()})
    plan = ff.Plan(shape, T, L.FFB_R2C)
    x = ff.DevArray.zeros(T, shape)
    xh = ff.DevArray.zeros(ff.cxtype(T), plan.spectral_shape)
    for _ in range(2):
        plan.mul(xh, x); plan.ldiv(x, xh)
    ff.prof_enable(True)
    for _ in range(5):
        plan.mul(xh, x); plan.ldiv(x, xh)
    rep = ff.prof_report()
    ff.prof_enable(False)
    tot = sum(r["ms"] for r in rep) / 10
    line = " ".join(f"{r['name'][4:]}={r['bytes'] / r['ms'] / 1e6:.0f}({1e3 * r['ms'] / r['launches']:.0f}us)" for r in sorted(rep, key=lambda r: r['name']))
    print(f"{shape} {np.dtype(T).name} {env}: avg {tot:.3f} ms/transform | {line}", flush=True)


for shape, T in (((8192, 8192), np.float64), ((4096, 4096), np.float64)):
    for env in ({}, {"FFB_FOURSTEP_MIN": 0}, {"FFB_W_COLS": 16}, {"FFB_W_COLS": 32}, {"FFB_W_COLS": 64}, {"FFB_W_COLS": 8}):
        run(shape, T, env)
for shape, T in (((1024, 1024, 256), np.float32), ((2048, 2048, 64), np.float32), ((2048, 2048, 64), np.float64), ((1024, 1024, 128), np.float64)):
    for env in ({}, {"FFB_W_COLS": 2}, {"FFB_W_COLS": 4}, {"FFB_W_COLS": 8}, {"FFB_W_COLS": 16}):
        run(shape, T, env)
run((8192, 8192), np.float32, {})
run((8192, 8192), np.float32, {"FFB_FOURSTEP_MIN": 0})
run((8192, 8192), np.float32, {"FFB_W_COLS": 64})
