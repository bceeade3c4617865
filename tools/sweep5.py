import os, sys, subprocess
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = '''
import os, sys
import numpy as np
sys.path.insert(0, %r)
import fourierflows_jl_b200 as ff
from fourierflows_jl_b200 import _lib as L
for shape, T in (((8192, 8192), np.float64), ((4096, 4096), np.float64), ((512, 512, 512), np.float64), ((1024, 1024, 512), np.float32), ((8192, 8192), np.float32)):
    plan = ff.Plan(shape, T, L.FFB_R2C)
    x = ff.DevArray.zeros(T, shape); xh = ff.DevArray.zeros(ff.cxtype(T), plan.spectral_shape)
    for _ in range(2):
        plan.mul(xh, x); plan.ldiv(x, xh)
    ff.prof_enable(True)
    for _ in range(5):
        plan.mul(xh, x); plan.ldiv(x, xh)
    rep = ff.prof_report(); ff.prof_enable(False)
    tot = sum(r["ms"] for r in rep) / 10
    print(shape, np.dtype(T).name, {k: v for k, v in os.environ.items() if k.startswith("FFB_")}, "avg %%.3f ms |" %% tot, " ".join("%%s=%%.0f(%%.0fus)" %% (r["name"][4:], r["bytes"]/r["ms"]/1e6, 1e3*r["ms"]/r["launches"]) for r in sorted(rep, key=lambda r: r["name"])), flush=True)
''' % root
for extra in ({}, {"FFB_SNAKE": "1"}):
    env = dict(os.environ); env.update(extra)
    subprocess.run([sys.executable, "-c", code], env=env, timeout=280)
