import os, sys, subprocess
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
code = '''
import os, sys
import numpy as np
sys.path.insert(0, %r)
import fourierflows_jl_b200 as ff
from fourierflows_jl_b200 import _lib as L
for shape, T in (((8192, 8192), np.float64), ((4096, 4096), np.float64), ((2048, 2048, 64), np.float32), ((1024,1024,128), np.float64), ((8192,8192), np.float32)):
    plan = ff.Plan(shape, T, L.FFB_R2C)
    x = ff.DevArray.zeros(T, shape); xh = ff.DevArray.zeros(ff.cxtype(T), plan.spectral_shape)
    for _ in range(2):
        plan.mul(xh, x); plan.ldiv(x, xh)
    ff.prof_enable(True)
    for _ in range(5):
        plan.mul(xh, x); plan.ldiv(x, xh)
    rep = ff.prof_report(); ff.prof_enable(False)
    tot = sum(r["ms"] for r in rep) / 10
    print(shape, np.dtype(T).name, os.environ.get("FFB_PF_AHEAD"), "avg %%.3f ms |" %% tot, " ".join("%%s=%%.0f" %% (r["name"][4:], r["bytes"]/r["ms"]/1e6) for r in sorted(rep, key=lambda r: r["name"]) if "rows" in r["name"]), flush=True)
''' % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for pf in ("0", "148", "296", "592", "1184", "2368"):
    env = dict(os.environ); env["FFB_PF_AHEAD"] = pf
    subprocess.run([sys.executable, "-c", code], env=env)
