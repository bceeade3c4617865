import os, sys, subprocess
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = '''
import os, sys
import numpy as np
sys.path.insert(0, %r)
import fourierflows_jl_b200 as ff
from fourierflows_jl_b200 import _lib as L
import scipy.fft as sfft
cases = [((8192, 8192), np.float64), ((4096, 4096), np.float64), ((2048, 2048, 64), np.float32), ((2048,2048,32), np.float64), ((1024,1024,128), np.float64), ((1024,1024,256), np.float32), ((8192,8192), np.float32), ((64, 4096), np.float64), ((32, 8192), np.float32), ((16, 2048, 4), np.float64)]
for shape, T in cases:
    plan = ff.Plan(shape, T, L.FFB_R2C)
    if np.prod(shape) <= 2**22:
        rng = np.random.default_rng(1)
        a = np.asfortranarray(rng.standard_normal(shape).astype(T))
        ref = sfft.rfftn(a.astype(np.float64), axes=tuple(range(len(shape)-1,-1,-1)))
        da = ff.DevArray.from_numpy(a)
        ah = plan * da
        e1 = np.linalg.norm(ah.to_numpy()-ref)/np.linalg.norm(ref)
        e2 = np.linalg.norm(plan.solve(ah).to_numpy()-a)/np.linalg.norm(a)
        print("  parity", shape, np.dtype(T).name, "fwd %%.2e rt %%.2e" %% (e1, e2), plan.describe(), flush=True)
        continue
    x = ff.DevArray.zeros(T, shape); xh = ff.DevArray.zeros(ff.cxtype(T), plan.spectral_shape)
    for _ in range(2):
        plan.mul(xh, x); plan.ldiv(x, xh)
    ff.prof_enable(True)
    for _ in range(5):
        plan.mul(xh, x); plan.ldiv(x, xh)
    rep = ff.prof_report(); ff.prof_enable(False)
    tot = sum(r["ms"] for r in rep) / 10
    print(shape, np.dtype(T).name, "CLUSTER_MIN", os.environ.get("FFB_CLUSTER_MIN"), "avg %%.3f ms |" %% tot, " ".join("%%s=%%.0f(%%.0fus)" %% (r["name"][4:], r["bytes"]/r["ms"]/1e6, 1e3*r["ms"]/r["launches"]) for r in sorted(rep, key=lambda r: r["name"]) if "rows" not in r["name"]), flush=True)
''' % root
for cm in ("0", "1024"):
    env = dict(os.environ); env["FFB_CLUSTER_MIN"] = cm
    subprocess.run([sys.executable, "-c", code], env=env, timeout=280)
