"""Streaming strided-pass kernel (fft_stream.cuh): parity against scipy.fft on ragged shapes, then per-kernel timing
against the plain register kernel.  Diagnostic tool for the GPU box, not part of the product."""
import os, sys
import numpy as np
import scipy.fft as sfft
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fourierflows_jl_b200 as ff
from fourierflows_jl_b200 import _lib as L


def relerr(a, b):
    return float(np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-300))


def check(shape, T):
    rng = np.random.default_rng(3)
    x = np.asfortranarray(rng.standard_normal(shape).astype(T))
    ref = sfft.rfftn(x.astype(np.float64), axes=tuple(range(len(shape) - 1, -1, -1)))
    plan = ff.Plan(shape, T, L.FFB_R2C)
    dx = ff.DevArray.from_numpy(x)
    xh = plan * dx
    e1 = relerr(xh.to_numpy(), ref)
    e2 = relerr(plan.solve(xh).to_numpy(), x)
    return e1, e2


def prof(shape, T, reps=4):
    plan = ff.Plan(shape, T, L.FFB_R2C)
    x = ff.DevArray.zeros(T, shape)
    xh = ff.DevArray.zeros(ff.cxtype(T), plan.spectral_shape)
    out = {}
    for name, fn in (("fwd", lambda: plan.mul(xh, x)), ("inv", lambda: plan.ldiv(x, xh))):
        for _ in range(2):
            fn()
        L.call("ffb_sync")
        ff.prof_enable(True)
        for _ in range(reps):
            fn()
        rep = ff.prof_report()
        ff.prof_enable(False)
        out[name] = (round(sum(r["ms"] for r in rep) / reps, 3), [(r["name"].replace("fft_", ""), round(r["ms"] / r["launches"], 3), round(r["bytes"] / r["ms"] / 1e6 / 6550.1, 2)) for r in rep])
    return out


def main():
    tol = {np.float64: 1e-13, np.float32: 2e-6}
    bad = 0
    quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
    for smin in (() if quick else ("256",)):
        os.environ["FFB_STREAM_MIN"] = smin
        for T in (np.float32, np.float64):
            for shape in ((64, 256), (34, 512), (30, 1024), (16, 2048), (6, 2048), (258, 512), (8, 256, 512), (36, 1024, 4), (4, 4, 2048), (62, 512, 256)):
                for wenv in (None, "4", "32"):
                    if wenv is None: os.environ.pop("FFB_W_STREAM", None)
                    else: os.environ["FFB_W_STREAM"] = wenv
                    for split in (None, "1"):
                        if split is None: os.environ.pop("FFB_STREAM_SPLIT", None)
                        else: os.environ["FFB_STREAM_SPLIT"] = split
                        try:
                            e1, e2 = check(shape, T)
                            ok = e1 < tol[T] * 10 and e2 < tol[T] * 10
                            bad += not ok
                            if not ok or (wenv is None and split is None):
                                print(f"parity {np.dtype(T).name} {shape} W={wenv} split={split}: fwd {e1:.2e} rt {e2:.2e} {'OK' if ok else 'FAIL'}", flush=True)
                        except Exception as ex:  # noqa: BLE001
                            bad += 1
                            print(f"parity {np.dtype(T).name} {shape} W={wenv} split={split} EXC {ex}", flush=True)
    os.environ.pop("FFB_W_STREAM", None); os.environ.pop("FFB_STREAM_SPLIT", None)
    print("FAILURES:", bad, flush=True)
    cases = [((2048, 2048, 256), np.float32), ((1024, 1024, 1024), np.float32), ((512, 512, 512), np.float32), ((512, 512, 512), np.float64),
             ((2048, 2048), np.float64), ((1024, 1024, 256), np.float64), ((256, 256, 256), np.float32), ((4096, 2048), np.float32)]
    for shape, T in cases:
        for smin, wenv, split in ((("0", None, None),) if quick else (("0", None, None), ("256", None, None), ("256", "8", None), ("256", "16", None), ("256", "32", None), ("256", None, "1"))):
            os.environ["FFB_STREAM_MIN"] = smin
            if wenv is None: os.environ.pop("FFB_W_STREAM", None)
            else: os.environ["FFB_W_STREAM"] = wenv
            if split is None: os.environ.pop("FFB_STREAM_SPLIT", None)
            else: os.environ["FFB_STREAM_SPLIT"] = split
            try:
                r = prof(shape, T)
                print(f"time {shape} {np.dtype(T).name} stream_min={smin} W={wenv} split={split}: fwd {r['fwd'][0]} inv {r['inv'][0]}  fwd-kernels {r['fwd'][1]}", flush=True)
            except Exception as ex:  # noqa: BLE001
                print(f"time {shape} stream_min={smin} W={wenv} EXC {ex}", flush=True)


if __name__ == "__main__":
    main()
