"""Diagnostic sweep run on the GPU box: FFT parity per size / mode / dtype against scipy.fft, plus timing of the
2-D / 3-D transforms against the HBM roofline.  Prints one line per case (never raises) so a single gpurun call
shows every failing configuration.  Not part of the product or the test-suite."""
import json
import os
import sys
import time

import numpy as np
import scipy.fft as sfft

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fourierflows_jl_b200 as ff  # noqa: E402
from fourierflows_jl_b200 import _lib as L  # noqa: E402
import ctypes as C  # noqa: E402


def relerr(a, b):
    return float(np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-300))


def check_r2c(shape, T, flags=0, nbatch=1):
    rng = np.random.default_rng(1)
    full = shape + ((nbatch,) if nbatch > 1 else ())
    x = np.asfortranarray(rng.standard_normal(full).astype(T))
    axes = tuple(range(len(shape) - 1, -1, -1))
    ref = sfft.rfftn(x.astype(np.float64), axes=axes)
    plan = ff.Plan(shape, T, L.FFB_R2C, nbatch=nbatch, flags=flags)
    dx = ff.DevArray.from_numpy(x)
    dxh = plan * dx
    got = dxh.to_numpy()
    e1 = relerr(got, ref)
    back = plan.solve(dxh).to_numpy()
    e2 = relerr(back, x)
    e3 = relerr(dx.to_numpy(), x)  # input preserved
    return e1, e2, e3, plan.describe()


def check_c2c(shape, T, flags=0):
    rng = np.random.default_rng(2)
    cT = np.complex64 if T == np.float32 else np.complex128
    x = np.asfortranarray((rng.standard_normal(shape) + 1j * rng.standard_normal(shape)).astype(cT))
    ref = sfft.fftn(x.astype(np.complex128))
    plan = ff.Plan(shape, T, L.FFB_C2C, flags=flags)
    dx = ff.DevArray.from_numpy(x)
    dxh = plan * dx
    e1 = relerr(dxh.to_numpy(), ref)
    e2 = relerr(plan.solve(dxh).to_numpy(), x)
    return e1, e2, plan.describe()


def time_plan(shape, T, reps=10):
    plan = ff.Plan(shape, T, L.FFB_R2C)
    x = ff.DevArray.zeros(T, shape)
    xh = ff.DevArray.zeros(ff.cxtype(T), plan.spectral_shape)
    st = C.c_void_p()
    L.call("ffb_get_stream", C.byref(st))
    import torch
    stream = torch.cuda.ExternalStream(st.value)
    res = {}
    with torch.cuda.stream(stream):
        for name, fn in (("fwd", lambda: plan.mul(xh, x)), ("inv", lambda: plan.ldiv(x, xh))):
            for _ in range(3):
                fn()
            L.call("ffb_sync")
            best = 1e9
            for _ in range(reps):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                fn()
                e1.record(stream)
                e1.synchronize()
                best = min(best, e0.elapsed_time(e1))
            res[name] = best
    es = np.dtype(T).itemsize
    P = np.prod(shape) * es
    S = np.prod(plan.spectral_shape) * es * 2
    alg = P + (2 * len(shape) - 1) * S
    return res, alg


def main():
    print("device:", ff.have_device())
    tol = {np.float64: 1e-13, np.float32: 2e-6}
    bad = 0
    for T in (np.float64, np.float32):
        for nx in (4, 6, 8, 10, 16, 30, 32, 34, 64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384, 32768):
            try:
                e1, e2, e3, d = check_r2c((nx,), T)
                ok = e1 < tol[T] * 10 and e2 < tol[T] * 10 and e3 == 0
                bad += not ok
                print(f"r2c1d {np.dtype(T).name} nx={nx:6d} fwd {e1:.2e} rt {e2:.2e} in-preserved {e3 == 0} {'OK' if ok else 'FAIL'} [{d}]")
            except Exception as ex:  # noqa: BLE001
                bad += 1
                print(f"r2c1d {np.dtype(T).name} nx={nx} EXC {type(ex).__name__}: {ex}")
        for nx in (2, 4, 8, 16, 30, 32, 64, 256, 1024, 4096, 8192, 16384):
            try:
                e1, e2, d = check_c2c((nx,), T)
                ok = e1 < tol[T] * 10 and e2 < tol[T] * 10
                bad += not ok
                print(f"c2c1d {np.dtype(T).name} nx={nx:6d} fwd {e1:.2e} rt {e2:.2e} {'OK' if ok else 'FAIL'} [{d}]")
            except Exception as ex:  # noqa: BLE001
                bad += 1
                print(f"c2c1d {np.dtype(T).name} nx={nx} EXC {type(ex).__name__}: {ex}")
        for shape in ((6, 8), (32, 64), (64, 32), (256, 128), (30, 16), (16, 30), (1024, 512), (128, 2048), (8, 4096), (16, 8192), (4096, 16)):
            for flags in (0, 1):
                try:
                    e1, e2, e3, d = check_r2c(shape, T, flags)
                    ok = e1 < tol[T] * 10 and e2 < tol[T] * 10 and e3 == 0
                    bad += not ok
                    print(f"r2c2d {np.dtype(T).name} {shape} flags={flags} fwd {e1:.2e} rt {e2:.2e} {'OK' if ok else 'FAIL'} [{d}]")
                except Exception as ex:  # noqa: BLE001
                    bad += 1
                    print(f"r2c2d {np.dtype(T).name} {shape} flags={flags} EXC {type(ex).__name__}: {ex}")
        for shape in ((6, 8, 10), (32, 30, 16), (64, 64, 64), (16, 128, 32), (128, 16, 256)):
            try:
                e1, e2, e3, d = check_r2c(shape, T)
                ok = e1 < tol[T] * 10 and e2 < tol[T] * 10 and e3 == 0
                bad += not ok
                print(f"r2c3d {np.dtype(T).name} {shape} fwd {e1:.2e} rt {e2:.2e} {'OK' if ok else 'FAIL'} [{d}]")
                e1, e2, d = check_c2c(shape, T)
                ok = e1 < tol[T] * 10 and e2 < tol[T] * 10
                bad += not ok
                print(f"c2c3d {np.dtype(T).name} {shape} fwd {e1:.2e} rt {e2:.2e} {'OK' if ok else 'FAIL'}")
            except Exception as ex:  # noqa: BLE001
                bad += 1
                print(f"3d {np.dtype(T).name} {shape} EXC {type(ex).__name__}: {ex}")
        try:
            e1, e2, e3, d = check_r2c((64, 32), T, nbatch=3)
            print(f"r2c2d batch3 {np.dtype(T).name} fwd {e1:.2e} rt {e2:.2e} {'OK' if e1 < tol[T] * 10 else 'FAIL'}")
        except Exception as ex:  # noqa: BLE001
            bad += 1
            print(f"batch EXC {type(ex).__name__}: {ex}")
    print("FAILURES:", bad)
    peak = 6550.1
    try:
        peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:  # noqa: BLE001
        pass
    for shape, T in (((4096, 4096), np.float64), ((8192, 8192), np.float64), ((8192, 8192), np.float32), ((512, 512, 512), np.float64),
                     ((512, 512, 512), np.float32), ((1024, 1024, 1024), np.float32), ((2048, 2048), np.float64), ((1024, 1024), np.float64)):
        try:
            res, alg = time_plan(shape, T)
            print(f"time {shape} {np.dtype(T).name}: fwd {res['fwd']:.3f} ms ({alg / res['fwd'] / 1e6:.0f} GB/s, {alg / res['fwd'] / 1e6 / peak:.2f}) "
                  f"inv {res['inv']:.3f} ms ({alg / res['inv'] / 1e6:.0f} GB/s, {alg / res['inv'] / 1e6 / peak:.2f})")
        except Exception as ex:  # noqa: BLE001
            print(f"time {shape} EXC {type(ex).__name__}: {ex}")


if __name__ == "__main__":
    main()
