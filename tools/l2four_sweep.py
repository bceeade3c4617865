"""Timing sweep of the strided-dimension strategies on the GPU box: two-kernel four-step vs the fused (L2-resident) four-step
for several chunk widths / look-ahead distances / thresholds, against the HBM roofline P + (2d-1) S of SURVEY 8d.
Usage: python tools/l2four_sweep.py [quick]"""
import ctypes as C
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fourierflows_jl_b200 as ff  # noqa: E402
from fourierflows_jl_b200 import _lib as L  # noqa: E402

KEYS = ("FFB_L2FOUR", "FFB_L2_CHUNK", "FFB_L2_AHEAD", "FFB_FOURSTEP_MIN", "FFB_L2_PF", "FFB_L2_ACQ", "FFB_ROWS_R8")


def time_plan(shape, T, reps=8):
    import torch
    plan = ff.Plan(shape, T, L.FFB_R2C)
    x = ff.DevArray.zeros(T, shape)
    xh = ff.DevArray.zeros(ff.cxtype(T), plan.spectral_shape)
    st = C.c_void_p()
    L.call("ffb_get_stream", C.byref(st))
    stream = torch.cuda.ExternalStream(st.value)
    res = {}
    with torch.cuda.stream(stream):
        for name, fn in (("fwd", lambda: plan.mul(xh, x)), ("inv", lambda: plan.ldiv(x, xh))):
            for _ in range(3):
                fn()
            L.call("ffb_sync")
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(reps):
                fn()
            e1.record(stream)
            e1.synchronize()
            res[name] = e0.elapsed_time(e1) / reps
    es = np.dtype(T).itemsize
    P = np.prod(shape) * es
    S = np.prod(plan.spectral_shape) * es * 2
    return res, P + (2 * len(shape) - 1) * S, plan.describe()


def main():
    quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
    peak = 6550.1
    try:
        peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:  # noqa: BLE001
        pass
    f64, f32 = np.float64, np.float32
    cases = [((8192, 8192), f64), ((4096, 4096), f64), ((8192, 8192), f32), ((1024, 1024, 1024), f32), ((2048, 2048, 256), f32),
             ((1024, 1024, 256), f64)]
    variants = [{"FFB_L2FOUR": "0"}, {}, {"FFB_L2_CHUNK": "2"}, {"FFB_L2_CHUNK": "4"}, {"FFB_L2_AHEAD": "10"}, {"FFB_L2_AHEAD": "40"},
                {"FFB_FOURSTEP_MIN": "1024"}, {"FFB_FOURSTEP_MIN": "1024", "FFB_L2FOUR": "0"}, {"FFB_FOURSTEP_MIN": "100000"}]
    if quick:
        cases = cases[:3]
        cases = [((2048, 2048, 256), f32), ((1024, 1024, 1024), f32), ((1024, 1024, 256), f64), ((2048, 2048), f64), ((4096, 4096), f32)]
        variants = [{}, {"FFB_FOURSTEP_MIN": "2048"}, {"FFB_FOURSTEP_MIN": "1024"}, {"FFB_FOURSTEP_MIN": "2048", "FFB_L2FOUR": "1"}]
    for shape, T in cases:
        for v in variants:
            for k in KEYS:
                os.environ.pop(k, None)
            os.environ.update(v)
            try:
                res, alg, desc = time_plan(shape, T)
                print(f"{'x'.join(map(str, shape)):>14} {np.dtype(T).name} {json.dumps(v):<52} fwd {res['fwd']:.3f} ms ({alg / res['fwd'] / 1e6 / peak:.3f}) "
                      f"inv {res['inv']:.3f} ms ({alg / res['inv'] / 1e6 / peak:.3f})  [{desc.strip()}]", flush=True)
            except Exception as ex:  # noqa: BLE001
                print(f"{shape} {np.dtype(T).name} {v} EXC {type(ex).__name__}: {ex}", flush=True)


if __name__ == "__main__":
    main()
