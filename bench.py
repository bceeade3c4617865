#!/usr/bin/env python
"""Benchmark of the pseudospectral time-stepping hot path (BASELINE.json metric: ETDRK4 steps/s and Gpt*steps/s).

    python bench.py --gpus N --steps K --warmup W            # this library on N B200s (torchrun for N > 1)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path restated (oracle), host cores

One "step" = one ETDRK4 `stepforward!` (4 calcN! with 5 2-D FFTs each + 4 fused stage kernels) of the 2-D vorticity
problem (user calcN! + dealias!) on the 8192^2 Float64 grid: config C3 of BASELINE.json / SURVEY 8d, the configuration
the metric is quoted on for one GPU.  Prints ONE JSON line (rank 0).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NU, DT, K0 = 1e-4, 1e-3, 64.0


def work(n, d=2):
    """N log2 N work model used to scale bounded CPU samples to the full grid."""
    return float(n) ** d * np.log2(float(n) ** d)


def hbm_peak():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"], "measured (MEASURED_PEAKS.json)"
    except Exception:  # noqa: BLE001
        return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------ byte model
def step_bytes(n, es=8):
    """Algorithmic HBM bytes of one ETDRK4 step of the 2-D vorticity problem."""
    nkr = n // 2 + 1
    S, P, R = nkr * n * 2 * es, n * n * es, nkr * n * es
    fft = P + 3 * S
    calcN = (S + R + 2 * S) + 3 * fft + 5 * P + 2 * fft + 3 * S       # prep (no zeta_h copy) + 3 irfft + products + 2 rfft + combine
    stages = (3 * S + 2 * R) * 2 + (4 * S + 2 * R) + (6 * S + 4 * R)   # substep12 x2, substep3, update (dense real Float64 coefficients)
    return 4 * calcN + stages, fft


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 8 for i in range(4) if r[4 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ reference arm / cpu baseline
def cpu_arm(n_sample, steps, warmup, n_full):
    """The reference's CPU path restated (oracle: NumPy + pocketfft with all host threads) on a bounded sample grid;
    steps/s scaled to the full grid by the N log2 N work model."""
    import oracle as fo
    cores = os.cpu_count() or 1
    fo.set_fft_workers(cores)
    prob = fo.TwoDNavierStokes.Problem(nx=n_sample, nu=NU, dt=DT, stepper="ETDRK4")
    z0 = fo.random_phase_field((n_sample, n_sample), 2 * np.pi, K0 * n_sample / n_full, slope=-1.0, seed=1234)
    prob.grid.rfftplan.mul(prob.sol, z0)
    fo.stepforward(prob, warmup)
    t0 = time.perf_counter()
    fo.stepforward(prob, steps)
    dt = (time.perf_counter() - t0) / steps
    scale = work(n_sample) / work(n_full)
    assert np.isfinite(prob.sol).all()
    return {"value": scale / dt, "unit": "steps/s", "cores": cores, "kind": "port",
            "sample": f"{steps} ETDRK4 step(s) of the same 2-D vorticity problem at {n_sample}^2 Float64 ({dt:.3f} s/step measured, pocketfft workers={cores}), "
                      f"scaled to {n_full}^2 by N*log2(N) (x{scale:.4f}); reference CPU path restated in NumPy, not FFTW (no Julia/FFTW in the image)"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_full = args.n
    n_sample = min(n_full, 2048)
    cb = cpu_arm(n_sample, args.steps, max(1, args.warmup), n_full)
    v = cb["value"]
    line = {"impl": "reference", "metric": "ETDRK4 steps/s (2-D vorticity, user calcN! + dealias!)", "value": v, "unit": "steps/s",
            "gpt_steps_per_s": v * n_full * n_full / 1e9, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 / v, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"C3: 2-D vorticity ETDRK4 {n_full}^2 Float64 (TwoDGrid, aliased_fraction=1/3, nu={NU}, dt={DT})"},
            "cpu_baseline": cb, "e2e": {"value": v, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ GPU arm
def run_gpu(args):
    import torch
    import fourierflows_jl_b200 as ff
    from fourierflows_jl_b200 import _lib as L
    import oracle as fo  # only for the synthetic initial condition and the cpu_baseline leg

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not ff.have_device():
        raise SystemExit("bench.py needs a CUDA device: libfourierflows_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    L.call("ffb_set_device", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    stream = torch.cuda.Stream()
    L.call("ffb_set_stream", stream.cuda_stream)
    n = args.n
    peak, peak_src = hbm_peak()

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    with torch.cuda.stream(stream):
        prob = ff.CProblem((n, n), 2 * np.pi, stepper="ETDRK4", dt=DT, calcN="vorticity2d", nu=NU, T=np.float64)
        z0 = fo.random_phase_field((n, n), 2 * np.pi, K0, slope=-1.0, seed=1234 + rank)
        prob.set_physical(z0)
        # ---------------- value: K steps, state resident in HBM ----------------
        prob.stepforward(args.warmup)
        barrier()
        sampler = ClockSampler(local)
        sampler.start()
        l0 = ff.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        prob.stepforward(args.steps)
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        launches = ff.launch_count() - l0
        clocks = sampler.stop()
        if world > 1:
            import torch.distributed as dist
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        ms_per_step = ms / args.steps
        assert np.isfinite(prob.sol.to_numpy()[:8, :8]).all()
        # ---------------- per-kernel roofline: same steps with CUDA events around every kernel launch ----------------
        ff.prof_enable(True)
        prob.stepforward(min(args.steps, 5))
        rep = ff.prof_report()
        ff.prof_enable(False)
        tot_ms = sum(r["ms"] for r in rep)
        rep.sort(key=lambda r: -r["ms"])
        top = rep[0]
        kernels = [{"name": r["name"], "share": round(r["ms"] / tot_ms, 4), "us_per_launch": round(1e3 * r["ms"] / r["launches"], 2),
                    "gbs": round(r["bytes"] / r["ms"] / 1e6, 1), "frac": round(r["bytes"] / r["ms"] / 1e6 / peak, 4)} for r in rep]
        roofline = {"bound": "hbm", "kernel": top["name"], "achieved": top["bytes"] / top["ms"] / 1e6, "peak": peak, "unit": "GB/s",
                    "frac": top["bytes"] / top["ms"] / 1e6 / peak, "traffic": None, "peak_source": peak_src,
                    "share_of_step": top["ms"] / tot_ms, "algorithmic_bytes_per_launch": top["bytes"] / top["launches"]}
        # ---------------- FFT % of HBM peak: standalone 2-D r2c / c2r at the same size ----------------
        plan = ff.Plan((n, n), np.float64, L.FFB_R2C)
        x = ff.DevArray.zeros(np.float64, (n, n))
        xh = ff.DevArray.zeros(np.complex128, plan.spectral_shape)
        fft_ms = {}
        for name, fn in (("rfft", lambda: plan.mul(xh, x)), ("irfft", lambda: plan.ldiv(x, xh))):
            for _ in range(3):
                fn()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            for _ in range(10):
                fn()
            b.record(stream)
            b.synchronize()
            fft_ms[name] = a.elapsed_time(b) / 10
        del x, xh, plan
        # ---------------- e2e: host buffers in, host buffers out, through the C ABI ----------------
        S = prob.sol.nbytes
        hp = C.c_void_p()
        L.call("ffb_host_alloc_pinned", C.byref(hp), S)
        L.call("ffb_d2h", hp, prob.sol.ptr, S)
        ke = max(1, min(args.steps, 5))
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(ke):
            L.call("ffb_h2d", prob.sol.ptr, hp, S)      # this step's input state from pinned host memory
            prob.stepforward(1)                          # public call: ffb_step
            L.call("ffb_d2h", hp, prob.sol.ptr, S)       # result back to the host (blocking)
        b.record(stream)
        barrier()
        e2e_ms = a.elapsed_time(b) / ke
        L.call("ffb_host_free_pinned", hp)
        dev_bytes = prob.device_bytes()

    total_bytes, fft_bytes = step_bytes(n)
    value = world * 1e3 / ms_per_step  # independent replicas when world > 1 (see config.parallelism)
    line = {
        "metric": "ETDRK4 steps/s (2-D vorticity, user calcN! + dealias!)", "value": value, "unit": "steps/s",
        "gpt_steps_per_s": value * n * n / 1e9,
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak" if world > 1 else "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"C3: 2-D vorticity ETDRK4 {n}^2 Float64 (TwoDGrid, aliased_fraction=1/3, nu={NU}, dt={DT}), random-phase IC seed 1234",
                   "grid": [n, n], "stepper": "ETDRK4", "parallelism": "single GPU" if world == 1 else f"{world} independent replicas",
                   "l2": f"every array is {S / 1e6:.0f} MB > 126 MB L2; no flush needed", "device_bytes": dev_bytes},
        "step_roofline": {"algorithmic_bytes_per_step": total_bytes, "ms_at_peak": total_bytes / peak / 1e6,
                          "frac": (total_bytes / peak / 1e6) / ms_per_step, "peak_gbs": peak, "peak_source": peak_src},
        "fft": {"rfft_ms": fft_ms["rfft"], "irfft_ms": fft_ms["irfft"], "algorithmic_bytes": fft_bytes,
                "rfft_frac_hbm_peak": fft_bytes / fft_ms["rfft"] / 1e6 / peak, "irfft_frac_hbm_peak": fft_bytes / fft_ms["irfft"] / 1e6 / peak},
        "roofline": roofline, "kernels": kernels,
        "e2e": {"value": world * 1e3 / e2e_ms, "unit": "steps/s", "h2d_bytes_per_step": S, "d2h_bytes_per_step": S, "ms_per_step": e2e_ms},
        "gpu_launches": launches, "clocks": clocks,
    }
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_arm(min(n, 4096) if args.cpu_sample == 0 else args.cpu_sample, 1, 0, n)
        print(json.dumps(line))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", type=int, default=8192, help="grid size per side (default: the C3 configuration)")
    ap.add_argument("--cpu-sample", type=int, default=0, help="grid size of the cpu_baseline sample (default 4096)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
