#!/usr/bin/env python
"""Benchmark of the pseudospectral time-stepping hot path (BASELINE.json metric: ETDRK4 steps/s and Gpt*steps/s at
8192^2 / 2048^3 on 1-8 B200; FFT % of HBM peak).

    python bench.py --gpus N --steps K --warmup W            # this library on N B200s (torchrun for N > 1)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path restated (oracle), host cores

Workloads (`--workload`; one "step" = one `stepforward!` of the named stepper):
  c3 (default at N = 1) : BASELINE configs[2] -- 2-D vorticity (user calcN! + dealias!) on TwoDGrid 8192^2 Float64, ETDRK4:
                          the configuration the metric is quoted on for one GPU.
  c5 (default at N > 1) : configs[4], ETDRK4 -- 3-D Burgers-like equation on ThreeDGrid (2048, 2048, 256*N) Float32,
                          slab-decomposed; weak scaling in z, reaching the 2048^3 grid of the metric at N = 8.  The line also
                          carries `weak_ref`: the same per-GPU problem on ONE GPU without exchange, timed in the same run.
  c4                    : configs[3] -- ThreeDGrid 1024^3 Float64, FilteredRK4, strong scaling over N = 1/2/4/8.
  c5-lsrk54             : configs[4], LSRK54 -- ThreeDGrid 2048^3 Float32, strong scaling over N = 2/4/8.
  c2                    : configs[1] -- TwoDGrid 4096^2 Float64 rfft + 2 x (ik, irfft) derivative round trip (1 GPU).
  c3-slab               : the C3 problem slab-decomposed over N GPUs (physical y-slabs <-> spectral kx blocks; north star: "2D grids beyond
                          one GPU's HBM"), strong scaling.
`value` is Gpt*steps/s (grid points x steps / s / 1e9) so that all N share one unit; `steps_per_s` is reported beside it.
Prints ONE JSON line (rank 0).
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

EXCHANGE = "auto"
P2P = 1             # N > 1: exchange through peer memory (--no-p2p falls back to the chunked NCCL all-to-all)
FUSED = 1           # fold calcN!'s spectral multiplies / products / dealias into the FFT passes (--no-fuse disables)
NVLINK_GBS = 770.0  # measured peer copy per direction per GPU (B200_PROFILING.md)
METRIC = "Gpt*steps/s (grid points x steps per second / 1e9)"


def hbm_peak():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"], "measured (MEASURED_PEAKS.json)"
    except Exception:  # noqa: BLE001
        return 6650.0, "fallback (B200_PROFILING.md)"


def nlogn(shape):
    n = float(np.prod([float(s) for s in shape]))
    return n * np.log2(n)


def stage_bytes(stepper, S, R, cw):
    """Algorithmic HBM bytes of one step's fused stage kernels: S = complex state array, R = dense real array of
    the state's width, cw = bytes of a dense real coefficient array (Float64 ETD coefficients for Float64 problems)."""
    filt = stepper.startswith("Filtered")
    base = stepper[len("Filtered"):] if filt else stepper
    f = R if filt else 0
    if base == "ETDRK4":
        return (3 * S + 2 * cw) * 2 + (4 * S + 2 * cw) + (6 * S + 4 * cw) + f
    if base == "RK4":
        return (4 * S + R) + 2 * (5 * S + R) + (7 * S + R) + f
    if base == "LSRK54":
        return (4 * S + R) + 4 * (5 * S + R) + f
    if base == "AB3":
        return 6 * S + R + f
    return 3 * S + R + f


PIPELINE_MAX_BYTES = 3 << 29   # e2e through the host-buffer pipeline when the state is at most 1.5 GB per GPU (6 pinned host buffers)
NCALC = {"ETDRK4": 4, "RK4": 4, "LSRK54": 5, "AB3": 1, "ForwardEuler": 1}


# ------------------------------------------------------------------------------------------------ workloads
class Workload:
    l2 = "no flush between timed steps: every array a step touches is far larger than the 126 MB L2"
    replicas = 1
    decomposed = False
    scaling = "strong"
    kind = "step"

    def points(self):
        return float(np.prod(self.shape))

    def config(self):
        """identical in both arms (the driver compares them)"""
        return {"workload": self.name, "grid": list(self.shape), "stepper": self.stepper, "precision": self.dtype, "l2": self.l2}


class Vorticity2D(Workload):
    """C3 of SURVEY 8d: TwoDGrid(nx=n, Lx=2pi, aliased_fraction=1/3), L = -nu*Krsq, nu=1e-4, dt=1e-3, random-phase IC."""
    nu, dt, K0 = 1e-4, 1e-3, 64.0
    dtype, T = "f64", np.float64
    key = "c3"

    def __init__(self, n, world, stepper="ETDRK4", slab=False):
        self.shape, self.world, self.stepper = (n, n), (world if slab else 1), stepper
        self.name = (f"C3: 2-D vorticity {stepper} {n}^2 Float64 (TwoDGrid, aliased_fraction=1/3, nu={self.nu}, dt={self.dt}), "
                     "random-phase IC seed 1234")
        self.decomposed = bool(slab) and world > 1
        self.key = "c3-slab" if slab else "c3"
        if self.decomposed:
            self.parallelism = f"slab decomposition x{world}: physical y-slabs <-> spectral kx blocks, one NCCL all-to-all per 2-D transform"
            self.replicas = 1
        else:
            self.parallelism = "single GPU" if world == 1 else f"{world} independent replicas"
            self.replicas = world

    def bytes_per_step(self):
        n, es = self.shape[0], 8
        nkr = n // 2 + 1
        S, P, R = nkr * n * 2 * es, n * n * es, nkr * n * es
        fft = P + 3 * S
        calcN = (S + R + 2 * S) + 3 * fft + 5 * P + 2 * fft + 3 * S      # prep (no zeta_h copy) + 3 irfft + products + 2 rfft + combine
        base = self.stepper.replace("Filtered", "")
        w = self.world if self.decomposed else 1
        nvlink = 5 * NCALC[base] * (S / w) * (w - 1) / w                  # five transforms per calcN!, one exchange each
        return NCALC[base] * calcN + stage_bytes(self.stepper, S, R, R), fft, nvlink

    def make_gpu(self, ff, fo, rank, comm, shape=None, seed=None):
        shape = shape or self.shape
        prob = ff.CProblem(shape, 2 * np.pi, stepper=self.stepper, dt=self.dt, calcN="vorticity2d", nu=self.nu, T=self.T, fused=FUSED, dist=comm)
        field = fo.random_phase_field(shape, 2 * np.pi, self.K0 * shape[0] / self.shape[0], slope=-1.0,
                                      seed=(1234 + (0 if comm is not None else rank)) if seed is None else seed)
        prob.set_physical(ff.physical_slab_2d(field, comm.nranks, comm.rank) if comm is not None else field)
        return prob

    def fft_plan(self, ff, L, comm):
        return ff.DistPlan(self.shape, self.T, comm) if comm is not None else ff.Plan(self.shape, self.T, L.FFB_R2C)

    def cpu_sizes(self):
        return [self.shape, (4096, 4096), (2048, 2048), (1024, 1024)]

    def make_cpu(self, fo, shape):
        prob = fo.TwoDNavierStokes.Problem(nx=shape[0], nu=self.nu, dt=self.dt, stepper=self.stepper)
        z0 = fo.random_phase_field(tuple(shape), 2 * np.pi, self.K0 * shape[0] / self.shape[0], slope=-1.0, seed=1234)
        prob.grid.rfftplan.mul(prob.sol, z0)
        return prob


class Burgers3D(Workload):
    """C4 / C5 of SURVEY 8d: ThreeDGrid, L = -kappa*Krsq, N = -1/2 im kr rfft(irfft(sol)^2) + dealias!; slab-decomposed over the
    ranks.  ETDRK4 in Float32 stores T-width coefficients (the reference's Float64 ones do not fit: SURVEY 8d C5)."""
    kappa, dt, K0 = 1e-3, 1e-3, 32.0
    decomposed = True

    def __init__(self, key, shape, world, stepper, T, scaling, note):
        self.key, self.shape, self.world, self.stepper, self.scaling = key, tuple(shape), world, stepper, scaling
        self.T, self.dtype = T, ("f32" if T == np.float32 else "f64")
        self.name = (f"{note}: 3-D Burgers-like {stepper} {'x'.join(map(str, self.shape))} {'Float32' if T == np.float32 else 'Float64'} "
                     f"(ThreeDGrid, aliased_fraction=1/3, kappa={self.kappa}, dt={self.dt}), random-phase IC")
        self.parallelism = ("single GPU" if world == 1 else
                            f"slab decomposition x{world}: physical z-slabs <-> spectral y-slabs, one exchange over NVLink per 3-D transform")

    def bytes_per_step(self, shape=None, world=None):
        nx, ny, nz = shape or self.shape
        w = world or self.world
        es = 4 if self.T == np.float32 else 8
        nkr = nx // 2 + 1
        S, P, R = nkr * ny * nz * 2 * es, nx * ny * nz * es, nkr * ny * nz * es
        fft = P + 5 * S
        calcN = 2 * fft + 2 * P + 2 * S
        base = self.stepper.replace("Filtered", "")
        nfft = 2 * NCALC[base]
        nvlink_per_gpu = nfft * (S / w) * (w - 1) / w   # S/P*(P-1)/P bytes out per GPU per transform
        return NCALC[base] * calcN + stage_bytes(self.stepper, S, R, R), fft, nvlink_per_gpu

    def make_gpu(self, ff, fo, rank, comm, shape=None, seed=None):
        shape = shape or self.shape
        kw = {"coef_dtype": np.float32} if self.T == np.float32 else {}
        prob = ff.CProblem(shape, 2 * np.pi, stepper=self.stepper, dt=self.dt, calcN="burgers3d", nu=self.kappa, T=self.T, dist=comm,
                           fused=FUSED, **kw)
        # synthetic field generated per slab (the full grid does not fit one host array at 2048^3): seeded white noise,
        # smoothed by the diffusion-dominated warm-up steps
        rng = np.random.default_rng((1234 + rank) if seed is None else seed)
        sl = rng.standard_normal(prob.physical_shape, dtype=np.float32).astype(self.T, copy=False)
        if P2P and comm is not None:
            prob.enable_p2p(EXCHANGE)
        prob.set_physical(np.asfortranarray(0.1 * sl))
        return prob

    def fft_plan(self, ff, L, comm):
        if comm is None:
            return ff.Plan(self.shape, self.T, L.FFB_R2C)
        plan = ff.DistPlan(self.shape, self.T, comm)
        return plan.enable_p2p(EXCHANGE) if P2P else plan

    def cpu_sizes(self):
        return [(512,) * 3, (256,) * 3, (128,) * 3]

    def make_cpu(self, fo, shape):
        prob = fo.Burgers3D.Problem(nx=shape[0], ny=shape[1], nz=shape[2], kappa=self.kappa, dt=self.dt, stepper=self.stepper, T=self.T)
        c0 = 0.1 * np.random.default_rng(1234).standard_normal(tuple(shape)).astype(self.T)
        prob.grid.rfftplan.mul(prob.sol, np.asfortranarray(c0))
        return prob


class DerivativeRoundTrip(Workload):
    """C2 of SURVEY 8d: TwoDGrid 4096^2 Float64: uh = rfft(u); ux = irfft(i kr uh); uy = irfft(i l uh) (one "step")."""
    dtype, T, key, stepper, kind = "f64", np.float64, "c2", "none (rfft + 2 x (ik, irfft))", "transform"
    l2 = "no flush between timed round trips: each of the five arrays a round trip streams (134 MB at 4096^2) is larger than the 126 MB L2"

    def __init__(self, n, world):
        self.shape, self.world, self.replicas = (n, n), 1, world
        self.name = f"C2: TwoDGrid {n}^2 Float64 rfft/irfft spectral-derivative round trip (d/dx, d/dy via kr, l), random-phase field seed 1234"
        self.parallelism = "single GPU" if world == 1 else f"{world} independent replicas"

    def bytes_per_step(self):
        n, es = self.shape[0], 8
        nkr = n // 2 + 1
        S, P = nkr * n * 2 * es, n * n * es
        fft = P + 3 * S
        return 3 * fft + 2 * 2 * S, fft, 0.0     # 3 transforms + 2 spectral multiplies (read + write S each)

    def fft_plan(self, ff, L, comm):
        return ff.Plan(self.shape, self.T, L.FFB_R2C)

    def cpu_sizes(self):
        return [self.shape, (2048, 2048)]


def make_workload(args, world):
    wl = args.workload
    if wl == "auto":
        wl = "c3" if world == 1 else "c5"
    if wl == "c3":
        return Vorticity2D(args.n, world)
    if wl == "c3-slab":
        return Vorticity2D(args.n, world, slab=True)
    if wl == "c2":
        return DerivativeRoundTrip(args.n if args.n != 8192 else 4096, world)
    if wl == "c5":
        return Burgers3D("c5", (args.n3, args.n3, args.nz_per_gpu * world), world, "ETDRK4", np.float32, "weak",
                         f"C5 (weak scaling in z, {args.nz_per_gpu} planes per GPU)")
    if wl == "c4":
        n = args.n3 if args.n3 != 2048 else 1024
        return Burgers3D("c4", (n, n, n), world, "FilteredRK4", np.float64, "strong", "C4 (strong scaling)")
    if wl == "c5-lsrk54":
        return Burgers3D("c5-lsrk54", (args.n3,) * 3, world, "LSRK54", np.float32, "strong", "C5-LSRK54 (strong scaling)")
    raise SystemExit(f"unknown workload {wl}")


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region (B200_PROFILING.md clocks line).  NVML in a thread (10 ms
    period, no process start-up inside the region); `nvidia-smi -lms` as fallback when pynvml is unusable."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    BITS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index
        self.nvml, self.handle, self.stop_flag, self.thread = None, None, False, None
        self.sm, self.pw, self.reasons, self.mx = [], [], set(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            h = None
            try:
                import torch
                uuid = str(torch.cuda.get_device_properties(index).uuid)
                h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
            except Exception:  # noqa: BLE001
                vis = os.environ.get("CUDA_VISIBLE_DEVICES")
                phys = int(vis.split(",")[index]) if vis and all(v.strip().isdigit() for v in vis.split(",")) else index
                h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            self.nvml, self.handle = pynvml, h
        except Exception:  # noqa: BLE001
            self.nvml = None

    def _sample(self):
        n, h = self.nvml, self.handle
        try:
            self.sm.append(float(n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM)))
            self.pw.append(n.nvmlDeviceGetPowerUsage(h) / 1000.0)
            fn = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or n.nvmlDeviceGetCurrentClocksThrottleReasons
            mask = int(fn(h))
            self.reasons |= {name for bit, name in self.BITS.items() if mask & bit}
        except Exception:  # noqa: BLE001
            pass

    def _loop(self):
        while not self.stop_flag:
            self._sample()
            time.sleep(0.01)

    def start(self):
        if self.nvml is not None:
            self.thread = threading.Thread(target=self._loop, daemon=True)
            self.thread.start()
            return
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
        if self.nvml is not None:
            self.stop_flag = True
            if self.thread:
                self.thread.join(timeout=1.0)
            return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.mx, "reasons": sorted(self.reasons),
                    "power_w_max": max(self.pw) if self.pw else None, "samples": len(self.sm), "source": "nvml"}
        if self.proc:
            self.proc.terminate()
        num = lambda s: s.replace(".", "").isdigit()
        sm = [float(r[1]) for r in self.rows if len(r) >= 8 and num(r[1])]
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and num(r[2])]
        pw = [float(r[3]) for r in self.rows if len(r) >= 8 and num(r[3])]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 8 for i in range(4) if r[4 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "source": "nvidia-smi"}


# ------------------------------------------------------------------------------------------------ CPU arms
def cpu_step_time(wl, fo, shape, steps, warmup):
    """seconds per step of the oracle (the reference's CPU path restated: NumPy + pocketfft, all host threads) at `shape`"""
    if wl.kind == "transform":
        g = fo.TwoDGrid(nx=shape[0], Lx=2 * np.pi, dense=False)
        u = fo.random_phase_field(tuple(shape), 2 * np.pi, 64.0 * shape[0] / 4096, slope=1.0, seed=1234)

        def once():
            uh = g.rfftplan * u
            return g.rfftplan.solve((1j * g.kr) * uh), g.rfftplan.solve((1j * g.l) * uh)
        for _ in range(warmup):
            once()
        t0 = time.perf_counter()
        for _ in range(steps):
            once()
        return (time.perf_counter() - t0) / steps
    prob = wl.make_cpu(fo, shape)
    if warmup:
        fo.stepforward(prob, warmup)
    t0 = time.perf_counter()
    fo.stepforward(prob, steps)
    dt = (time.perf_counter() - t0) / steps
    assert np.isfinite(prob.sol).all()
    return dt


def cpu_cost_guess(wl, shape, cores):
    """rough seconds per oracle step (calibrated on 16 EPYC cores: 8192^2 Float64 ETDRK4 vorticity ~ 27 s, 256^3 Float32 ETDRK4
    Burgers ~ 1.6 s); only used to pick which sample sizes fit the time budget"""
    pts = float(np.prod(shape))
    per_pt = 4.0e-7 if len(shape) == 2 else 1.0e-7
    if wl.kind == "transform":
        per_pt = 0.6e-7
    return pts * per_pt * max(1.0, 16.0 / cores) ** 0.5


def cpu_arm(wl, steps, warmup, budget_s):
    """Times the oracle on the largest grid of `wl.cpu_sizes()` that fits `budget_s` -- the workload's own grid when possible
    (then nothing is modelled) -- for (warmup + steps) steps; when the sample is smaller than the workload, ONE step on the next
    larger size that still fits is timed as well so that the N*log2(N) scaling used for the headline can be checked."""
    import oracle as fo
    cores = os.cpu_count() or 1
    fo.set_fft_workers(cores)
    sizes = wl.cpu_sizes()
    fits = [s for s in sizes if cpu_cost_guess(wl, s, cores) * (steps + warmup) <= budget_s] or [sizes[-1]]
    sample = fits[0]
    t = cpu_step_time(wl, fo, sample, steps, warmup)
    loop_grid, loop_t = sample, t
    full = tuple(sample) == tuple(wl.shape)
    check = None
    if not full:
        bigger = [s for s in sizes if np.prod(s) > np.prod(sample) and cpu_cost_guess(wl, s, cores) <= 0.6 * budget_s]
        if bigger:
            tb = cpu_step_time(wl, fo, bigger[0], 1, 0)
            model = t * nlogn(bigger[0]) / nlogn(sample)
            check = {"grid": list(bigger[0]), "s_per_step_measured": tb, "s_per_step_modelled_from_sample": model, "measured_over_model": tb / model}
            if tb / model > 1.0:   # the larger measurement is the better basis for the headline: scale from it
                sample, t = bigger[0], tb
                full = tuple(sample) == tuple(wl.shape)
    scale = 1.0 if full else nlogn(sample) / nlogn(wl.shape)
    sps = scale / t
    how = "measured at the workload's own grid (nothing modelled)" if full else \
        f"scaled to {'x'.join(map(str, wl.shape))} by N*log2(N) (x{scale:.3e}) from the largest grid that fits the time / memory budget"
    out = {"value": sps * wl.points() / 1e9, "unit": "Gpt*steps/s", "steps_per_s": sps, "cores": cores, "kind": "port", "same_grid": bool(full),
           "sample": f"{wl.stepper} step(s) of the same problem on a {'x'.join(map(str, sample))} {wl.dtype} grid: {t:.3f} s/step "
                     f"(pocketfft workers={cores}; NumPy elementwise single-threaded like Julia's CPU broadcast), {how}; reference CPU path "
                     "restated in NumPy + pocketfft, not FFTW (no Julia / FFTW in the image)",
           "sample_grid": list(sample), "sample_s_per_step": t, "loop_grid": list(loop_grid), "loop_s_per_step": loop_t,
           "loop_steps": steps, "loop_warmup": warmup}
    if check:
        out["model_check"] = check
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = make_workload(args, args.gpus)
    cb = cpu_arm(wl, args.steps, max(0, args.warmup), args.cpu_budget)
    line = {"impl": "reference", "metric": "ETDRK4 " + METRIC if "ETDRK4" in wl.stepper else wl.stepper + " " + METRIC,
            "value": cb["value"], "unit": "Gpt*steps/s", "steps_per_s": cb["steps_per_s"], "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * cb["loop_s_per_step"],
            "ms_per_step_is": "the oracle's step in the K-step loop on cpu_baseline.loop_grid; `value` comes from cpu_baseline.sample_grid",
            "higher_is_better": True, "scaling": wl.scaling if args.gpus > 1 else "strong",
            "vs_baseline": None, "dtype": wl.dtype, "data": "synthetic", "config": wl.config(),
            "cpu_baseline": cb, "e2e": {"value": cb["value"], "unit": "Gpt*steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ GPU arm
def parity_probe(wl, ff, fo, rank, world, comm, all_reduce_sum):
    """Before anything is timed: the benchmarked code path (same equation, stepper, precision, fusion and exchange) on a grid the
    CPU oracle finishes in seconds, compared with the oracle as the relative L2 error of the GLOBAL state."""
    tol = 1e-12 if wl.T == np.float64 else 1e-5
    if wl.kind == "transform":
        return None
    if wl.decomposed and len(wl.shape) == 2:
        # 2-D slab decomposition: transforms and two steps of the benchmarked problem at 1024^2, global rel-L2 vs the oracle
        shape = (1024, 1024)
        P = world
        rng = np.random.default_rng(98)
        x = np.asfortranarray(rng.standard_normal(shape).astype(wl.T))
        ref = fo.RfftPlan(shape, wl.T) * x.astype(np.float64)
        plan = ff.DistPlan(shape, wl.T, comm)
        xh = plan * ff.DevArray.from_numpy(ff.physical_slab_2d(x, P, rank))
        r = ff.spectral_slab_2d(ref, P, rank)
        n1, d1 = all_reduce_sum(float(np.sum(np.abs(xh.to_numpy() - r) ** 2))), all_reduce_sum(float(np.sum(np.abs(r) ** 2)))
        back = plan.solve(xh).to_numpy() - ff.physical_slab_2d(x, P, rank)
        n2, d2 = all_reduce_sum(float(np.sum(back ** 2))), all_reduce_sum(float(np.sum(ff.physical_slab_2d(x, P, rank) ** 2)))
        del plan, xh
        prob = wl.make_gpu(ff, fo, rank, comm, shape=shape, seed=1234)
        oprob = wl.make_cpu(fo, shape)
        prob.stepforward(2)
        fo.stepforward(oprob, 2)
        r = ff.spectral_slab_2d(oprob.sol, P, rank)
        num, den = all_reduce_sum(float(np.sum(np.abs(prob.sol.to_numpy() - r) ** 2))), all_reduce_sum(float(np.sum(np.abs(r) ** 2)))
        prob.close()
        e_fft, e_step = max((n1 / d1) ** 0.5, (n2 / d2) ** 0.5), (num / den) ** 0.5
        return {"grid": list(shape), "fft_rel_l2": e_fft, "step_rel_l2_after_2": e_step, "tol_per_step": tol, "exchange": "nccl",
                "ok": bool(e_fft <= tol and e_step <= 2 * tol)}
    if not wl.decomposed:
        shape = (1024, 1024)
        prob = wl.make_gpu(ff, fo, 0, None, shape=shape, seed=1234)
        oprob = wl.make_cpu(fo, shape)
        e0 = float(np.linalg.norm(prob.sol.to_numpy() - oprob.sol) / np.linalg.norm(oprob.sol))
        prob.stepforward(2)
        fo.stepforward(oprob, 2)
        e = float(np.linalg.norm(prob.sol.to_numpy() - oprob.sol) / np.linalg.norm(oprob.sol))
        prob.close()
        return {"grid": list(shape), "fft_rel_l2": e0, "step_rel_l2_after_2": e, "tol_per_step": tol, "ok": bool(e0 <= tol and e <= 2 * tol)}
    shape = (256, 256, 256)
    fo.set_fft_workers(max(1, (os.cpu_count() or 1) // max(world, 1)))
    P = max(world, 1)
    rng = np.random.default_rng(99)
    x = np.asfortranarray(rng.standard_normal(shape).astype(wl.T))
    ref = fo.RfftPlan(shape, wl.T) * x.astype(np.float64)
    sl = (lambda a: ff.physical_slab(a, P, rank)) if comm is not None else (lambda a: a)
    ssl = (lambda a: ff.spectral_slab(a, P, rank)) if comm is not None else (lambda a: a)
    if comm is not None:
        plan = ff.DistPlan(shape, wl.T, comm)
        if P2P:
            plan.enable_p2p(EXCHANGE)
    else:
        plan = ff.Plan(shape, wl.T, ff._lib.FFB_R2C)
    xl = ff.DevArray.from_numpy(sl(x))
    xh = plan * xl
    d = xh.to_numpy() - ssl(ref)
    n1, d1 = all_reduce_sum(float(np.sum(np.abs(d) ** 2))), all_reduce_sum(float(np.sum(np.abs(ssl(ref)) ** 2)))
    back = plan.solve(xh).to_numpy() - sl(x)
    n2, d2 = all_reduce_sum(float(np.sum(back.astype(np.float64) ** 2))), all_reduce_sum(float(np.sum(sl(x).astype(np.float64) ** 2)))
    del plan, xl, xh
    e_fft = max((n1 / d1) ** 0.5, (n2 / d2) ** 0.5)
    ob = fo.Burgers3D.Problem(nx=shape[0], kappa=wl.kappa, dt=wl.dt, stepper=wl.stepper, T=wl.T)
    c0 = fo.random_phase_field(shape, 2 * np.pi, 8.0, slope=0, seed=1234, T=wl.T)
    ob.grid.rfftplan.mul(ob.sol, c0)
    kw = {"coef_dtype": np.float32} if wl.T == np.float32 else {}
    cp = ff.CProblem(shape, 2 * np.pi, stepper=wl.stepper, dt=wl.dt, calcN="burgers3d", nu=wl.kappa, T=wl.T, dist=comm, fused=FUSED, **kw)
    if comm is not None and P2P:
        cp.enable_p2p(EXCHANGE)
    cp.set_physical(sl(c0))
    cp.stepforward(2)
    fo.stepforward(ob, 2)
    r = ssl(ob.sol)
    num, den = all_reduce_sum(float(np.sum(np.abs(cp.sol.to_numpy() - r) ** 2))), all_reduce_sum(float(np.sum(np.abs(r) ** 2)))
    e_step = (num / den) ** 0.5
    used = getattr(cp, "exchange", "nccl") if comm is not None else None
    cp.close()
    return {"grid": list(shape), "fft_rel_l2": e_fft, "step_rel_l2_after_2": e_step, "tol_per_step": tol, "exchange": used,
            "ok": bool(e_fft <= tol and e_step <= 2 * tol)}


def run_gpu(args):
    import torch
    import fourierflows_jl_b200 as ff
    from fourierflows_jl_b200 import _lib as L
    import oracle as fo  # checker only: synthetic initial condition, the parity probe and the cpu_baseline leg

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not ff.have_device():
        raise SystemExit("bench.py needs a CUDA device: libfourierflows_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    L.call("ffb_set_device", local)
    comm = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    wl = make_workload(args, world)
    if world > 1 and wl.decomposed:
        comm = ff.Dist.from_torch()
    stream = torch.cuda.Stream()
    L.call("ffb_set_stream", stream.cuda_stream)
    peak, peak_src = hbm_peak()

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if world == 1:
            return v
        import torch.distributed as dist
        t = torch.tensor([v], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def all_reduce_sum(v):
        if comm is None:
            return v
        import torch.distributed as dist
        t = torch.tensor([v], device="cuda", dtype=torch.float64)
        dist.all_reduce(t)
        return float(t.item())

    def timed(fn, reps):
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(reps):
            fn()
        b.record(stream)
        barrier()
        return max_over_ranks(a.elapsed_time(b)) / reps

    with torch.cuda.stream(stream):
        parity = None if args.no_parity else parity_probe(wl, ff, fo, rank, world, comm, all_reduce_sum)
        # ---------------- weak-scaling reference: this rank's share of the problem on ONE GPU, no exchange ----------------
        weak_ref = None
        if world > 1 and wl.decomposed and wl.scaling == "weak" and not args.no_weak_ref:
            shape1 = (wl.shape[0], wl.shape[1], wl.shape[2] // world)
            p1 = wl.make_gpu(ff, fo, rank, None, shape=shape1)
            p1.stepforward(args.warmup)
            ms1 = timed(lambda: p1.stepforward(1), max(2, min(args.steps, 6)))
            p1.close()
            del p1
            b1, _, _ = wl.bytes_per_step(shape=shape1, world=1)
            weak_ref = {"grid": list(shape1), "ms_per_step_1gpu_no_exchange": ms1, "hbm_frac_1gpu": b1 / peak / 1e6 / ms1}
        if wl.kind == "transform":
            line = run_transform_workload(args, wl, ff, L, fo, timed, stream, peak, peak_src, world, local, barrier)
            if rank == 0:
                print(json.dumps(line))
            return
        prob = wl.make_gpu(ff, fo, rank, comm)
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
        ms_per_step = max_over_ranks(e0.elapsed_time(e1)) / args.steps
        launches = ff.launch_count() - l0
        clocks = sampler.stop()
        head = prob.sol.view(shape=(64,), dtype=prob.sol.dtype).to_numpy()
        assert np.isfinite(head).all(), "state diverged"
        # ---------------- per-kernel roofline: the same steps with CUDA events around every kernel launch ----------------
        ff.prof_enable(True)
        prob.stepforward(min(args.steps, 4))
        rep = ff.prof_report()
        ff.prof_enable(False)
        tot_ms = sum(r["ms"] for r in rep)
        rep.sort(key=lambda r: -r["ms"])
        hbm_rep = [r for r in rep if r["bytes"] > 0]
        top = hbm_rep[0]
        kernels = [{"name": r["name"], "share": round(r["ms"] / tot_ms, 4), "us_per_launch": round(1e3 * r["ms"] / r["launches"], 2),
                    "gbs": round(r["bytes"] / r["ms"] / 1e6, 1), "frac": round(r["bytes"] / r["ms"] / 1e6 / peak, 4)} for r in rep]
        traffic, traffic_src = ncu_traffic(top["name"])
        roofline = {"bound": "hbm", "kernel": top["name"], "achieved": top["bytes"] / top["ms"] / 1e6, "peak": peak, "unit": "GB/s",
                    "frac": top["bytes"] / top["ms"] / 1e6 / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                    "share_of_step": top["ms"] / tot_ms, "algorithmic_bytes_per_launch": top["bytes"] / top["launches"]}
        # ---------------- FFT % of HBM peak: standalone r2c / c2r at the same size ----------------
        plan = BorrowedPlan(prob, L) if (wl.decomposed and comm is not None) else wl.fft_plan(ff, L, comm)   # 2048^3 on 2 GPUs has no room for a second plan
        x = ff.DevArray.zeros(wl.T, plan.physical_shape)
        xh = ff.DevArray.zeros(ff.cxtype(wl.T), plan.spectral_shape)
        fft_ms = {}
        for name, fn in (("rfft", lambda: plan.mul(xh, x)), ("irfft", lambda: plan.ldiv(x, xh))):
            for _ in range(3):
                fn()
            fft_ms[name] = timed(fn, 10)
        del x, xh, plan
        # ---------------- e2e: host buffers in, host buffers out, through the C ABI ----------------
        S = prob.sol.nbytes
        hp = C.c_void_p()
        L.call("ffb_host_alloc_pinned", C.byref(hp), S)
        L.call("ffb_d2h", hp, prob.sol.ptr, S)

        def e2e_step():
            L.call("ffb_h2d", prob.sol.ptr, hp, S)      # this step's input state (this rank's slab) from pinned host memory
            prob.stepforward(1)                          # public call: ffb_step
            L.call("ffb_d2h", hp, prob.sol.ptr, S)       # result back to the host (blocking)
        e2e_serial_ms = timed(e2e_step, max(1, min(args.steps, 5)))
        e2e_ms, e2e_mode = e2e_serial_ms, {"mode": "blocking: ffb_h2d, ffb_step, ffb_d2h one after the other"}
        if S <= PIPELINE_MAX_BYTES and "AB3" not in wl.stepper:
            # the same work through the host-buffer pipeline (ffb_pipeline_*): every step is an independent host state -- uploaded from
            # pinned memory, stepped once, downloaded -- and the copies of neighbouring steps run beside the step on two copy streams
            depth, nsub = 3, max(6, min(args.steps, 12))
            host_state = np.frombuffer((C.c_char * S).from_address(hp.value), dtype=np.uint8)
            ins = [ff.PinnedBuffer((S,), np.uint8) for _ in range(depth)]
            outs = [ff.PinnedBuffer((S,), np.uint8) for _ in range(depth)]
            for b in ins:
                b.array[:] = host_state
            pipe = prob.pipeline(depth)

            def e2e_pipelined():
                for i in range(nsub):
                    pipe.submit(ins[i % depth], outs[i % depth], 1)
                pipe.wait_all()
            e2e_pipelined()                                                                       # warm-up, and the check below
            e2e_step()                                                                            # blocking form on the same input
            same = bool(np.array_equal(outs[0].array, host_state) and np.array_equal(outs[depth - 1].array, host_state))
            e2e_ms = timed(e2e_pipelined, 1) / nsub
            e2e_mode = {"mode": f"pipelined (ffb_pipeline_*, depth {depth}, {nsub} submissions timed): every step is an independent host state, uploaded from "
                                "pinned memory, stepped once and downloaded; the copies of neighbouring steps overlap the step on two copy streams",
                        "matches_blocking_form_bitwise": same,
                        "blocking": {"ms_per_step": e2e_serial_ms, "value": wl.replicas * 1e3 / e2e_serial_ms * wl.points() / 1e9}}
            pipe.close()
            del host_state
            for b in ins + outs:
                b.close()
        L.call("ffb_host_free_pinned", hp)
        dev_bytes = prob.device_bytes()

    total_bytes, fft_bytes, nvlink_bytes = wl.bytes_per_step()
    share = world if wl.decomposed else 1
    sps = wl.replicas * 1e3 / ms_per_step
    gpt = sps * wl.points() / 1e9
    hbm_ms = total_bytes / share / peak / 1e6
    nvl_ms = nvlink_bytes / NVLINK_GBS / 1e6
    nfft = (2 if len(wl.shape) == 3 else 5) * NCALC[wl.stepper.replace("Filtered", "")]
    line = {
        "metric": ("ETDRK4 " if "ETDRK4" in wl.stepper else wl.stepper + " ") + METRIC, "value": gpt, "unit": "Gpt*steps/s", "steps_per_s": sps,
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": wl.scaling if world > 1 else "strong", "vs_baseline": None, "dtype": wl.dtype, "data": "synthetic",
        "config": wl.config(),
        "run": {"parallelism": wl.parallelism, "l2": "every array is far larger than the 126 MB L2; no flush needed", "device_bytes_per_gpu": dev_bytes,
                "calcN_fusion": bool(FUSED), "exchange": exchange_desc(prob, world, wl)},
        "parity": parity,
        "step_roofline": {"algorithmic_hbm_bytes_per_step": total_bytes, "nvlink_bytes_per_gpu_per_step": nvlink_bytes,
                          "hbm_ms_at_peak": hbm_ms, "nvlink_ms_at_770": nvl_ms, "ms_at_roofline_overlapped": max(hbm_ms, nvl_ms),
                          "frac": max(hbm_ms, nvl_ms) / ms_per_step, "frac_non_overlapped": (hbm_ms + nvl_ms) / ms_per_step,
                          "peak_gbs": peak, "peak_source": peak_src},
        "fft": {"rfft_ms": fft_ms["rfft"], "irfft_ms": fft_ms["irfft"], "algorithmic_bytes": fft_bytes,
                "rfft_frac_hbm_peak": fft_bytes / share / fft_ms["rfft"] / 1e6 / peak, "irfft_frac_hbm_peak": fft_bytes / share / fft_ms["irfft"] / 1e6 / peak,
                "alltoall_gbs_per_gpu_rfft": (nvlink_bytes / nfft) / fft_ms["rfft"] / 1e6 if nvlink_bytes else None},
        "roofline": roofline, "kernels": kernels,
        "e2e": {"value": wl.replicas * 1e3 / e2e_ms * wl.points() / 1e9, "unit": "Gpt*steps/s", "steps_per_s": wl.replicas * 1e3 / e2e_ms,
                "h2d_bytes_per_step": S * world, "d2h_bytes_per_step": S * world, "ms_per_step": e2e_ms, **e2e_mode},
        "gpu_launches": launches, "clocks": clocks,
    }
    if weak_ref is not None:
        weak_ref["weak_efficiency_t1_over_tN"] = weak_ref["ms_per_step_1gpu_no_exchange"] / ms_per_step
        line["weak_ref"] = weak_ref
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_arm(wl, 1, 0, args.cpu_budget)
        print(json.dumps(line))
    prob.close()
    del prob
    if comm is not None:
        comm.close()
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


class BorrowedPlan:
    """the problem's own slab plan (`ffb_problem_plan`): same exchange, no second set of scratch / receive buffers"""

    def __init__(self, prob, L):
        self._L, self._h = L, C.c_void_p()
        L.call("ffb_problem_plan", prob._h, C.byref(self._h))
        self.physical_shape, self.spectral_shape = prob.physical_shape, prob.spectral_shape

    def mul(self, out, a):
        self._L.call("ffb_fft_forward", self._h, a.ptr, out.ptr)

    def ldiv(self, out, ah):
        self._L.call("ffb_fft_inverse", self._h, ah.ptr, out.ptr)


def run_transform_workload(args, wl, ff, L, fo, timed, stream, peak, peak_src, world, local, barrier):
    """C2: the derivative round trip through the public API (rfftplan mul / spectral_mul / ldiv), parity vs the analytic derivative"""
    n = wl.shape[0]
    g = ff.TwoDGrid(ff.GPU(), nx=n, Lx=2 * np.pi)
    og = fo.TwoDGrid(nx=n, Lx=2 * np.pi, dense=False)
    xg, yg = og.x.reshape(-1, 1), og.y.reshape(1, -1)
    u = np.asfortranarray(np.sin(3 * xg + 2 * yg) + 0.5 * np.cos(7 * xg - 5 * yg))
    du = ff.DevArray.from_numpy(u)
    uh, dh = ff.DevArray((g.nkr, g.nl), np.complex128), ff.DevArray((g.nkr, g.nl), np.complex128)
    ux, uy = ff.DevArray((n, n), np.float64), ff.DevArray((n, n), np.float64)

    def once():
        g.rfftplan.mul(uh, du)
        ff.spectral_mul(dh, uh, g, coef=1j, px=1)
        g.rfftplan.ldiv(ux, dh)
        ff.spectral_mul(dh, uh, g, coef=1j, py=1)
        g.rfftplan.ldiv(uy, dh)
    once()
    ex = 3 * np.cos(3 * xg + 2 * yg) - 3.5 * np.sin(7 * xg - 5 * yg)
    ey = 2 * np.cos(3 * xg + 2 * yg) + 2.5 * np.sin(7 * xg - 5 * yg)
    err = max(float(np.linalg.norm(ux.to_numpy() - ex) / np.linalg.norm(ex)), float(np.linalg.norm(uy.to_numpy() - ey) / np.linalg.norm(ey)))
    for _ in range(args.warmup):
        once()
    sampler = ClockSampler(local)
    sampler.start()
    l0 = ff.launch_count()
    ms = timed(once, args.steps)
    launches = ff.launch_count() - l0
    clocks = sampler.stop()
    total_bytes, fft_bytes, _ = wl.bytes_per_step()
    fft_ms = {"rfft": timed(lambda: g.rfftplan.mul(uh, du), 10), "irfft": timed(lambda: g.rfftplan.ldiv(ux, uh), 10)}
    hu = np.empty((n, n), dtype=np.float64, order="F")

    def e2e():
        L.call("ffb_h2d", du.ptr, u.ctypes.data, u.nbytes)
        once()
        L.call("ffb_d2h", hu.ctypes.data, ux.ptr, hu.nbytes)
        L.call("ffb_d2h", hu.ctypes.data, uy.ptr, hu.nbytes)
    e2e_ms = timed(e2e, 3)
    rps = wl.replicas * 1e3 / ms
    line = {"metric": "derivative round trips/s (rfft + 2 x (ik, irfft)); value in " + METRIC.replace("steps", "round trips"),
            "value": rps * wl.points() / 1e9, "unit": "Gpt*round-trips/s", "steps_per_s": rps, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": wl.dtype,
            "data": "synthetic", "config": wl.config(), "parity": {"rel_l2_vs_analytic_derivative": err, "tol": 1e-12, "ok": bool(err <= 1e-12)},
            "step_roofline": {"algorithmic_hbm_bytes_per_step": total_bytes, "hbm_ms_at_peak": total_bytes / peak / 1e6,
                              "frac": total_bytes / peak / 1e6 / ms, "peak_gbs": peak, "peak_source": peak_src},
            "roofline": {"bound": "hbm", "kernel": "whole round trip (5 launches chains)", "achieved": total_bytes / ms / 1e6, "peak": peak, "unit": "GB/s",
                         "frac": total_bytes / ms / 1e6 / peak, "traffic": None, "peak_source": peak_src},
            "fft": {"rfft_ms": fft_ms["rfft"], "irfft_ms": fft_ms["irfft"], "algorithmic_bytes": fft_bytes,
                    "rfft_frac_hbm_peak": fft_bytes / fft_ms["rfft"] / 1e6 / peak, "irfft_frac_hbm_peak": fft_bytes / fft_ms["irfft"] / 1e6 / peak},
            "e2e": {"value": wl.replicas * 1e3 / e2e_ms * wl.points() / 1e9, "unit": "Gpt*round-trips/s", "h2d_bytes_per_step": u.nbytes,
                    "d2h_bytes_per_step": 2 * hu.nbytes, "ms_per_step": e2e_ms},
            "gpu_launches": launches, "clocks": clocks}
    if not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_arm(wl, 2, 1, args.cpu_budget)
    return line


def ncu_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel`, from the committed `ncu --set full` summaries
    (profiles/r0*_ncu_full_*_kernels.csv; mean over the captured launches; newest round first).  (None, None) when the kernel was
    not captured."""
    here = os.path.dirname(os.path.abspath(__file__))
    try:
        files = sorted((f for f in os.listdir(os.path.join(here, "profiles")) if "_ncu_full_" in f and f.endswith("_kernels.csv")), reverse=True)
    except OSError:
        return None, None
    for fn in files:
        try:
            lines = [l for l in open(os.path.join(here, "profiles", fn)) if not l.startswith("#")]
            cols = lines[0].strip().split(",")
            vals = [float(l.split(",")[cols.index("traffic_MB")]) for l in lines[1:] if l.split(",")[0] in (kernel + "_fwd", kernel + "_inv", kernel)]
        except (OSError, ValueError, IndexError):
            continue
        if vals:
            return sum(vals) / len(vals) * 1e6, f"profiles/{fn} ({len(vals)} launches)"
    return None, None


def exchange_desc(prob, world, wl):
    if world == 1 or not wl.decomposed:
        return None
    names = {"peer-store": "peer stores over NVLink fused into the FFT pass (blocked receive layout)",
             "copy-engine": "kx-chunked copy-engine pushes into IPC-mapped peer buffers over NVLink, overlapped with the neighbouring chunks' passes",
             "nccl": "chunked NCCL all-to-all"}
    used = getattr(prob, "exchange", "nccl") if (P2P and len(wl.shape) == 3) else "nccl"
    import fourierflows_jl_b200 as ff
    tuned = [t for t in ff.dist.AUTOTUNE_LOG if tuple(t["shape"]) == tuple(wl.shape)][:1] if EXCHANGE == "auto" else []
    return {"used": used, "how": names.get(used, used), "requested": EXCHANGE, "autotune_ms_fwd_plus_inv": tuned[0]["ms"] if tuned else None}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="auto", choices=["auto", "c2", "c3", "c3-slab", "c4", "c5", "c5-lsrk54"])
    ap.add_argument("--n", type=int, default=8192, help="C3 / C2 grid size per side (C2 default 4096)")
    ap.add_argument("--n3", type=int, default=2048, help="3-D grid size (C5: x and y; C4 default 1024)")
    ap.add_argument("--nz-per-gpu", type=int, default=256, help="C5 z-planes per GPU (weak scaling; 256 x 8 = 2048)")
    ap.add_argument("--cpu-budget", type=float, default=150.0, help="seconds of oracle time the CPU legs may spend")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the parity probe that runs before the timed region")
    ap.add_argument("--no-weak-ref", action="store_true", help="N > 1, weak scaling: skip the single-GPU reference of the per-GPU problem")
    ap.add_argument("--no-p2p", action="store_true", help="N > 1: same as --exchange nccl")
    ap.add_argument("--exchange", default="auto", choices=["auto", "nccl", "peer-store", "copy-engine"],
                    help="N > 1: how the slab transpose moves data between GPUs (auto: measured at plan time, fastest kept)")
    ap.add_argument("--no-fuse", action="store_true", help="run calcN! as separate elementwise kernels (the byte model of SURVEY 8d)")
    args = ap.parse_args()
    global FUSED, P2P, EXCHANGE
    FUSED = 0 if args.no_fuse else 1
    EXCHANGE = args.exchange
    P2P = 0 if (args.no_p2p or EXCHANGE == "nccl") else 1
    if args.impl == "reference":
        run_reference(args)
    elif args.gpus > 1 and "WORLD_SIZE" not in os.environ:
        # `python bench.py --gpus N` without a launcher: start one rank per GPU ourselves (the driver uses torchrun directly)
        import socket
        s = socket.socket()
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
        s.close()
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
               "--master-port", str(port), os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
