#!/usr/bin/env python
"""Benchmark of the pseudospectral time-stepping hot path (BASELINE.json metric: ETDRK4 steps/s and Gpt*steps/s at
8192^2 / 2048^3 on 1-8 B200; FFT % of HBM peak).

    python bench.py --gpus N --steps K --warmup W            # this library on N B200s (torchrun for N > 1)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path restated (oracle), host cores

Workloads (one "step" = one ETDRK4 `stepforward!`: 4 calcN! + 4 fused stage kernels):
  N = 1 : config C3 -- 2-D vorticity (user calcN! + dealias!) on TwoDGrid 8192^2 Float64, the configuration the metric
          is quoted on for one GPU.
  N > 1 : config C5 -- 3-D Burgers-like equation on ThreeDGrid (2048, 2048, 256*N) Float32, slab-decomposed with NCCL
          all-to-all transposes; weak scaling in z, reaching the 2048^3 grid of the metric at N = 8.
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
P2P = 1             # N > 1: fused pass + collective exchange (--no-p2p falls back to the chunked NCCL all-to-all)
FUSED = 1           # fold calcN!'s spectral multiplies / products / dealias into the FFT passes (--no-fuse disables)
NVLINK_GBS = 770.0  # measured peer copy per direction per GPU (B200_PROFILING.md)


def hbm_peak():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"], "measured (MEASURED_PEAKS.json)"
    except Exception:  # noqa: BLE001
        return 6650.0, "fallback (B200_PROFILING.md)"


def nlogn(shape):
    n = float(np.prod([float(s) for s in shape]))
    return n * np.log2(n)


# ------------------------------------------------------------------------------------------------ workloads
class Vorticity2D:
    """C3 of SURVEY 8d: TwoDGrid(nx=n, Lx=2pi, aliased_fraction=1/3), L = -nu*Krsq, nu=1e-4, dt=1e-3, random-phase IC."""
    nu, dt, K0 = 1e-4, 1e-3, 64.0
    dtype, T = "f64", np.float64

    def __init__(self, n, world):
        self.shape, self.world = (n, n), 1
        self.name = f"C3: 2-D vorticity ETDRK4 {n}^2 Float64 (TwoDGrid, aliased_fraction=1/3, nu={self.nu}, dt={self.dt}), random-phase IC seed 1234"
        self.parallelism = "single GPU" if world == 1 else f"{world} independent replicas"
        self.replicas = world

    def points(self):
        return float(np.prod(self.shape))

    def bytes_per_step(self):
        n, es = self.shape[0], 8
        nkr = n // 2 + 1
        S, P, R = nkr * n * 2 * es, n * n * es, nkr * n * es
        fft = P + 3 * S
        calcN = (S + R + 2 * S) + 3 * fft + 5 * P + 2 * fft + 3 * S      # prep (no zeta_h copy) + 3 irfft + products + 2 rfft + combine
        stages = (3 * S + 2 * R) * 2 + (4 * S + 2 * R) + (6 * S + 4 * R)  # dense real Float64 coefficients
        return 4 * calcN + stages, fft, 0.0

    def make_gpu(self, ff, fo, rank, comm):
        prob = ff.CProblem(self.shape, 2 * np.pi, stepper="ETDRK4", dt=self.dt, calcN="vorticity2d", nu=self.nu, T=self.T, fused=FUSED)
        prob.set_physical(fo.random_phase_field(self.shape, 2 * np.pi, self.K0, slope=-1.0, seed=1234 + rank))
        return prob

    def fft_plan(self, ff, L, comm):
        return ff.Plan(self.shape, self.T, L.FFB_R2C)

    def make_cpu(self, fo, n_sample):
        prob = fo.TwoDNavierStokes.Problem(nx=n_sample, nu=self.nu, dt=self.dt, stepper="ETDRK4")
        z0 = fo.random_phase_field((n_sample, n_sample), 2 * np.pi, self.K0 * n_sample / self.shape[0], slope=-1.0, seed=1234)
        prob.grid.rfftplan.mul(prob.sol, z0)
        return prob, (n_sample, n_sample)


class Burgers3D:
    """C5 of SURVEY 8d: ThreeDGrid Float32, L = -kappa*Krsq, N = -1/2 im kr rfft(irfft(sol)^2) + dealias!, ETDRK4 with
    T-width coefficients (the reference's Float64 coefficients do not fit: SURVEY 8d C5), slab-decomposed."""
    kappa, dt, K0 = 1e-3, 1e-3, 32.0
    dtype, T = "f32", np.float32

    def __init__(self, nxy, nz_per_gpu, world):
        self.shape, self.world = (nxy, nxy, nz_per_gpu * world), world
        self.name = (f"C5: 3-D Burgers-like ETDRK4 {self.shape} Float32 (ThreeDGrid, aliased_fraction=1/3, kappa={self.kappa}, dt={self.dt}), "
                     f"slab-decomposed over {world} GPUs, weak scaling in z ({nz_per_gpu} planes per GPU), random-phase IC")
        self.parallelism = f"slab decomposition x{world}: physical z-slabs <-> spectral y-slabs, one exchange over NVLink per 3-D transform (see config.exchange)"
        self.replicas = 1

    def points(self):
        return float(np.prod(self.shape))

    def bytes_per_step(self):
        nx, ny, nz = self.shape
        es = 4
        nkr = nx // 2 + 1
        S, P, R = nkr * ny * nz * 2 * es, nx * ny * nz * es, nkr * ny * nz * es
        fft = P + 5 * S
        calcN = 2 * fft + 2 * P + 2 * S
        stages = (3 * S + 2 * R) * 2 + (4 * S + 2 * R) + (6 * S + 4 * R)
        w = self.world
        nvlink_per_gpu = 8 * (S / w) * (w - 1) / w   # 8 transforms per step, S/P*(P-1)/P bytes out per GPU each
        return 4 * calcN + stages, fft, nvlink_per_gpu

    def make_gpu(self, ff, fo, rank, comm):
        prob = ff.CProblem(self.shape, 2 * np.pi, stepper="ETDRK4", dt=self.dt, calcN="burgers3d", nu=self.kappa, T=self.T,
                           coef_dtype=np.float32, dist=comm)
        # synthetic random field generated per slab on the device side of the API (rfft of uniform noise would need the
        # full grid on one host): seeded white noise, smoothed by a few diffusion-dominated steps during warm-up
        rng = np.random.default_rng(1234 + rank)
        sl = rng.standard_normal(prob.physical_shape, dtype=np.float32)
        if P2P and comm is not None:
            prob.enable_p2p(EXCHANGE)   # exchange through IPC-mapped peer memory over NVLink instead of the NCCL all-to-all
        prob.set_physical(np.asfortranarray(0.1 * sl))
        return prob

    def fft_plan(self, ff, L, comm):
        plan = ff.DistPlan(self.shape, self.T, comm)
        return plan.enable_p2p(EXCHANGE) if P2P else plan

    def make_cpu(self, fo, n_sample):
        prob = fo.Burgers3D.Problem(nx=n_sample, kappa=self.kappa, dt=self.dt, stepper="ETDRK4", T=self.T)
        c0 = 0.1 * np.random.default_rng(1234).standard_normal((n_sample,) * 3).astype(np.float32)
        prob.grid.rfftplan.mul(prob.sol, np.asfortranarray(c0))
        return prob, (n_sample,) * 3


def make_workload(args, world):
    if world == 1 or args.workload == "c3":
        return Vorticity2D(args.n, world)
    return Burgers3D(args.n3, args.nz_per_gpu, world)


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
def cpu_arm(wl, n_sample, steps, warmup):
    """The reference's CPU path restated (oracle: NumPy + pocketfft with all host threads) on a bounded sample grid;
    throughput scaled to the full grid by the N log2 N work model."""
    import oracle as fo
    cores = os.cpu_count() or 1
    fo.set_fft_workers(cores)
    prob, sshape = wl.make_cpu(fo, n_sample)
    fo.stepforward(prob, warmup)
    t0 = time.perf_counter()
    fo.stepforward(prob, steps)
    dt = (time.perf_counter() - t0) / steps
    assert np.isfinite(prob.sol).all()
    scale = nlogn(sshape) / nlogn(wl.shape)
    sps = scale / dt
    return {"value": sps * wl.points() / 1e9, "unit": "Gpt*steps/s", "steps_per_s": sps, "cores": cores, "kind": "port",
            "sample": f"{steps} ETDRK4 step(s) of the same problem on a {'x'.join(map(str, sshape))} {wl.dtype} grid ({dt:.3f} s/step measured, "
                      f"pocketfft workers={cores}), scaled to {'x'.join(map(str, wl.shape))} by N*log2(N) (x{scale:.3e}); reference CPU path "
                      f"restated in NumPy + pocketfft, not FFTW (no Julia/FFTW in the image)"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = make_workload(args, args.gpus)
    n_sample = min(wl.shape[0], 2048) if len(wl.shape) == 2 else min(wl.shape[0], 256)
    cb = cpu_arm(wl, n_sample, args.steps, max(1, args.warmup))
    line = {"impl": "reference", "metric": "ETDRK4 Gpt*steps/s (grid points x steps per second / 1e9)", "value": cb["value"], "unit": "Gpt*steps/s",
            "steps_per_s": cb["steps_per_s"], "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 / cb["steps_per_s"], "higher_is_better": True, "scaling": "weak" if args.gpus > 1 else "strong",
            "vs_baseline": None, "dtype": wl.dtype, "data": "synthetic", "config": {"workload": wl.name, "grid": list(wl.shape)},
            "cpu_baseline": cb, "e2e": {"value": cb["value"], "unit": "Gpt*steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
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
    comm = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    wl = make_workload(args, world)
    if world > 1 and wl.replicas == 1:
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

    with torch.cuda.stream(stream):
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
        roofline = {"bound": "hbm", "kernel": top["name"], "achieved": top["bytes"] / top["ms"] / 1e6, "peak": peak, "unit": "GB/s",
                    "frac": top["bytes"] / top["ms"] / 1e6 / peak, "traffic": ncu_traffic(top["name"])[0], "traffic_source": ncu_traffic(top["name"])[1],
                    "peak_source": peak_src,
                    "share_of_step": top["ms"] / tot_ms, "algorithmic_bytes_per_launch": top["bytes"] / top["launches"]}
        # ---------------- FFT % of HBM peak: standalone r2c / c2r at the same size ----------------
        plan = wl.fft_plan(ff, L, comm)
        x = ff.DevArray.zeros(wl.T, plan.physical_shape)
        xh = ff.DevArray.zeros(ff.cxtype(wl.T), plan.spectral_shape)
        fft_ms = {}
        for name, fn in (("rfft", lambda: plan.mul(xh, x)), ("irfft", lambda: plan.ldiv(x, xh))):
            for _ in range(3):
                fn()
            barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            for _ in range(10):
                fn()
            b.record(stream)
            barrier()
            fft_ms[name] = max_over_ranks(a.elapsed_time(b)) / 10
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
            L.call("ffb_h2d", prob.sol.ptr, hp, S)      # this step's input state (this rank's slab) from pinned host memory
            prob.stepforward(1)                          # public call: ffb_step
            L.call("ffb_d2h", hp, prob.sol.ptr, S)       # result back to the host (blocking)
        b.record(stream)
        barrier()
        e2e_ms = max_over_ranks(a.elapsed_time(b)) / ke
        L.call("ffb_host_free_pinned", hp)
        dev_bytes = prob.device_bytes()

    total_bytes, fft_bytes, nvlink_bytes = wl.bytes_per_step()
    sps = wl.replicas * 1e3 / ms_per_step
    gpt = sps * wl.points() / 1e9
    hbm_ms = total_bytes / (world if wl.replicas == 1 else 1) / peak / 1e6
    nvl_ms = nvlink_bytes / NVLINK_GBS / 1e6
    fftw = (world if wl.replicas == 1 else 1)
    line = {
        "metric": "ETDRK4 Gpt*steps/s (grid points x steps per second / 1e9)", "value": gpt, "unit": "Gpt*steps/s", "steps_per_s": sps,
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak" if world > 1 else "strong", "vs_baseline": None, "dtype": wl.dtype, "data": "synthetic",
        "config": {"workload": wl.name, "grid": list(wl.shape), "stepper": "ETDRK4", "parallelism": wl.parallelism,
                   "l2": "every array is far larger than the 126 MB L2; no flush needed", "device_bytes_per_gpu": dev_bytes,
                   "calcN_fusion": bool(FUSED) and world == 1, "exchange": exchange_desc(prob, world, wl)},
        "step_roofline": {"algorithmic_hbm_bytes_per_step": total_bytes, "nvlink_bytes_per_gpu_per_step": nvlink_bytes,
                          "hbm_ms_at_peak": hbm_ms, "nvlink_ms_at_770": nvl_ms, "ms_at_roofline_overlapped": max(hbm_ms, nvl_ms),
                          "frac": max(hbm_ms, nvl_ms) / ms_per_step, "frac_non_overlapped": (hbm_ms + nvl_ms) / ms_per_step,
                          "peak_gbs": peak, "peak_source": peak_src},
        "fft": {"rfft_ms": fft_ms["rfft"], "irfft_ms": fft_ms["irfft"], "algorithmic_bytes": fft_bytes,
                "rfft_frac_hbm_peak": fft_bytes / fftw / fft_ms["rfft"] / 1e6 / peak, "irfft_frac_hbm_peak": fft_bytes / fftw / fft_ms["irfft"] / 1e6 / peak,
                "alltoall_gbs_per_gpu_rfft": (nvlink_bytes / 8) / fft_ms["rfft"] / 1e6 if nvlink_bytes else None},
        "roofline": roofline, "kernels": kernels,
        "e2e": {"value": wl.replicas * 1e3 / e2e_ms * wl.points() / 1e9, "unit": "Gpt*steps/s", "steps_per_s": wl.replicas * 1e3 / e2e_ms,
                "h2d_bytes_per_step": S * world, "d2h_bytes_per_step": S * world, "ms_per_step": e2e_ms},
        "gpu_launches": launches, "clocks": clocks,
    }
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_arm(wl, args.cpu_sample or min(wl.shape[0], 4096), 1, 0)
        print(json.dumps(line))
    del prob
    if comm is not None:
        comm.close()
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


def ncu_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel`, from the committed `ncu --set full` summaries
    (profiles/r01_ncu_full_*_kernels.csv; mean over the captured launches).  (None, None) when the kernel was not captured."""
    here = os.path.dirname(os.path.abspath(__file__))
    for fn in ("r01_ncu_full_step_kernels.csv", "r01_ncu_full_fft3d_f32_kernels.csv"):
        try:
            lines = [l for l in open(os.path.join(here, "profiles", fn)) if not l.startswith("#")]
        except OSError:
            continue
        cols = lines[0].strip().split(",")
        vals = [float(l.split(",")[cols.index("traffic_MB")]) for l in lines[1:] if l.split(",")[0] in (kernel + "_fwd", kernel + "_inv", kernel)]
        if vals:
            return sum(vals) / len(vals) * 1e6, f"profiles/{fn} ({len(vals)} launches)"
    return None, None


def exchange_desc(prob, world, wl):
    if world == 1 or wl.replicas != 1:
        return None
    names = {"peer-store": "peer stores over NVLink fused into the FFT pass (blocked receive layout) + barrier",
             "copy-engine": "kx-chunked copy-engine pushes into IPC-mapped peer buffers over NVLink, overlapped with the neighbouring chunks' passes",
             "nccl": "chunked NCCL all-to-all"}
    used = getattr(prob, "exchange", "nccl") if P2P else "nccl"
    import fourierflows_jl_b200 as ff
    tuned = ff.dist.AUTOTUNE_LOG[:1] if EXCHANGE == "auto" else []   # first entry: the problem's plan
    return {"used": used, "how": names[used], "requested": EXCHANGE, "autotune_ms_fwd_plus_inv": tuned[0]["ms"] if tuned else None}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="auto", choices=["auto", "c3", "c5"])
    ap.add_argument("--n", type=int, default=8192, help="C3 grid size per side")
    ap.add_argument("--n3", type=int, default=2048, help="C5 grid size in x and y")
    ap.add_argument("--nz-per-gpu", type=int, default=256, help="C5 z-planes per GPU (weak scaling; 256 x 8 = 2048)")
    ap.add_argument("--cpu-sample", type=int, default=0, help="grid size of the cpu_baseline sample (default 4096 for C3)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
