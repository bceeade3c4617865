"""Worker run under torchrun (one process per GPU): slab-decomposed FFT and Burgers problem against the CPU oracle.
Exit code 0 = parity green on every rank.  Usage: torchrun --nproc-per-node P tests/dist_worker.py [quick]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    import fourierflows_jl_b200 as ff
    from fourierflows_jl_b200 import _lib as L
    import oracle as fo
    from util import relerr

    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    L.call("ffb_set_device", local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, P = dist.get_rank(), dist.get_world_size()
    comm = ff.Dist.from_torch()
    worst = 0.0
    for shape, T, tol in (((64, 32, 64), np.float64, 1e-12), ((32, 64, 16 * P), np.float32, 1e-5), ((256, 256, 256), np.float64, 1e-12)):
        rng = np.random.default_rng(5)
        x = np.asfortranarray(rng.standard_normal(shape).astype(T))
        ref = fo.RfftPlan(shape, T) * x.astype(np.float64)
        for nch in (0, 1, -1, -2, -3):   # -1: fused pass + collective (peer stores over NVLink), -2: copy-engine pushes, -3: autotuned
            plan = ff.DistPlan(shape, T, comm, nchunks=max(nch, 0))
            if nch < 0:
                plan.enable_p2p({-1: "peer-store", -2: "copy-engine", -3: "auto"}[nch])
            xl = ff.DevArray.from_numpy(ff.physical_slab(x, P, rank))
            xh = plan * xl
            e1 = relerr(xh.to_numpy(), ff.spectral_slab(ref, P, rank))
            back = plan.solve(xh)
            e2 = relerr(back.to_numpy(), ff.physical_slab(x, P, rank))
            for _ in range(3):   # repeated transforms exercise the double-buffered receive buffers
                xh = plan * xl
                back = plan.solve(xh)
            e2 = max(e2, relerr(back.to_numpy(), ff.physical_slab(x, P, rank)), relerr(xh.to_numpy(), ff.spectral_slab(ref, P, rank)))
            e3 = 0.0 if np.array_equal(xl.to_numpy(), ff.physical_slab(x, P, rank)) else 1.0
            worst = max(worst, e1 / tol, e2 / tol, e3)
            if rank == 0:
                print(f"dist fft {shape} {np.dtype(T).name} chunks={nch}: fwd {e1:.2e} rt {e2:.2e} [{plan.describe()}]", flush=True)
    # 2-D slab decomposition (physical y-slabs <-> spectral kx blocks): transforms and the 2-D vorticity problem (config C3's equation)
    for shape, T, tol in (((64, 32), np.float64, 1e-12), ((256, 4096), np.float32, 1e-5), ((128, 8192), np.float64, 1e-12)):
        rng = np.random.default_rng(6)
        x = np.asfortranarray(rng.standard_normal(shape).astype(T))
        ref = fo.RfftPlan(shape, T) * x.astype(np.float64)
        plan = ff.DistPlan(shape, T, comm)
        xl = ff.DevArray.from_numpy(ff.physical_slab_2d(x, P, rank))
        xh = plan * xl
        back = plan.solve(xh)
        d = xh.to_numpy() - ff.spectral_slab_2d(ref, P, rank)
        acc = torch.tensor([float(np.sum(np.abs(d) ** 2)), float(np.sum(np.abs(ff.spectral_slab_2d(ref, P, rank)) ** 2))], device="cuda", dtype=torch.float64)
        dist.all_reduce(acc)
        e1 = float(torch.sqrt(acc[0] / acc[1]).item())
        e2 = relerr(back.to_numpy(), ff.physical_slab_2d(x, P, rank))
        worst = max(worst, e1 / tol, e2 / tol)
        if rank == 0:
            print(f"dist fft2d {shape} {np.dtype(T).name}: fwd {e1:.2e} rt {e2:.2e} [{plan.describe()}]", flush=True)
    for stepper, T, tol, fused in (("ETDRK4", np.float64, 1e-12, 1), ("FilteredRK4", np.float32, 1e-5, 0), ("LSRK54", np.float64, 1e-12, 1)):
        n = 128
        ob = fo.TwoDNavierStokes.Problem(nx=n, nu=1e-3, dt=2e-3, stepper=stepper, T=T)
        z0 = fo.random_phase_field((n, n), 2 * np.pi, 8.0, slope=-1, seed=1234, T=T)
        ob.grid.rfftplan.mul(ob.sol, z0)
        cp = ff.CProblem((n, n), 2 * np.pi, stepper=stepper, dt=2e-3, calcN="vorticity2d", nu=1e-3, T=T, dist=comm, fused=fused)
        cp.set_physical(ff.physical_slab_2d(z0, P, rank))
        for s in range(3):
            cp.stepforward(1)
            fo.stepforward(ob, 1)
            ref = ff.spectral_slab_2d(ob.sol, P, rank)
            acc = torch.tensor([float(np.sum(np.abs(cp.sol.to_numpy() - ref) ** 2)), float(np.sum(np.abs(ref) ** 2))], device="cuda", dtype=torch.float64)
            dist.all_reduce(acc)
            e = float(torch.sqrt(acc[0] / acc[1]).item())
            worst = max(worst, e / ((s + 1) * tol))
        if rank == 0:
            print(f"dist vorticity2d {stepper} {np.dtype(T).name} fused={fused}: rel-L2 after 3 steps {e:.2e}", flush=True)
        cp.close()
    # slab-decomposed Burgers problem (configs C4 / C5 shape), C-driven, vs the single-process oracle
    # fused = 1: square folded into the x pass, -1/2 im kr and the dealias mask into the last z pass (NCCL and peer-store exchanges;
    # the copy-engine exchange runs the unfused kernels)
    for stepper, T, tol, exch, fused in (("FilteredRK4", np.float64, 1e-12, "copy-engine", 0), ("ETDRK4", np.float32, 1e-5, "peer-store", 0),
                                         ("LSRK54", np.float32, 1e-5, None, 0), ("ETDRK4", np.float32, 1e-5, "peer-store", 1),
                                         ("FilteredRK4", np.float64, 1e-12, "peer-store", 1), ("LSRK54", np.float64, 1e-12, None, 1),
                                         ("RK4", np.float64, 1e-12, "copy-engine", 1)):
        n = (64, 64, 64)
        ob = fo.Burgers3D.Problem(nx=64, kappa=1e-3, dt=1e-3, stepper=stepper, T=T)
        c0 = fo.random_phase_field(n, 2 * np.pi, 4.0, slope=0, seed=1234, T=T)
        ob.grid.rfftplan.mul(ob.sol, c0)
        cp = ff.CProblem(n, 2 * np.pi, stepper=stepper, dt=1e-3, calcN="burgers3d", nu=1e-3, T=T, dist=comm, fused=fused)
        if exch is not None:
            cp.enable_p2p(exch)
        cp.set_physical(ff.physical_slab(c0, P, rank))
        for s in range(3):
            cp.stepforward(1)
            fo.stepforward(ob, 1)
            # relative L2 error of the GLOBAL state: slabs that hold only high wavenumbers are ~1e-20 and have no meaningful
            # relative error of their own
            ref = ff.spectral_slab(ob.sol, P, rank)
            acc = torch.tensor([float(np.sum(np.abs(cp.sol.to_numpy() - ref) ** 2)), float(np.sum(np.abs(ref) ** 2))], device="cuda", dtype=torch.float64)
            dist.all_reduce(acc)
            e = float(torch.sqrt(acc[0] / acc[1]).item())
            worst = max(worst, e / ((s + 1) * tol))
        if rank == 0:
            print(f"dist burgers {stepper} {np.dtype(T).name} exchange={exch or 'nccl'} fused={fused}: rel-L2 after 3 steps {e:.2e}", flush=True)
    t = torch.tensor([worst], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ok = float(t.item()) <= 1.0
    if rank == 0:
        print("DIST PARITY", "OK" if ok else f"FAILED (worst ratio {float(t.item()):.2f})", flush=True)
    comm.close()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
