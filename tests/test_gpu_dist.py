"""Slab-decomposed transforms on the device.  With one visible GPU the distributed code path runs with P = 1
(segmented strides, chunked exchange with the self-copy); with >= 2 GPUs a torchrun worker checks P = 2 (and 4, 8)."""
import os
import subprocess
import sys

import numpy as np
import pytest

import oracle as fo
from util import relerr

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def ff():
    import fourierflows_jl_b200 as ff
    assert ff.have_device()
    return ff


def ngpus():
    import ctypes as C
    import fourierflows_jl_b200 as ff
    n = C.c_int(0)
    ff._lib.load().ffb_device_count(C.byref(n))
    return n.value


@pytest.mark.parametrize("T,tol", [(np.float64, 1e-12), (np.float32, 1e-5)])
@pytest.mark.parametrize("nch", [1, 4])
def test_single_rank_slab_plan_matches_oracle(ff, T, tol, nch):
    comm = ff.Dist(0, 1, ff.Dist.unique_id())
    shape = (64, 32, 128)
    rng = np.random.default_rng(3)
    x = np.asfortranarray(rng.standard_normal(shape).astype(T))
    plan = ff.DistPlan(shape, T, comm, nchunks=nch)
    assert "slab" in plan.describe()
    xh = plan * ff.DevArray.from_numpy(x)
    assert relerr(xh.to_numpy(), fo.RfftPlan(shape, T) * x.astype(np.float64)) <= tol
    assert relerr(plan.solve(xh).to_numpy(), x) <= tol
    cp = ff.CProblem((32, 32, 32), 2 * np.pi, stepper="FilteredRK4", dt=1e-3, calcN="burgers3d", nu=1e-3, T=T, dist=comm)
    ob = fo.Burgers3D.Problem(nx=32, kappa=1e-3, dt=1e-3, stepper="FilteredRK4", T=T)
    c0 = fo.random_phase_field((32, 32, 32), 2 * np.pi, 4.0, slope=0, T=T)
    cp.set_physical(c0)
    ob.grid.rfftplan.mul(ob.sol, c0)
    cp.stepforward(2)
    fo.stepforward(ob, 2)
    assert relerr(cp.sol.to_numpy(), ob.sol) <= 2 * tol
    del cp, plan
    comm.close()


@pytest.mark.parametrize("T,tol", [(np.float64, 1e-12), (np.float32, 1e-5)])
@pytest.mark.parametrize("shape", [(64, 32), (256, 4096), (128, 8192)], ids=lambda s: "x".join(map(str, s)))
def test_single_rank_2d_slab_plan_matches_oracle(ff, T, tol, shape):
    """2-D slab decomposition with P = 1 (the same kernels: block-segmented half spectrum, padded (kb + 1, ny) slab; single-pass and
    four-step y lines): transforms and the fused vorticity problem against the oracle"""
    comm = ff.Dist(0, 1, ff.Dist.unique_id())
    rng = np.random.default_rng(4)
    x = np.asfortranarray(rng.standard_normal(shape).astype(T))
    plan = ff.DistPlan(shape, T, comm)
    assert "slab2d" in plan.describe()
    xh = plan * ff.DevArray.from_numpy(x)
    ref = fo.RfftPlan(shape, T) * x.astype(np.float64)
    assert relerr(xh.to_numpy(), ff.spectral_slab_2d(ref, 1, 0)) <= tol
    assert relerr(plan.solve(xh).to_numpy(), x) <= tol
    for fused in (0, 1):
        n = shape[0]
        cp = ff.CProblem((n, n), 2 * np.pi, stepper="ETDRK4", dt=2e-3, calcN="vorticity2d", nu=1e-3, T=T, dist=comm, fused=fused)
        ob = fo.TwoDNavierStokes.Problem(nx=n, nu=1e-3, dt=2e-3, stepper="ETDRK4", T=T)
        z0 = fo.random_phase_field((n, n), 2 * np.pi, 8.0, slope=-1, seed=1234, T=T)
        cp.set_physical(z0)
        ob.grid.rfftplan.mul(ob.sol, z0)
        cp.stepforward(2)
        fo.stepforward(ob, 2)
        assert relerr(cp.sol.to_numpy(), ff.spectral_slab_2d(ob.sol, 1, 0)) <= 2 * tol
        cp.close()
    del plan
    comm.close()


def test_unsupported_decompositions_fail_loudly(ff):
    comm = ff.Dist(0, 1, ff.Dist.unique_id())
    with pytest.raises(ff.FFBError):
        ff.DistPlan((64, 30, 64), np.float64, comm)      # non power-of-two y
    comm.close()


@pytest.mark.parametrize("P", [1, 2, 4, 8])
def test_multi_gpu_worker(P):
    """torchrun worker: every exchange (NCCL, copy-engine, peer-store, autotuned) and the fused slab-decomposed Burgers problem against the
    oracle.  P = 1 runs on the single-GPU box too (the rank is its own peer: same kernels, same blocked layouts, IPC-free)."""
    if ngpus() < P:
        pytest.skip(f"needs {P} GPUs")
    env = dict(os.environ)
    port = 29500 + P
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={P}", "--master-addr", "127.0.0.1",
                          "--master-port", str(port), os.path.join(ROOT, "tests", "dist_worker.py")], capture_output=True, text=True, env=env, timeout=900)
    assert res.returncode == 0 and "DIST PARITY OK" in res.stdout, res.stdout[-3000:] + res.stderr[-3000:]
