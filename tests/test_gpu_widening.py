"""SURVEY 8f rows beyond the core path: the closed elementwise vocabulary incl. `jacobianh` and multi-field arrays (8f-2) and the
asynchronous output path (8f-3), on the device, against the oracle / the reference's analytic known answers."""
import os

import numpy as np
import pytest

import oracle as fo
from util import relerr

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ff():
    import fourierflows_jl_b200 as ff
    assert ff.have_device()
    return ff


def test_jacobian_kats_on_device(ff):
    """test/runtests.jl:243-280: J(a, a) = 0, J(sin1, sin2), J(exp1, exp2) on TwoDGrid(64 x 128, 2pi x 3pi); atol = nx*ny*10*eps"""
    nx, ny, Lx, Ly = 64, 128, 2 * np.pi, 3 * np.pi
    g = ff.TwoDGrid(ff.GPU(), nx=nx, Lx=Lx, ny=ny, Ly=Ly)
    og = fo.TwoDGrid(nx=nx, Lx=Lx, ny=ny, Ly=Ly)
    x, y = np.asarray(og.x).reshape(-1, 1), np.asarray(og.y).reshape(1, -1)
    k0, l0 = 2 * np.pi / Lx, 2 * np.pi / Ly
    k1, l1, k2, l2 = 2 * k0, 6 * l0, 3 * k0, -3 * l0
    s1, s2 = np.asfortranarray(np.sin(k1 * x + l1 * y)), np.asfortranarray(np.sin(k2 * x + l2 * y))
    e1, e2 = np.asfortranarray(np.exp(1j * (k1 * x + l1 * y))), np.asfortranarray(np.exp(1j * (k2 * x + l2 * y)))
    atol = nx * ny * 10 * np.finfo(np.float64).eps
    d = ff.DevArray.from_numpy
    assert np.linalg.norm(ff.jacobian(d(s1), d(s1), g).to_numpy()) <= atol
    assert np.linalg.norm(ff.jacobian(d(s1), d(s2), g).to_numpy() - (k1 * l2 - k2 * l1) * np.cos(k1 * x + l1 * y) * np.cos(k2 * x + l2 * y)) <= atol
    assert np.linalg.norm(ff.jacobian(d(e1), d(e2), g).to_numpy() - (k2 * l1 - k1 * l2) * np.exp(1j * ((k1 + k2) * x + (l1 + l2) * y))) <= atol
    # random fields against the oracle, both element types and both precisions, also at a four-step size
    for shape, T, tol in (((64, 128), np.float64, 1e-12), ((128, 64), np.float32, 1e-5), ((128, 4096), np.float64, 1e-12)):
        gg = ff.TwoDGrid(ff.GPU(), nx=shape[0], Lx=2 * np.pi, ny=shape[1], Ly=3.0, T=T)
        og2 = fo.TwoDGrid(nx=shape[0], Lx=2 * np.pi, ny=shape[1], Ly=3.0, T=T)
        rng = np.random.default_rng(8)
        a, b = (np.asfortranarray(rng.standard_normal(shape).astype(T)) for _ in range(2))
        assert relerr(ff.jacobianh(d(a), d(b), gg).to_numpy(), fo.jacobianh(a, b, og2)) <= 10 * tol
        cT = np.complex64 if T == np.float32 else np.complex128
        ac, bc = (np.asfortranarray((rng.standard_normal(shape) + 1j * rng.standard_normal(shape)).astype(cT)) for _ in range(2))
        assert relerr(ff.jacobianh(d(ac), d(bc), gg).to_numpy(), fo.jacobianh(ac, bc, og2)) <= 10 * tol


@pytest.mark.parametrize("T,tol", [(np.float64, 1e-14), (np.float32, 1e-6)])
def test_multi_field_spectral_mul_and_dealias(ff, T, tol):
    """trailing field dimension (examples/OneDShallowWaterGeostrophicAdjustment.jl:114-126: `sol[:, 1..3]`): one call of the
    vocabulary kernels sweeps every field; `dealias!` zeroes `fh[kalias, :]` of all of them (src/domains.jl:471-473)"""
    nx, ny, nf = 64, 48, 3
    g = ff.TwoDGrid(ff.GPU(), nx=nx, Lx=2 * np.pi, ny=ny, Ly=2.0, T=T)
    og = fo.TwoDGrid(nx=nx, Lx=2 * np.pi, ny=ny, Ly=2.0, T=T)
    rng = np.random.default_rng(9)
    cT = np.complex64 if T == np.float32 else np.complex128
    sh = (g.nkr, ny, nf)
    a = np.asfortranarray((rng.standard_normal(sh) + 1j * rng.standard_normal(sh)).astype(cT))
    out = ff.DevArray(sh, cT)
    ff.spectral_mul(out, ff.DevArray.from_numpy(a), g, coef=-2j, px=1, py=2, w=g.invKrsq, dealias=True)
    ref = np.asfortranarray(((((-2j * og.kr) * og.l) * og.l) * og.invKrsq)[..., None] * a)
    for f in range(nf):
        r = np.asfortranarray(ref[..., f])
        fo.dealias(r, og)
        ref[..., f] = r
    got = out.to_numpy()
    assert relerr(got, ref) <= tol and np.array_equal(got == 0, ref == 0)
    # 1-D grid, three fields (the shallow-water example's layout): u_h, v_h, eta_h in one array
    g1 = ff.OneDGrid(ff.GPU(), nx=128, Lx=2 * np.pi, T=T)
    og1 = fo.OneDGrid(nx=128, Lx=2 * np.pi, T=T)
    s = np.asfortranarray((rng.standard_normal((g1.nkr, 3)) + 1j * rng.standard_normal((g1.nkr, 3))).astype(cT))
    o1 = ff.DevArray(s.shape, cT)
    ff.spectral_mul(o1, ff.DevArray.from_numpy(s), g1, coef=1j, px=1)
    assert relerr(o1.to_numpy(), (1j * og1.kr)[:, None] * s) <= tol
    ds = ff.DevArray.from_numpy(s)
    ff.dealias(ds, g1)
    r = s.copy(order="F")
    fo.dealias(r, og1)
    assert np.array_equal(ds.to_numpy(), r)


def test_async_snapshots_do_not_see_later_steps(ff, tmp_path):
    """`saveoutput` through the snapshot ring: the copy is stream-ordered with the steps (it holds the state of the step it was
    taken at even though stepping continues before it is waited for), and the files hold what a blocking download would"""
    n = 256
    cp = ff.CProblem((n, n), 2 * np.pi, stepper="RK4", dt=2e-3, calcN="vorticity2d", nu=1e-3)
    z0 = fo.random_phase_field((n, n), 2 * np.pi, 8.0, slope=-1, seed=1234)
    cp.set_physical(z0)
    snap = ff.AsyncSnapshot(cp.sol.nbytes, nbuf=3)
    want, slots = [], []
    for k in range(3):
        cp.stepforward(2)
        slots.append(snap.begin(cp.sol))          # non-blocking
        want.append(None)
        cp.stepforward(1)                         # more steps are enqueued behind the staging copy
    with pytest.raises(ff.FFBError):
        snap.begin(cp.sol)                        # ring full: slots must be released first
    # reference run with blocking downloads at the same steps
    cq = ff.CProblem((n, n), 2 * np.pi, stepper="RK4", dt=2e-3, calcN="vorticity2d", nu=1e-3)
    cq.set_physical(z0)
    for k in range(3):
        cq.stepforward(2)
        want[k] = cq.sol.to_numpy()
        cq.stepforward(1)
    for k, s in enumerate(slots):
        got = np.array(snap.wait(s, cp.sol.shape, cp.sol.dtype))
        assert np.array_equal(got, want[k])
        snap.release(s)
    assert snap.begin(cp.sol) in (0, 1, 2)
    snap.close()
    # Output / saveoutput mirror
    prob = ff.Diffusion.Problem(ff.GPU(), nx=64, Lx=2 * np.pi, kappa=0.01, dt=1e-3, stepper="RK4")
    ff.Diffusion.set_c(prob, np.cos(np.asarray(prob.grid.x)))
    out = ff.Output(prob, str(tmp_path / "run.npz"), {"sol": lambda p: p.sol})
    ff.stepforward(prob, 5)
    ff.saveoutput(out)
    held = prob.sol.to_numpy()
    ff.stepforward(prob, 5)
    out.close()
    f = np.load(str(tmp_path / "run_snapshot_5.npz"))
    assert np.array_equal(f["snapshots/sol/5"], held) and float(f["snapshots/t/5"]) == float(5 * 1e-3) or np.isclose(float(f["snapshots/t/5"]), 5e-3)


@pytest.mark.parametrize("kw", [dict(n=(256, 256), stepper="ETDRK4", calcN="vorticity2d", nu=1e-3, T=np.float64, fused=1),
                                dict(n=(128, 64), stepper="FilteredRK4", calcN="vorticity2d", nu=1e-3, T=np.float32, fused=0),
                                dict(n=(32, 32, 32), stepper="LSRK54", calcN="burgers3d", nu=1e-2, T=np.float32, fused=1)],
                         ids=["etdrk4-2d-f64", "filtered-rk4-2d-f32", "lsrk54-3d-f32"])
def test_host_pipeline_equals_blocking_form(ff, kw):
    """ffb_pipeline_*: independent host states (pinned), uploaded / stepped / downloaded with the copies of neighbouring submissions
    overlapping the steps, must equal h2d + ffb_step + d2h bit for bit -- more submissions than slots (the ring wraps), nsteps = 2,
    outputs read only after their ticket.  AB3 keeps history across steps: FFB_EUNSUPPORTED."""
    kw = dict(kw)
    n = kw.pop("n")
    prob = ff.CProblem(n, 2 * np.pi, dt=1e-3, **kw)
    cT = prob.sol.dtype
    rng = np.random.default_rng(5)
    nsub, depth, nsteps = 5, 2, 2
    states = []
    for _ in range(nsub):
        prob.set_physical(np.asfortranarray(rng.standard_normal(prob.physical_shape).astype(prob.T)))
        states.append(prob.sol.to_numpy().copy())
    # blocking form
    expect = []
    for s in states:
        prob.sol.copy_from_host(s)
        prob.stepforward(nsteps)
        expect.append(prob.sol.to_numpy().copy())
    ins = [ff.PinnedBuffer(prob.spectral_shape, cT) for _ in range(nsub)]
    outs = [ff.PinnedBuffer(prob.spectral_shape, cT) for _ in range(nsub)]
    for b, s in zip(ins, states):
        b.array[...] = s
    pipe = prob.pipeline(depth)
    tickets = [pipe.submit(ins[i], outs[i], nsteps) for i in range(nsub)]
    assert tickets == [i % depth for i in range(nsub)]
    pipe.wait_all()
    for i in range(nsub):
        assert np.isfinite(outs[i].array).all()
        assert np.array_equal(outs[i].array, expect[i]), f"submission {i}"
    # a second round on the same pipeline, nsteps = 0: the state comes back unchanged
    t = pipe.submit(ins[0], outs[1], 0)
    pipe.wait(t)
    assert np.array_equal(outs[1].array, states[0])
    with pytest.raises(ValueError):
        pipe.submit(ff.PinnedBuffer((4,), cT), outs[0], 1)
    pipe.close()
    for b in ins + outs:
        b.close()
    prob.close()
    ab3 = ff.CProblem((64, 64), 2 * np.pi, stepper="AB3", dt=1e-3, calcN="vorticity2d", nu=1e-3)
    with pytest.raises(ff.FFBError) as ei:
        ab3.pipeline(2)
    assert ei.value.code == ff._lib.FFB_EUNSUPPORTED
    ab3.close()
