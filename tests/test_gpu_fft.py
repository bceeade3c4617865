"""FFT parity through the C ABI (`ffb_plan_create` / `ffb_fft_forward` / `ffb_fft_inverse`).

1. the reference's closed-form known-answer tests (test/test_fft.jl, test/test_ifft.jl, test/createffttestfunctions.jl)
   at rtol 1e-12, sizes 1-D 32, 2-D 32x64, 3-D 32x30x16;
2. parity with the CPU oracle (scipy.fft) on seeded random fields: rel-L2 <= 1e-12 (Float64), <= 1e-5 (Float32).
"""
import numpy as np
import pytest

import oracle as fo
from util import TOL_FFT, isapprox, relerr

pytestmark = pytest.mark.gpu
rtol_fft = 1e-12  # test/runtests.jl:21


@pytest.fixture(scope="module")
def ff():
    import fourierflows_jl_b200 as ff
    assert ff.have_device(), "CUDA extension must be present on the GPU box (no fallback)"
    return ff


def dev(ff, a):
    return ff.DevArray.from_numpy(np.asfortranarray(a))


def test_kat_1d_cosmx(ff):
    g = ff.OneDGrid(ff.GPU(), nx=32, Lx=2 * np.pi)
    og = fo.OneDGrid(nx=32, Lx=2 * np.pi)
    m, phi = 5, np.pi / 3
    k0 = og.k[1]
    assert np.array_equal(g.k.to_numpy(), og.k) and np.array_equal(g.kr.to_numpy(), og.kr)
    f1 = np.cos(m * k0 * og.x + phi)
    th = np.zeros(g.nk, dtype=complex)
    thr = np.zeros(g.nkr, dtype=complex)
    for i in range(g.nk):
        if abs(og.k[i]) == m * k0:
            th[i] = -np.exp(np.sign(og.k[i]) * 1j * phi) * g.nx / 2
    for i in range(g.nkr):
        if abs(og.k[i]) == m * k0:
            thr[i] = -np.exp(np.sign(og.kr[i]) * 1j * phi) * g.nx / 2
    f1h = (g.fftplan * dev(ff, f1.astype(complex))).to_numpy()
    df1 = dev(ff, f1)
    f1hr = (g.rfftplan * df1).to_numpy()
    out = ff.DevArray.zeros(np.complex128, (g.nkr,))
    ff.mul_(out, g.rfftplan, df1)                         # mul!(f1hr_mul, g.rfftplan, f1)
    assert isapprox(f1h, th, rtol_fft)                    # test_fft_cosmx
    assert isapprox(f1hr, thr, rtol_fft)                  # test_rfft_cosmx
    assert isapprox(out.to_numpy(), thr, rtol_fft)        # test_rfft_mul_cosmx
    assert isapprox(f1, g.fftplan.solve(dev(ff, f1h)).to_numpy().real, rtol_fft)   # test_ifft_cosmx
    f1b = ff.DevArray.zeros(np.float64, (g.nx,))
    ff.ldiv_(f1b, g.rfftplan, dev(ff, f1hr))              # test_irfft_mul_cosmx
    assert isapprox(f1, f1b.to_numpy(), rtol_fft)


def test_kat_2d(ff):
    g = ff.TwoDGrid(ff.GPU(), nx=32, Lx=2 * np.pi, ny=64, Ly=3 * np.pi)
    og = fo.TwoDGrid(nx=32, Lx=2 * np.pi, ny=64, Ly=3 * np.pi)
    x, y = og.x.reshape(-1, 1), og.y.reshape(1, -1)
    m, n = 5, 2
    k0, l0 = og.k[1, 0], og.l[0, 1]
    f1 = np.cos(m * k0 * x) * np.cos(n * l0 * y)
    f2 = np.sin(m * k0 * x + n * l0 * y)
    sh, shr = (g.nk, g.nl), (g.nkr, g.nl)
    K, Lw = np.broadcast_to(og.k, sh), np.broadcast_to(og.l, sh)
    Kr, Lr = np.broadcast_to(og.kr, shr), np.broadcast_to(og.l, shr)
    f1h_th = np.where((np.abs(K) == m * k0) & (np.abs(Lw) == n * l0), -g.nx * g.ny / 4, 0).astype(complex)
    f2h_th = -1j * (np.where((K == m * k0) & (Lw == n * l0), -g.nx * g.ny / 2, 0) + np.where((K == -m * k0) & (Lw == -n * l0), g.nx * g.ny / 2, 0))
    f1hr_th = np.where((np.abs(Kr) == m * k0) & (np.abs(Lr) == n * l0), -g.nx * g.ny / 4, 0).astype(complex)
    f2hr_th = -1j * np.where((Kr == m * k0) & (Lr == n * l0), -g.nx * g.ny / 2, 0)
    assert isapprox((g.fftplan * dev(ff, f1.astype(complex))).to_numpy(), f1h_th, rtol_fft)
    assert isapprox((g.fftplan * dev(ff, f2.astype(complex))).to_numpy(), f2h_th, rtol_fft)
    for f, th in ((f1, f1hr_th), (f2, f2hr_th)):
        fh = g.rfftplan * dev(ff, f)
        assert isapprox(fh.to_numpy(), th, rtol_fft)
        fb = ff.DevArray.zeros(np.float64, (g.nx, g.ny))
        g.rfftplan.ldiv(fb, fh)
        assert isapprox(f, fb.to_numpy(), rtol_fft)
        assert isapprox(f, g.fftplan.solve(g.fftplan * dev(ff, f.astype(complex))).to_numpy().real, rtol_fft)


def test_kat_3d_32x30x16(ff):
    kw = dict(nx=32, Lx=2 * np.pi, ny=30, Ly=3 * np.pi, nz=16, Lz=4.0)
    g = ff.ThreeDGrid(ff.GPU(), **kw)
    og = fo.ThreeDGrid(**kw)
    x, y, z = og.x.reshape(-1, 1, 1), og.y.reshape(1, -1, 1), og.z.reshape(1, 1, -1)
    mx, my, mz = 5, 2, 3
    k0, l0, m0 = og.k[1, 0, 0], og.l[0, 1, 0], og.m[0, 0, 1]
    f1 = np.cos(mx * k0 * x) * np.cos(my * l0 * y) * np.cos(mz * m0 * z)
    f2 = np.sin(mx * k0 * x + my * l0 * y + mz * m0 * z)
    shr = (g.nkr, g.nl, g.nm)
    Kr, Lr, Mr = np.broadcast_to(og.kr, shr), np.broadcast_to(og.l, shr), np.broadcast_to(og.m, shr)
    N3 = g.nx * g.ny * g.nz
    f1hr_th = np.where((np.abs(Kr) == mx * k0) & (np.abs(Lr) == my * l0) & (np.abs(Mr) == mz * m0), N3 / 8, 0).astype(complex)
    f2hr_th = -1j * np.where((Kr == mx * k0) & (Lr == my * l0) & (Mr == mz * m0), N3 / 2, 0)
    for f, th in ((f1, f1hr_th), (f2, f2hr_th)):
        fh = g.rfftplan * dev(ff, f)
        assert isapprox(fh.to_numpy(), th, rtol_fft)
        assert isapprox(f, g.rfftplan.solve(fh).to_numpy(), rtol_fft)
    sh = (g.nk, g.nl, g.nm)
    K, Lw, M = np.broadcast_to(og.k, sh), np.broadcast_to(og.l, sh), np.broadcast_to(og.m, sh)
    f1h_th = np.where((np.abs(K) == mx * k0) & (np.abs(Lw) == my * l0) & (np.abs(M) == mz * m0), N3 / 8, 0).astype(complex)
    f1h = g.fftplan * dev(ff, f1.astype(complex))
    assert isapprox(f1h.to_numpy(), f1h_th, rtol_fft)
    assert isapprox(f1, g.fftplan.solve(f1h).to_numpy().real, rtol_fft)


SHAPES = [(4,), (6,), (16,), (30,), (34,), (256,), (1024,), (8192,), (16384,), (32768,), (6, 8), (32, 64), (34, 16), (512, 256), (64, 4096),
          (2048, 32), (6, 8, 10), (32, 30, 16), (64, 32, 128), (256, 16, 16)]


@pytest.mark.parametrize("T", [np.float64, np.float32])
@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: "x".join(map(str, s)))
def test_rfft_matches_oracle(ff, shape, T):
    """ragged / non-power-of-two / maximum in-kernel sizes, both precisions; the input of the inverse is preserved"""
    rng = np.random.default_rng(7)
    x = np.asfortranarray(rng.standard_normal(shape).astype(T))
    oplan = fo.RfftPlan(shape, T)
    ref = oplan * x.astype(np.float64)
    plan = ff.Plan(shape, T, ff._lib.FFB_R2C)
    dx = dev(ff, x)
    xh = plan * dx
    tol = TOL_FFT[np.dtype(T)]
    assert relerr(xh.to_numpy(), ref) <= tol
    assert np.array_equal(dx.to_numpy(), x)
    before = xh.to_numpy()
    back = plan.solve(xh)
    assert relerr(back.to_numpy(), x) <= tol
    assert np.array_equal(xh.to_numpy(), before), "ldiv! must preserve its input in this implementation"


@pytest.mark.parametrize("T", [np.float64, np.float32])
@pytest.mark.parametrize("shape", [(2,), (30,), (64,), (8192,), (16384,), (16, 30), (128, 64), (8, 8, 8), (32, 30, 16)], ids=lambda s: "x".join(map(str, s)))
def test_fft_c2c_matches_oracle(ff, shape, T):
    rng = np.random.default_rng(8)
    cT = np.complex64 if T == np.float32 else np.complex128
    x = np.asfortranarray((rng.standard_normal(shape) + 1j * rng.standard_normal(shape)).astype(cT))
    ref = fo.FftPlan(shape, T) * x.astype(np.complex128)
    plan = ff.Plan(shape, T, ff._lib.FFB_C2C)
    xh = plan * dev(ff, x)
    tol = TOL_FFT[np.dtype(T)]
    assert relerr(xh.to_numpy(), ref) <= tol
    assert relerr(plan.solve(xh).to_numpy(), x) <= tol
    # in place
    d = dev(ff, x)
    plan.mul(d, d)
    assert relerr(d.to_numpy(), ref) <= tol


def test_generic_and_register_paths_agree(ff):
    """the arbitrary-size path is an independent implementation: both must agree on power-of-two sizes"""
    rng = np.random.default_rng(9)
    x = np.asfortranarray(rng.standard_normal((256, 128)))
    a = ff.Plan((256, 128), np.float64, ff._lib.FFB_R2C)
    b = ff.Plan((256, 128), np.float64, ff._lib.FFB_R2C, flags=ff._lib.FFB_PLAN_FORCE_GENERIC)
    assert "pow2" in a.describe() and "pow2" not in b.describe()
    assert relerr((a * dev(ff, x)).to_numpy(), (b * dev(ff, x)).to_numpy()) <= 1e-14


def test_batched_fields(ff):
    """trailing field dimension (examples/OneDShallowWaterGeostrophicAdjustment.jl:114-126 shape requirement)"""
    rng = np.random.default_rng(10)
    x = np.asfortranarray(rng.standard_normal((64, 32, 3)))
    plan = ff.Plan((64, 32), np.float64, ff._lib.FFB_R2C, nbatch=3)
    xh = (plan * dev(ff, x)).to_numpy()
    op = fo.RfftPlan((64, 32), np.float64)
    for f in range(3):
        assert relerr(xh[:, :, f], op * x[:, :, f]) <= 1e-13


def test_odd_sizes_raise_domain_error(ff):
    """runtests.jl:133-138"""
    for shape in ((5,), (5, 4), (4, 5), (5, 4, 6), (4, 5, 6), (4, 6, 5)):
        with pytest.raises(ff.DomainError):
            ff.Plan(shape, np.float64, ff._lib.FFB_R2C)
    with pytest.raises(ff.DomainError):
        ff.OneDGrid(ff.GPU(), nx=5, Lx=1)
    with pytest.raises(ff.DomainError):
        ff.TwoDGrid(ff.GPU(), nx=4, Lx=1, ny=5, Ly=2)
    with pytest.raises(ff.DomainError):
        ff.ThreeDGrid(ff.GPU(), nx=4, Lx=1, ny=6, Ly=2, nz=5, Lz=3)


@pytest.mark.parametrize("shape", [(64,), (64, 64), (30, 16), (16, 8, 8)], ids=lambda s: "x".join(map(str, s)))
def test_c2r_ignores_non_hermitian_parts_like_fftw(ff, shape):
    """`ldiv!` on a spectrum that is not Hermitian-consistent (e.g. `-im*kr*...` at the Nyquist column, as every
    calcN! produces): FFTW, cuFFT and pocketfft ignore Im X[0] and Im X[nx/2] of each x-line; so must we."""
    rng = np.random.default_rng(21)
    sh = (shape[0] // 2 + 1,) + tuple(shape[1:])
    xh = np.asfortranarray(rng.standard_normal(sh) + 1j * rng.standard_normal(sh))
    ref = fo.RfftPlan(shape, np.float64).solve(xh)
    plan = ff.Plan(shape, np.float64, ff._lib.FFB_R2C)
    assert relerr(plan.solve(dev(ff, xh)).to_numpy(), ref) <= 1e-13
    gen = ff.Plan(shape, np.float64, ff._lib.FFB_R2C, flags=ff._lib.FFB_PLAN_FORCE_GENERIC)
    assert relerr(gen.solve(dev(ff, xh)).to_numpy(), ref) <= 1e-13


@pytest.mark.parametrize("T,tol", [(np.float64, 1e-12), (np.float32, 1e-5)])
@pytest.mark.parametrize("shape", [(64, 32), (128, 4096), (32, 16, 64)], ids=lambda s: "x".join(map(str, s)))
def test_fused_transforms_match_unfused_reference_expressions(ff, shape, T, tol):
    """ffb_fft_inverse_ex / ffb_fft_forward_ex (spectral factor, physical product, accumulate + dealias folded into the
    passes) against the same expressions evaluated by the oracle; covers single-pass and four-step strided passes."""
    nd = len(shape)
    G = {2: ff.TwoDGrid, 3: ff.ThreeDGrid}[nd]
    OG = {2: fo.TwoDGrid, 3: fo.ThreeDGrid}[nd]
    kw = dict(nx=shape[0], Lx=2 * np.pi, ny=shape[1], Ly=3 * np.pi, T=T)
    if nd == 3:
        kw.update(nz=shape[2], Lz=4.0)
    g, og = G(ff.GPU(), **kw), OG(**kw)
    rng = np.random.default_rng(31)
    sh = (g.nkr,) + tuple(shape[1:])
    cT = np.complex64 if T == np.float32 else np.complex128
    ah = np.asfortranarray((rng.standard_normal(sh) + 1j * rng.standard_normal(sh)).astype(cT))
    zeta = np.asfortranarray(rng.standard_normal(shape).astype(T))
    plan, oplan = g.rfftplan, og.rfftplan
    # inverse_ex: irfft(im * l * invKrsq .* ah) .* zeta
    ref = oplan.solve(((1j * og.l) * og.invKrsq) * ah) * zeta
    out = ff.DevArray(shape, T)
    plan.ldiv_ex(out, dev(ff, ah), coef=1j, l=g.l, w=g.invKrsq, mul=dev(ff, zeta))
    assert relerr(out.to_numpy(), ref) <= tol
    ref = oplan.solve(((-1j * og.kr) * og.invKrsq) * ah)
    plan.ldiv_ex(out, dev(ff, ah), coef=-1j, kx=g.kr, w=g.invKrsq)
    assert relerr(out.to_numpy(), ref) <= tol
    # forward_ex: dealias!(-im*kr .* acc - im*l .* rfft(zeta))  (and the m-vector / malias in 3-D)
    acc = np.asfortranarray((rng.standard_normal(sh) + 1j * rng.standard_normal(sh)).astype(cT))
    ref = (-1j * og.kr) * acc - (1j * og.l) * (oplan * zeta)
    ref = np.asfortranarray(ref)
    fo.dealias(ref, og)
    outh = ff.DevArray(sh, cT)
    alias = [g.kralias, g.lalias] + ([g.malias] if nd == 3 else [])
    plan.mul_ex(outh, dev(ff, zeta), coef=-1j, l=g.l, acc=dev(ff, acc), acoef=-1j, akx=g.kr, alias=alias)
    assert relerr(outh.to_numpy(), ref) <= tol
    assert np.array_equal(outh.to_numpy() == 0, ref == 0), "the dealiased box must be exactly zero"
    if nd == 3:
        ref = np.asfortranarray(((-0.5j * og.kr) * og.m) * (oplan * zeta))
        fo.dealias(ref, og)
        plan.mul_ex(outh, dev(ff, zeta), coef=-0.5j, kx=g.kr, m=g.m, alias=alias)
        assert relerr(outh.to_numpy(), ref) <= tol


@pytest.mark.parametrize("T,tol", [(np.float64, 1e-12), (np.float32, 1e-5)], ids=["f64", "f32"])
@pytest.mark.parametrize("shape", [(64, 32), (64, 4096), (128, 8192), (16, 8, 4096), (30, 48)], ids=lambda s: "x".join(map(str, s)))
def test_inverse_multi_shares_the_first_sub_pass_and_matches_separate_transforms(ff, shape, T, tol, monkeypatch):
    """ffb_fft_inverse_multi (zeta, u*zeta, v*zeta of the vorticity calcN! from one `sol`): against the oracle expressions, and
    bit-for-bit against separate ffb_fft_inverse_ex calls, for the shared four-step sub-pass (last dimension >= 4096, ragged column
    tiles, 3-D), the single-pass fallback, FFB_MULTI_A=0, and -- error code -- arbitrary sizes."""
    nd = len(shape)
    G = {2: ff.TwoDGrid, 3: ff.ThreeDGrid}[nd]
    OG = {2: fo.TwoDGrid, 3: fo.ThreeDGrid}[nd]
    kw = dict(nx=shape[0], Lx=2 * np.pi, ny=shape[1], Ly=3 * np.pi, T=T)
    if nd == 3:
        kw.update(nz=shape[2], Lz=4.0)
    g, og = G(ff.GPU(), **kw), OG(**kw)
    rng = np.random.default_rng(77)
    sh = (g.nkr,) + tuple(shape[1:])
    cT = np.complex64 if T == np.float32 else np.complex128
    ah = np.asfortranarray((rng.standard_normal(sh) + 1j * rng.standard_normal(sh)).astype(cT))
    plan, oplan = g.rfftplan, og.rfftplan
    dah = dev(ff, ah)
    outs = [ff.DevArray(shape, T) for _ in range(3)]
    variants = lambda: [dict(), dict(coef=1j, l=g.l, w=g.invKrsq, mul=outs[0]), dict(coef=-1j, kx=g.kr, w=g.invKrsq, mul=outs[0])]
    if any(n & (n - 1) for n in shape):
        with pytest.raises(ff.FFBError) as ei:
            plan.ldiv_multi(outs, dah, variants())
        assert ei.value.code == ff._lib.FFB_EUNSUPPORTED
        return
    zeta = oplan.solve(ah)
    refs = [zeta, oplan.solve(((1j * og.l) * og.invKrsq) * ah) * zeta, oplan.solve(((-1j * og.kr) * og.invKrsq) * ah) * zeta]
    got = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("FFB_MULTI_A", mode)
        for o in outs:
            o.fill_zero()
        plan.ldiv_multi(outs, dah, variants())
        got[mode] = [o.to_numpy() for o in outs]
        for a, r in zip(got[mode], refs):
            assert relerr(a, r) <= tol
        assert np.array_equal(dah.to_numpy(), ah), "the input is preserved"
    for a, b in zip(got["1"], got["0"]):
        assert np.array_equal(a, b)
    # separate calls
    sep = [ff.DevArray(shape, T) for _ in range(3)]
    plan.ldiv(sep[0], dah)
    plan.ldiv_ex(sep[1], dah, coef=1j, l=g.l, w=g.invKrsq, mul=sep[0])
    plan.ldiv_ex(sep[2], dah, coef=-1j, kx=g.kr, w=g.invKrsq, mul=sep[0])
    for a, b in zip(got["1"], sep):
        assert np.array_equal(a, b.to_numpy())
    # twice on one plan (scratch reuse), two variants only
    monkeypatch.setenv("FFB_MULTI_A", "1")
    plan.ldiv_multi(outs[:2], dah, variants()[:2])
    assert np.array_equal(outs[1].to_numpy(), got["1"][1])


@pytest.mark.parametrize("T,tol", [(np.float64, 1e-12), (np.float32, 1e-5)], ids=["f64", "f32"])
@pytest.mark.parametrize("shape", [(64, 1024), (66, 2048), (128, 4096), (32, 8192), (16, 16384), (32, 1024, 8), (16, 8, 2048), (8, 4096, 2)],
                         ids=lambda s: "x".join(map(str, s)))
def test_fused_l2_four_step_matches_two_kernel_form_and_oracle(ff, shape, T, tol, monkeypatch):
    """The persistent fused four-step pass (both sub-passes in one kernel, intermediate in an L2-resident ring; every
    instantiated N1 x N2 split, ragged column chunks, outer slices, in-place execution) against the oracle, and bit-for-bit
    against the two-kernel four-step it replaces (same arithmetic, different schedule)."""
    monkeypatch.setenv("FFB_FOURSTEP_MIN", "1024")
    if T == np.float64 and max(shape[1:]) > 8192:
        pytest.skip("Float64 register kernels stop at 8192; 16384 = 128 x 128 is covered in Float32")
    rng = np.random.default_rng(43)
    x = np.asfortranarray(rng.standard_normal(shape).astype(T))
    ref = fo.RfftPlan(shape, T) * x.astype(np.float64)
    outs = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("FFB_L2FOUR", mode)   # 1: one persistent kernel, intermediate in L2 (opt-in); 0: two kernels (default)
        plan = ff.Plan(shape, T, ff._lib.FFB_R2C)
        assert "four-step" in plan.describe()
        xh = plan * dev(ff, x)
        outs[mode] = xh.to_numpy()
        assert relerr(outs[mode], ref) <= tol
        back = plan.solve(xh)
        assert relerr(back.to_numpy(), x) <= tol
        assert np.array_equal(xh.to_numpy(), outs[mode]), "the inverse transform must preserve its input"
    assert np.array_equal(outs["1"], outs["0"])
    # twice in a row on one plan: the kernel leaves its ticket / completion counters zeroed
    monkeypatch.setenv("FFB_L2FOUR", "1")
    plan = ff.Plan(shape, T, ff._lib.FFB_R2C)
    a = (plan * dev(ff, x)).to_numpy()
    b = (plan * dev(ff, x)).to_numpy()
    assert np.array_equal(a, b) and np.array_equal(a, outs["1"])
    # complex plan, in place
    cT = np.complex64 if T == np.float32 else np.complex128
    z = np.asfortranarray((rng.standard_normal(shape) + 1j * rng.standard_normal(shape)).astype(cT))
    cplan = ff.Plan(shape, T, ff._lib.FFB_C2C)
    dz = dev(ff, z)
    cplan.mul(dz, dz)
    assert relerr(dz.to_numpy(), np.fft.fftn(z.astype(np.complex128))) <= tol


@pytest.mark.parametrize("T,tol", [(np.float64, 1e-12), (np.float32, 1e-5)], ids=["f64", "f32"])
def test_free_fft_functions(ff, T, tol):
    """`fft / ifft / rfft / irfft` (FFTW names re-exported by the reference, src/FourierFlows.jl:72) as allocating whole-array
    transforms; plans are cached per (shape, type, kind) and the least recently used one is dropped."""
    rng = np.random.default_rng(12)
    cT = np.complex64 if T == np.float32 else np.complex128
    for shape in ((30,), (64,), (16, 30), (128, 64), (8, 8, 8), (32, 30, 16)):      # six shapes: more than the cache holds
        x = np.asfortranarray(rng.standard_normal(shape).astype(T))
        xh = ff.rfft(dev(ff, x))
        assert xh.shape == (shape[0] // 2 + 1,) + shape[1:] and xh.dtype == cT
        assert relerr(xh.to_numpy(), fo.RfftPlan(shape, T) * x.astype(np.float64)) <= tol
        assert relerr(ff.irfft(xh, shape[0]).to_numpy(), x) <= tol
        z = np.asfortranarray((rng.standard_normal(shape) + 1j * rng.standard_normal(shape)).astype(cT))
        zh = ff.fft(dev(ff, z))
        assert relerr(zh.to_numpy(), np.fft.fftn(z.astype(np.complex128))) <= tol
        assert relerr(ff.ifft(zh).to_numpy(), z) <= tol
    from fourierflows_jl_b200 import utils
    assert len(utils._FREE_PLANS) <= utils._FREE_PLANS_MAX
    with pytest.raises(TypeError):
        ff.rfft(dev(ff, z))
    with pytest.raises(ValueError):
        ff.irfft(xh, 2 * shape[0])
