"""Grid-side kernels through the C ABI against the oracle: wavenumbers (bit-exact), Ksq/invKsq, `dealias!`
(exact zeros, retained box), `makefilter`, alias ranges.  Restates test/test_grid.jl on the device."""
import numpy as np
import pytest

import oracle as fo
from util import relerr

pytestmark = pytest.mark.gpu
nx, Lx, ny, Ly, nz, Lz = 6, 2 * np.pi, 8, 4 * np.pi, 10, 3.0


@pytest.fixture(scope="module")
def ff():
    import fourierflows_jl_b200 as ff
    assert ff.have_device()
    return ff


def grids(ff, T=np.float64, **kw):
    d = ff.GPU()
    return ((ff.OneDGrid(d, nx=nx, Lx=Lx, T=T, **kw), fo.OneDGrid(nx=nx, Lx=Lx, T=T, **kw)),
            (ff.TwoDGrid(d, nx=nx, Lx=Lx, ny=ny, Ly=Ly, T=T, **kw), fo.TwoDGrid(nx=nx, Lx=Lx, ny=ny, Ly=Ly, T=T, **kw)),
            (ff.ThreeDGrid(d, nx=nx, Lx=Lx, ny=ny, Ly=Ly, nz=nz, Lz=Lz, T=T, **kw), fo.ThreeDGrid(nx=nx, Lx=Lx, ny=ny, Ly=Ly, nz=nz, Lz=Lz, T=T, **kw)))


@pytest.mark.parametrize("T", [np.float64, np.float32])
def test_wavenumbers_and_spacing_bit_exact(ff, T):
    for g, og in grids(ff, T):
        assert g.k.to_numpy().dtype == np.dtype(T)
        assert np.array_equal(g.k.to_numpy(), og.k.ravel()) and np.array_equal(g.kr.to_numpy(), og.kr.ravel())
        assert g.dx == og.dx and np.array_equal(g.x, og.x) and g.nkr == og.nkr
        if g.ndim >= 2:
            assert np.array_equal(g.l.to_numpy(), og.l.ravel()) and np.array_equal(g.y, og.y)
        if g.ndim == 3:
            assert np.array_equal(g.m.to_numpy(), og.m.ravel()) and np.array_equal(g.z, og.z)
    # kr[nkr] is the +Nyquist wavenumber (test_grid.jl:48)
    g, og = grids(ff, T)[0]
    assert g.kr.to_numpy()[-1] == abs(g.k.to_numpy()[g.nkr - 1]) > 0


@pytest.mark.parametrize("T", [np.float64, np.float32])
def test_dense_ksq_arrays(ff, T):
    for g, og in grids(ff, T)[1:]:
        for name in ("Ksq", "invKsq", "Krsq", "invKrsq"):
            got, ref = getattr(g, name).to_numpy(), getattr(og, name)
            assert got.shape == ref.shape
            assert np.array_equal(got, ref), name
        assert g.invKsq.to_numpy().ravel()[0] == 0 and g.invKrsq.to_numpy().ravel()[0] == 0
    g, og = grids(ff, T)[0]
    assert np.array_equal(g.invkrsq.to_numpy(), og.invkrsq)


def test_dealias_matches_reference_rule(ff):
    """testdealias (test_grid.jl:80-133) and equality with the oracle, including a trailing field dimension"""
    for g, og in grids(ff):
        for lead in (g.nkr, g.nk):
            shape = (lead,) + tuple(g.shape[1:]) + (2,)
            rng = np.random.default_rng(3)
            fh = np.asfortranarray(rng.standard_normal(shape) + 1j * rng.standard_normal(shape))
            d = ff.DevArray.from_numpy(fh)
            assert ff.dealias(d, g) is None
            ref = fh.copy(order="F")
            fo.dealias(ref, og)
            assert np.array_equal(d.to_numpy(), ref)
    g, og = grids(ff)[2]
    fh = ff.DevArray.from_numpy(np.ones((g.nkr, g.nl, g.nm), dtype=complex, order="F"))
    ff.dealias(fh, g)
    out = fh.to_numpy()
    kmax = round(float(og.kr.max()) * 2 / 3)
    lmax = round(float(np.abs(og.l).max()) * 2 / 3)
    mmax = round(float(np.abs(og.m).max()) * 2 / 3)
    keep = (og.kr < kmax) & (og.l < lmax) & (og.l >= -lmax) & (og.m < mmax) & (og.m >= -mmax)
    out[keep] = 0  # the reference asserts only that nothing survives outside the box (test_grid.jl:117-133)
    assert np.abs(out).sum() == 0


def test_no_dealias_with_zero_fraction(ff):
    for g, og in grids(ff, aliased_fraction=0):
        fh = ff.DevArray.from_numpy(np.ones((g.nkr,) + tuple(g.shape[1:]), dtype=complex, order="F"))
        assert ff.dealias(fh, g) is None
        assert np.all(fh.to_numpy() == 1)
        assert g.kalias is None and g.kralias is None


@pytest.mark.parametrize("a", [0, 1 / 3, 1 / 2, 1 / 4])
def test_aliased_fraction_ranges(ff, a):
    """test_aliased_fraction (test_grid.jl:203-227) on 16 x 32 x 34"""
    g = ff.ThreeDGrid(ff.GPU(), nx=16, Lx=Lx, ny=32, Ly=Lx, nz=34, Lz=Lx, aliased_fraction=a)
    og = fo.ThreeDGrid(nx=16, Lx=Lx, ny=32, Ly=Lx, nz=34, Lz=Lx, aliased_fraction=a)
    assert (g.kalias, g.kralias, g.lalias, g.malias) == (og.kalias, og.kralias, og.lalias, og.malias)
    assert g.aliased_fraction == og.aliased_fraction


@pytest.mark.parametrize("T", [np.float64, np.float32])
def test_makefilter(ff, T):
    """testmakefilter (test_grid.jl:165-173) + agreement with the oracle"""
    for g, og in grids(ff, T):
        f = ff.makefilter(g).to_numpy()
        ref = fo.makefilter(og)
        K = fo.fforacle._nondimK(og, True)
        assert f.dtype == np.dtype(T) and f.shape == ref.shape
        assert np.all(f[K < 0.65] == 1)
        assert np.all(np.abs(f[K > 0.999]) <= 1e-12)
        assert relerr(f, ref) <= (1e-14 if T == np.float64 else 1e-6)
    g, og = grids(ff, T)[1]
    f = ff.makefilter(g, T, (g.nkr, g.nl, 3), innerK=0.0, outerK=1 / 16).to_numpy()
    ref = fo.makefilter(og, T, (og.nkr, og.nl, 3), innerK=0.0, outerK=1 / 16)
    assert f.shape == (g.nkr, g.nl, 3) and relerr(f, ref) <= (1e-13 if T == np.float64 else 1e-6)
    fc = ff.makefilter(g, realvars=False).to_numpy()
    assert relerr(fc, fo.makefilter(og, realvars=False)) <= (1e-14 if T == np.float64 else 1e-6)


def test_parseval_and_cpu_device_rejected(ff):
    """parsevalsum2 (test_utils.jl:82-96 analogue) and the no-CPU-fallback rule"""
    g = ff.TwoDGrid(ff.GPU(), nx=64, Lx=2 * np.pi, ny=128, Ly=3 * np.pi)
    og = fo.TwoDGrid(nx=64, Lx=2 * np.pi, ny=128, Ly=3 * np.pi)
    x, y = og.x.reshape(-1, 1), og.y.reshape(1, -1)
    u = np.exp(-(x ** 2 + y ** 2) / 0.5)
    uh = g.rfftplan * ff.DevArray.from_numpy(np.asfortranarray(u))
    integral = float(np.sum(u ** 2) * og.dx * og.dy)
    assert abs(ff.parsevalsum2(uh, g) - integral) <= 1e-13 * integral
    uhc = g.fftplan * ff.DevArray.from_numpy(np.asfortranarray(u.astype(complex)))
    assert abs(ff.parsevalsum2(uhc, g) - integral) <= 1e-13 * integral
    with pytest.raises(ff.FFBError):
        ff.OneDGrid(ff.CPU(), nx=8, Lx=1.0)
