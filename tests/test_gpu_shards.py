"""Device-side shard bookkeeping and algebra against the host restatements of the
reference's Scatter!/Gather!/dot/almxfl!/alm2cl (healpixmpi.jl_b200/alm.py, map.py; src/alm.jl, src/map.jl,
src/cl.jl).  Index work is bit-exact; the two reductions differ from numpy only in summation order."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def Hm():
    import healpixmpi_jl_b200 as H

    return H


def _global(Hm, rng, nside, lmax, ncomp):
    nalm = Hm.numberOfAlms(lmax)
    galm = rng.standard_normal((ncomp, nalm)) + 1j * rng.standard_normal((ncomp, nalm))
    gmap = rng.standard_normal((ncomp, 12 * nside * nside))
    return galm, gmap


@pytest.mark.parametrize("nside,lmax", [(4, 11), (16, 40), (64, 191)])
@pytest.mark.parametrize("ncomp", [1, 2])
def test_scatter_gather_single_rank(Hm, nside, lmax, ncomp):
    rng = np.random.default_rng(nside + ncomp)
    galm, gmap = _global(Hm, rng, nside, lmax, ncomp)
    plan = Hm.Plan(nside, lmax, Hm.COMM_SELF, 2)
    # host restatement of the reference
    alms = [Hm.Alm(lmax, lmax, galm[c].copy()) for c in range(ncomp)]
    d_alm = Hm.DAlm()
    Hm.Scatter_(alms if ncomp == 2 else alms[0], d_alm)
    d_map = Hm.DMap()
    Hm.Scatter_(Hm.HealpixMap(nside, gmap[0].copy()) if ncomp == 1 else
                Hm.PolarizedHealpixMap(nside, None, gmap[0].copy(), gmap[1].copy()), d_map)
    # device scatter
    a_loc = np.full((ncomp, plan.nalm_loc), np.nan + 0j)
    plan.scatter_alm(galm, a_loc, ncomp, 0)
    assert np.array_equal(a_loc, d_alm.alm.T)
    m_loc = np.full((ncomp, plan.npix_loc), np.nan)
    plan.scatter_map(gmap, m_loc, ncomp, 0)
    assert np.array_equal(m_loc, d_map.pixels.T)
    # gather and allgather invert it exactly
    for root in (0, -1):
        back = np.full_like(galm, np.nan)
        plan.gather_alm(a_loc, back, ncomp, root)
        assert np.array_equal(back, galm)
        backm = np.full_like(gmap, np.nan)
        plan.gather_map(m_loc, backm, ncomp, root)
        assert np.array_equal(backm, gmap)
    plan.destroy()


def test_scatter_gather_device_pointers(Hm):
    import torch

    nside, lmax = 32, 95
    rng = np.random.default_rng(5)
    galm, gmap = _global(Hm, rng, nside, lmax, 1)
    plan = Hm.Plan(nside, lmax, Hm.COMM_SELF, 1)
    d_g = torch.from_numpy(galm.view(np.float64).copy()).cuda()
    d_loc = torch.zeros((1, plan.nalm_loc, 2), dtype=torch.float64, device="cuda")
    plan.scatter_alm(d_g, d_loc, 1, 0)
    host = np.zeros((1, plan.nalm_loc), dtype=np.complex128)
    plan.scatter_alm(galm, host, 1, 0)
    assert np.array_equal(d_loc.cpu().numpy().view(np.complex128)[..., 0], host)
    d_back = torch.zeros_like(d_g)
    plan.gather_alm(d_loc, d_back, 1, 0)
    assert torch.equal(d_back, d_g)
    # the scattered shard feeds the transform directly, everything device resident
    d_map = torch.zeros((1, plan.npix_loc), dtype=torch.float64, device="cuda")
    plan.alm2map(d_loc, d_map, 1)
    d_full = torch.zeros((1, 12 * nside * nside), dtype=torch.float64, device="cuda")
    plan.gather_map(d_map, d_full, 1, -1)
    ref = np.zeros((1, plan.npix_loc))
    plan.alm2map(host, ref, 1)
    full_ref = np.zeros((1, 12 * nside * nside))
    plan.gather_map(ref, full_ref, 1, 0)
    assert np.array_equal(d_full.cpu().numpy(), full_ref)
    plan.destroy()


def test_dot_reference_known_answer(Hm):
    """test/test_MPI_dot.jl:17-30: Alm(5, 5, [1..21]) scattered, d_alm . d_alm == 6531."""
    plan = Hm.Plan(4, 5, Hm.COMM_SELF, 1)
    galm = np.arange(1, Hm.numberOfAlms(5) + 1, dtype=np.float64).astype(np.complex128)[None, :]
    loc = np.zeros((1, plan.nalm_loc), dtype=np.complex128)
    plan.scatter_alm(galm, loc, 1, 0)
    assert plan.dot(loc[0], loc[0]) == 6531.0
    plan.destroy()


@pytest.mark.parametrize("nside,lmax", [(8, 23), (64, 150)])
def test_dot_almxfl_alm2cl(Hm, nside, lmax):
    rng = np.random.default_rng(lmax)
    plan = Hm.Plan(nside, lmax, Hm.COMM_SELF, 2)
    nalm = Hm.numberOfAlms(lmax)
    a = [Hm.Alm(lmax, lmax, rng.standard_normal(nalm) + 1j * rng.standard_normal(nalm)) for _ in range(2)]
    da = Hm.DAlm()
    Hm.Scatter_(a, da)
    loc = np.ascontiguousarray(da.alm.T)  # (2, nalm_loc)
    for c1, c2 in ((0, 0), (0, 1)):
        ref = Hm.dot(da, da, c1 + 1, c2 + 1)
        got = plan.dot(loc[c1], loc[c2])
        assert abs(got - ref) <= 1e-13 * abs(ref) + 1e-13
    # almxfl with a short fl (zero-extended, src/alm.jl:699-701)
    fl = rng.standard_normal(lmax - 3)
    # the reference scales min(ncomp, ncol(fl)) columns: a vector fl touches the first component only
    mval, mstart0 = plan.alm_info()
    for fls in (fl[None, :], np.stack([fl, 2.0 * fl])):
        x = loc.copy()
        plan.almxfl(x, fls if fls.shape[0] > 1 else fl, 2)
        ref = loc.copy()
        for c in range(fls.shape[0]):
            flx = np.concatenate([fls[c], np.zeros(lmax + 1 - fls.shape[1])])
            for mi, m in enumerate(mval):
                sl = slice(mstart0[mi] + m, mstart0[mi] + lmax + 1)
                ref[c, sl] = loc[c, sl] * flx[m:]
        assert np.array_equal(x, ref)
    # alm2cl (src/cl.jl:5-38)
    cl = plan.alm2cl(loc[0], loc[1])
    refcl = np.zeros(lmax + 1)
    for mi, m in enumerate(mval):
        sl = slice(mstart0[mi] + m, mstart0[mi] + lmax + 1)
        w = 2.0 if m > 0 else 1.0
        refcl[m:] += w * (loc[0, sl] * np.conj(loc[1, sl])).real / (2 * np.arange(m, lmax + 1) + 1)
    assert np.allclose(cl, refcl, rtol=1e-13, atol=1e-15)
    plan.destroy()
