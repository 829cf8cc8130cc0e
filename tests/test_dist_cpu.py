"""World-size-2 (and 3) gloo tests of the N>1 host logic on CPU: Scatter!/Gather! round trips over a
real process group and the reference-faithful leg redistribution (communicate_alm2map! /
communicate_map2alm!, src/sht.jl:5-44,106-135) checked against single-rank oracle legs."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, nside, lmax, q):
    try:
        sys.path.insert(0, ROOT)
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        import torch.distributed as dist

        dist.init_process_group("gloo", rank=rank, world_size=world)
        import healpixmpi_jl_b200 as H
        from oracle import oracle as O

        comm = H.TorchComm()
        assert comm.rank() == rank and comm.size() == world
        rng = np.random.default_rng(5)  # same stream on every rank
        nalm = H.numberOfAlms(lmax)
        npix = 12 * nside * nside
        # ---- alm scatter / gather / allgather (test/test_MPI_alm.jl, test_MPI_alm_pol.jl) ----
        alms = [H.Alm(lmax, lmax, rng.standard_normal(nalm) + 1j * rng.standard_normal(nalm)) for _ in range(3)]
        dt, dp = H.DAlm(), H.DAlm()
        H.Scatter_(alms if rank == 0 else None, dt, dp, comm)
        ref_loc, mval, mstart0 = O.scatter_alm(np.stack([a.alm for a in alms[1:]]), lmax, rank, world)
        assert np.array_equal(dp.info.mval, mval) and np.array_equal(dp.info.mstart, mstart0 + 1)
        assert np.array_equal(dp.alm.T, ref_loc) and dp.info.maxnm == O.rr_nm(lmax, 0, world)
        out = H.Alm(lmax, lmax) if rank == 0 else None
        H.Gather_(dp, out, 2)
        if rank == 0:
            assert np.array_equal(out.alm, alms[2].alm)
        out = H.Alm(lmax, lmax)
        H.Allgather_(dt, out)
        assert np.array_equal(out.alm, alms[0].alm)
        d2 = H.dot(dt, dt)
        w = np.concatenate([np.full(lmax + 1 - m, 1.0 if m == 0 else 2.0) for m in range(lmax + 1)])
        a0 = alms[0].alm.copy()
        a0[:lmax + 1] = a0[:lmax + 1].real
        assert abs(d2 - np.sum(w * np.abs(a0) ** 2)) < 1e-9 * d2
        # ---- map scatter / gather (test/test_MPI_map.jl, test_MPI_map_pol.jl, test_MPI_arr.jl) ----
        pm = H.PolarizedHealpixMap(nside, *[rng.standard_normal(npix) for _ in range(3)])
        di, dq = H.DMap(), H.DMap()
        H.Scatter_(pm if rank == 0 else None, di, dq, comm)
        ref_pix, info = O.scatter_map(np.stack([pm.q.pixels, pm.u.pixels]), nside, rank, world)
        assert np.array_equal(dq.pixels.T, ref_pix)
        assert np.array_equal(dq.info.rings, info["rings"]) and np.array_equal(dq.info.rstart, info["rstart0"] + 1)
        assert np.array_equal(dq.info.thetatot, info["thetatot"]) and np.array_equal(dq.info.phi0, info["phi0"])
        o = H.HealpixMap(nside) if rank == 0 else None
        H.Gather_(dq, o, 2)
        if rank == 0:
            assert np.array_equal(o.pixels, pm.u.pixels)
        o = H.HealpixMap(nside)
        H.Allgather_(di, o)
        assert np.array_equal(o.pixels, pm.i.pixels)
        arr = H.Scatter_array(pm.i.pixels if rank == 0 else None, nside, comm)
        assert np.array_equal(arr, di.pixels[:, 0])
        # ---- leg redistribution against single-rank oracle legs, spin 0 and spin 2 ----
        for spin, d_alm, d_map in ((0, dt, di), (2, dp, dq)):
            ncomp = d_alm.ncomp
            full = np.stack([a.alm for a in (alms[:1] if spin == 0 else alms[1:])])
            loc, mval, mstart0 = O.scatter_alm(full, lmax, rank, world)
            thetatot = d_map.info.thetatot
            # alm-side leg of this rank: my m's x ALL rings in thetatot order   (loc_nm, tot_nr, ncomp)
            leg_a = O.alm2leg(loc, spin, lmax, mval, mstart0, thetatot).astype(np.complex128)  # [c, ring, m]
            in_leg = np.asfortranarray(np.transpose(leg_a, (2, 1, 0)))
            a_buf, b_buf = H.alloc_leg_buffers(d_alm, d_map)
            assert a_buf.shape == in_leg.shape
            H.communicate_alm2map_(in_leg, b_buf, comm)
            # expected map-side leg: ALL m x my rings, computed directly
            allm = np.arange(lmax + 1)
            ms_all = O.rr_mstart1(lmax, 1, allm) - 1
            exp = O.alm2leg(full, spin, lmax, allm, ms_all, d_map.info.theta).astype(np.complex128)
            assert np.array_equal(b_buf, np.transpose(exp, (2, 1, 0)))
            # and back (communicate_map2alm!)
            back = np.zeros_like(a_buf)
            H.communicate_map2alm_(b_buf, back, comm)
            assert np.array_equal(back, in_leg)
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback

        q.put((rank, "FAIL: " + traceback.format_exc()))


@pytest.mark.parametrize("world,nside,lmax", [(2, 4, 11), (3, 4, 9), (2, 8, 23)])
def test_gloo_multi_rank_host_logic(world, nside, lmax):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, nside, lmax, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, msg in res:
        assert msg == "ok", f"rank {rank}: {msg}"
