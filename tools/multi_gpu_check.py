"""N-rank check of the fused plan (NCCL exchange): torchrun --nproc-per-node N tools/multi_gpu_check.py
Each rank scatters its RR shard from a common seeded alm / map, runs alm2map_ / adjoint_alm2map_
through the public API, gathers, and rank 0 compares with (a) the oracle at small size and (b) a
single-rank plan on its own GPU at large size (bit-comparable up to summation order)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import healpixmpi_jl_b200 as H

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
comm = H.TorchComm()
ok = True

def rel(a, b):
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))

for (nside, lmax, use_oracle) in [(8, 23, True), (64, 191, True), (64, 100, True), (1024, 3071, False)]:
    nalm = H.numberOfAlms(lmax); npix = 12 * nside * nside
    rng = np.random.default_rng(1000 + nside)  # identical on every rank
    for ncomp in (1, 2):
        spin = 0 if ncomp == 1 else 2
        alms = [H.Alm(lmax, lmax, (rng.standard_normal(nalm) + 1j * rng.standard_normal(nalm))) for _ in range(ncomp)]
        maps = [rng.standard_normal(npix) for _ in range(ncomp)]
        d_alm, d_map = H.DAlm(), H.DMap()
        H.Scatter_(alms if ncomp == 2 else alms[0], d_alm, comm=comm)
        H.Scatter_(H.HealpixMap(nside, maps[0]) if ncomp == 1 else H.PolarizedHealpixMap(nside, None, maps[0], maps[1]), d_map, comm=comm)
        d_map_in = d_map.pixels.copy(order='F')
        t0 = time.time()
        H.alm2map_(d_alm, d_map)
        t1 = time.time()
        got_map = [H.HealpixMap(nside) for _ in range(ncomp)]
        for c in range(ncomp):
            H.Gather_(d_map, got_map[c] if rank == 0 else None, c + 1)
        d_map.pixels = d_map_in
        d_alm_out = H.DAlm(); H.Scatter_(alms if ncomp == 2 else alms[0], d_alm_out, comm=comm)
        H.adjoint_alm2map_(d_map, d_alm_out)
        got_alm = [H.Alm(lmax, lmax) for _ in range(ncomp)]
        for c in range(ncomp):
            H.Gather_(d_alm_out, got_alm[c] if rank == 0 else None, c + 1)
        if rank == 0:
            if use_oracle:
                from oracle import oracle as O
                refm, _ = O.full_alm2map(np.stack([a.alm for a in alms]), spin, nside, lmax)
                refm = refm.astype(np.float64)
                refa = O.full_adjoint(np.stack(maps), spin, nside, lmax).astype(np.complex128)
                if spin == 2:
                    for m in (0, 1):
                        for l in range(m, min(2, lmax + 1)):
                            refa[:, alms[0].index(l, m)] = 0
                tag = "oracle"
            else:
                p1 = H.Plan(nside, lmax, H.COMM_SELF, ncomp)
                mv, ms0 = p1.alm_info(); mi = p1.map_info()
                idx = alms[0].each_ell_idx(mv)
                aloc = np.ascontiguousarray(np.stack([a.alm[idx] for a in alms]))
                mloc = np.zeros((ncomp, p1.npix_loc))
                p1.alm2map(aloc, mloc, ncomp)
                from oracle import oracle as O
                nphi, _, _, fp = O.healpix_rings(nside, mi["rings"])
                pix = np.concatenate([np.arange(f, f + n) for f, n in zip(fp, nphi)])
                refm = np.zeros((ncomp, npix)); refm[:, pix] = mloc
                fin = np.ascontiguousarray(np.stack(maps)[:, pix])
                aout = np.zeros((ncomp, p1.nalm_loc), dtype=np.complex128)
                p1.adjoint_alm2map(fin, aout, ncomp)
                refa = np.zeros((ncomp, nalm), dtype=np.complex128); refa[:, idx] = aout
                p1.destroy(); tag = "1-rank plan"
            em = rel(np.stack([g.pixels for g in got_map]), refm)
            ea = rel(np.stack([g.alm for g in got_alm]), refa)
            good = em < 5e-13 and ea < 5e-13
            ok &= good
            print(f"P={world} nside={nside} lmax={lmax} spin={spin}: map rel {em:.2e} alm rel {ea:.2e} vs {tag} "
                  f"(alm2map_ {1e3*(t1-t0):.1f} ms incl. first-call setup) {'OK' if good else 'FAIL'}", flush=True)
    # ---- device-side scatter / gather / algebra against the host restatements ----
    if nside <= 64:
        plan = H.Plan(nside, lmax, comm, 2)
        rng = np.random.default_rng(77 + nside)
        galm = rng.standard_normal((2, nalm)) + 1j * rng.standard_normal((2, nalm))
        gmap = rng.standard_normal((2, npix))
        d_a, d_m = H.DAlm(), H.DMap()
        H.Scatter_([H.Alm(lmax, lmax, galm[0].copy()), H.Alm(lmax, lmax, galm[1].copy())], d_a, comm=comm)
        H.Scatter_(H.PolarizedHealpixMap(nside, None, gmap[0].copy(), gmap[1].copy()), d_m, comm=comm)
        a_loc = np.full((2, plan.nalm_loc), np.nan + 0j)
        m_loc = np.full((2, plan.npix_loc), np.nan)
        root = world - 1
        plan.scatter_alm(galm if rank == root else None, a_loc, 2, root)
        plan.scatter_map(gmap if rank == root else None, m_loc, 2, root)
        good = np.array_equal(a_loc, d_a.alm.T) and np.array_equal(m_loc, d_m.pixels.T)
        back_a = np.full_like(galm, np.nan); back_m = np.full_like(gmap, np.nan)
        plan.gather_alm(a_loc, back_a, 2, -1)
        plan.gather_map(m_loc, back_m, 2, -1)
        good &= np.array_equal(back_a, galm) and np.array_equal(back_m, gmap)
        back_a[:] = np.nan
        plan.gather_alm(a_loc, back_a if rank == 0 else None, 2, 0)
        if rank == 0:
            good &= np.array_equal(back_a, galm)
        ref = H.dot(d_a, d_a, 1, 2)
        got = plan.dot(a_loc[0], a_loc[1])
        good &= abs(got - ref) <= 1e-13 * abs(ref) + 1e-13
        cl = plan.alm2cl(a_loc[0], a_loc[1])
        l_of = np.concatenate([np.arange(m, lmax + 1) for m in range(lmax + 1)])
        w_of = np.concatenate([np.full(lmax + 1 - m, 1.0 if m == 0 else 2.0) for m in range(lmax + 1)])
        refcl = np.bincount(l_of, weights=w_of * (galm[0] * np.conj(galm[1])).real / (2 * l_of + 1), minlength=lmax + 1)
        good &= np.allclose(cl, refcl, rtol=1e-12, atol=1e-14)
        flags = comm.allgather(bool(good))
        if rank == 0:
            ok &= all(flags)
            print(f"P={world} nside={nside} lmax={lmax}: device scatter/gather/allgather bit-exact, dot/alm2cl vs host "
                  f"{'OK' if all(flags) else 'FAIL ' + str(flags)}", flush=True)
        plan.destroy()
    H.clear_plans()
dist.barrier()
if rank == 0:
    print("ALL OK" if ok else "FAILED", flush=True)
dist.destroy_process_group()
sys.exit(0 if ok else 1)
