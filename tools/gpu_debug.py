"""Stage-by-stage comparison of the CUDA library against the FP64 CPU port at a large size
(debugging aid; the port itself is validated against the long-double oracle on the CPU)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import healpixmpi_jl_b200 as H
from oracle import oracle as O, cpu_port as P
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import test_gpu_parity as T

nside, lmax = int(sys.argv[1]), int(sys.argv[2])
spins = [int(s) for s in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0, 2]
rng = np.random.default_rng(1)
rings = np.arange(1, 4 * nside)
nphi, theta, phi0, fp = O.healpix_rings(nside, rings)
mval = np.arange(lmax + 1); mstart0 = O.rr_mstart1(lmax, 1, mval) - 1
nalm = H.numberOfAlms(lmax)
npix = 12 * nside * nside

def report(name, got, ref, axis_names):
    d = np.abs(got - ref) ** 2
    tot = np.sqrt(d.sum() / (np.abs(ref) ** 2).sum())
    print(f"{name}: rel L2 {tot:.3e}  max abs {np.sqrt(d.max()):.3e} finite={np.isfinite(got.view(np.float64)).all()}")
    for ax, nm in axis_names:
        other = tuple(i for i in range(got.ndim) if i != ax)
        num = d.sum(axis=other); den = (np.abs(ref) ** 2).sum(axis=other)
        rel = np.sqrt(num / np.maximum(den, 1e-300))
        worst = np.argsort(-rel)[:8]
        print(f"   worst along {nm}: " + ", ".join(f"{int(i)}:{rel[i]:.1e}" for i in worst), " median %.1e" % np.median(rel))
        bad = np.nonzero(rel > 1e-10)[0]
        if len(bad):
            print(f"   {len(bad)} bad along {nm}; first {bad[:12]} last {bad[-12:]}")

for spin in spins:
    nc = 1 if spin == 0 else 2
    alm = T._rand_alm(rng, nc, nalm)
    t = time.time(); ref = P.alm2leg(alm, spin, lmax, mval, mstart0, theta); tp = time.time() - t
    t = time.time(); got = T._stage_alm2leg(H, alm, spin, lmax, mval, mstart0, theta); tg = time.time() - t
    print(f"== spin {spin} alm2leg (port {tp:.2f}s gpu-call {tg:.2f}s)")
    report("alm2leg", got, ref, [(1, "ring"), (2, "m")])
    leg = ref
    t = time.time(); got = T._stage_leg2alm(H, leg, spin, lmax, mval, mstart0, theta, nalm); tg = time.time() - t
    refa = P.leg2alm(leg, spin, lmax, mval, mstart0, theta, nalm)
    print(f"== spin {spin} leg2alm (gpu-call {tg:.2f}s)")
    report("leg2alm", got, refa, [])
    # per-m error
    rel = []
    for m in mval:
        a = slice(mstart0[m] + m, mstart0[m] + lmax + 1)
        rel.append(np.sqrt((np.abs(got[:, a] - refa[:, a]) ** 2).sum() / max((np.abs(refa[:, a]) ** 2).sum(), 1e-300)))
    rel = np.array(rel); bad = np.nonzero(rel > 1e-10)[0]
    print("   per-m worst", [(int(i), float(rel[i])) for i in np.argsort(-rel)[:6]], "nbad", len(bad), bad[:10], bad[-10:])
    if spin == 0:
        nsel = min(len(rings), 400)
        sel = np.unique(np.linspace(0, len(rings) - 1, nsel).astype(int))
        rs0 = np.concatenate([[0], np.cumsum(nphi[sel])[:-1]]).astype(np.int64)
        lg = leg[:, sel, :]
        refm = P.leg2map(lg, nphi[sel], phi0[sel], rs0, int(nphi[sel].sum()))
        gotm = T._stage_leg2map(H, lg, nphi[sel], phi0[sel], rs0, int(nphi[sel].sum()))
        print("== leg2map")
        per = np.array([np.sqrt(((gotm[0, a:a + n] - refm[0, a:a + n]) ** 2).sum() / (refm[0, a:a + n] ** 2).sum()) for a, n in zip(rs0, nphi[sel])])
        print("   total", np.sqrt(((gotm - refm) ** 2).sum() / (refm ** 2).sum()), "worst rings", [(int(rings[sel][i]), int(nphi[sel][i]), float(per[i])) for i in np.argsort(-per)[:8]])
        mp = rng.standard_normal((1, int(nphi[sel].sum())))
        refl = P.map2leg(mp, lmax + 1, nphi[sel], phi0[sel], rs0)
        gotl = T._stage_map2leg(H, mp, lmax + 1, nphi[sel], phi0[sel], rs0)
        per = np.sqrt((np.abs(gotl - refl) ** 2).sum(axis=(0, 2)) / (np.abs(refl) ** 2).sum(axis=(0, 2)))
        print("== map2leg total", np.sqrt((np.abs(gotl - refl) ** 2).sum() / (np.abs(refl) ** 2).sum()), "worst rings", [(int(rings[sel][i]), int(nphi[sel][i]), float(per[i])) for i in np.argsort(-per)[:8]])
