"""Parity evidence at the BASELINE sizes (run on the GPU box; CPU part = long-double oracle, OpenMP).

  python tools/parity_report.py full 1024 3071 0      full-map relative L2 of alm2map_ vs oracle
  python tools/parity_report.py sampled 4096 8191 0   sampled rings (all m) + sampled m (all rings)
"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import healpixmpi_jl_b200 as H
from oracle import oracle as O

mode, nside, lmax, spin = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
ncomp = 1 if spin == 0 else 2
rng = np.random.default_rng(4096 + spin)
plan = H.Plan(nside, lmax, H.COMM_SELF, ncomp)
mval, mstart0 = plan.alm_info()
info = plan.map_info()
alm = (rng.standard_normal((ncomp, plan.nalm_loc)) + 1j * rng.standard_normal((ncomp, plan.nalm_loc))) * np.sqrt(0.5)
mp = np.zeros((ncomp, plan.npix_loc))
plan.alm2map(alm, mp, ncomp)
print(f"# {mode} nside={nside} lmax={lmax} spin={spin}: GPU alm2map {plan.timings()['total']:.1f} ms (host buffers)", flush=True)
loc = {int(r): (int(s), int(n)) for r, s, n in zip(info["rings"], info["rstart0"], info["nphi"])}

def rel(a, b):
    return float(np.sqrt(np.sum(np.abs(a - b) ** 2) / np.sum(np.abs(b) ** 2)))

if mode == "full":
    sel = np.arange(1, 4 * nside)
else:
    # pseudo-random north rings, pixel-weighted (ring index ~ sqrt in the caps) + the special ones
    rs = np.random.default_rng(99)
    north = sorted(set([1, 2, nside, 2 * nside - 1, 2 * nside] +
                       [int(x) for x in np.sqrt(rs.uniform(1, nside ** 2, 10))] +
                       [int(x) for x in rs.uniform(nside, 2 * nside, 14)]))
    sel = np.array(north + [4 * nside - r for r in north if r != 2 * nside])
nphi, theta, phi0, fp = O.healpix_rings(nside, sel)
cut = os.environ.get("ORACLE_CUT", "0") == "1"
O.set_reference_cut(cut)
print(f"# oracle applies the reference's m > mlim cut: {cut}", flush=True)
t = time.time()
leg = O.alm2leg(alm, spin, lmax, mval, mstart0, theta)
rs0 = np.concatenate([[0], np.cumsum(nphi)[:-1]]).astype(np.int64)
ref = O.leg2map(leg, nphi, phi0, rs0, int(nphi.sum())).astype(np.float64)
print(f"# oracle {time.time()-t:.0f} s for {len(sel)} rings", flush=True)
got = np.concatenate([mp[:, loc[int(r)][0]:loc[int(r)][0] + loc[int(r)][1]] for r in sel], axis=1)
per = np.array([rel(got[:, a:a + n], ref[:, a:a + n]) for a, n in zip(rs0, nphi)])
num = np.array([np.sum((got[:, a:a + n] - ref[:, a:a + n]) ** 2) for a, n in zip(rs0, nphi)])
den = np.array([np.sum(ref[:, a:a + n] ** 2) for a, n in zip(rs0, nphi)])
print(f"relative L2 over the compared pixels ({int(nphi.sum())} px x {ncomp} comp): {np.sqrt(num.sum()/den.sum()):.3e}")
polar = nphi < 64
if (~polar).any():
    print(f"   excluding rings with < 64 pixels: {np.sqrt(num[~polar].sum()/den[~polar].sum()):.3e}")
order = np.argsort(-per)[:8]
print("   worst rings (ring, nphi, rel L2):", [(int(sel[i]), int(nphi[i]), float(f'{per[i]:.2e}')) for i in order])
print(f"   median per-ring rel L2 {np.median(per):.2e}")
if mode == "sampled":
    # adjoint: sampled m (all rings) against the oracle
    f = rng.standard_normal((ncomp, plan.npix_loc))
    out = np.zeros((ncomp, plan.nalm_loc), dtype=np.complex128)
    plan.adjoint_alm2map(f, out, ncomp)
    allr = info["rings"]
    nphi_a, theta_a, phi0_a, _ = O.healpix_rings(nside, allr)
    msel = np.array(sorted(set([0, 1, 2, 3, 10, 100, 1000, lmax // 2, lmax - 100, lmax - 1, lmax])))
    t = time.time()
    # map-side legs of the sampled m in long double: scipy's pocketfft is templated on long double
    # (independent of our FFT), then the un-alias + phase shift of SURVEY A9 in long double
    import scipy.fft
    legs = np.zeros((ncomp, len(allr), len(msel)), dtype=np.clongdouble)
    for c in range(ncomp):
        for ir, (s0, n) in enumerate(zip(info["rstart0"], nphi_a)):
            n = int(n)
            X = scipy.fft.rfft(f[c, s0:s0 + n].astype(np.longdouble))
            k = msel % n
            up = k > n // 2
            v = np.where(up, np.conj(X[np.where(up, n - k, 0)]), X[np.where(up, 0, k)])
            ang = msel.astype(np.longdouble) * np.longdouble(phi0_a[ir])
            legs[c, ir, :] = v * (np.cos(ang) - 1j * np.sin(ang))
    ms0 = O.rr_mstart1(lmax, 1, np.arange(lmax + 1))[msel] - 1
    refa = O.leg2alm(legs, spin, lmax, msel, ms0, theta_a, plan.nalm_loc).astype(np.complex128)
    print(f"# oracle adjoint {time.time()-t:.0f} s for m in {list(msel)}", flush=True)
    for k, m in enumerate(msel):
        lo = max(m, 2 if spin else 0)
        a = slice(mstart0[m] + lo, mstart0[m] + lmax + 1)
        print(f"   adjoint m={int(m):5d}: rel L2 {rel(out[:, a], refa[:, a]):.2e}")
