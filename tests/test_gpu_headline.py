"""Parity at the BASELINE sizes against the long-double oracle, with explicit per-class thresholds
(VERDICT round 1, item 1a): sampled rings (all m) of alm2map! and sampled m rows (all rings) of
adjoint_alm2map!, through the fused C-ABI plan with host buffers (the ring-block pipeline).

Classes: bulk rings (>= 64 pixels), the equator neighbourhood (|ring - 2 nside| <= 8, where the two-step spin-0
recurrence is ill-conditioned), polar rings (< 64 pixels, conditioning eps l / sin(theta)), adjoint rows of low
m (fed by the polar rings) and of high m.  north_star tolerance: relative L2 <= 1e-12.
"""
import numpy as np
import pytest

from oracle import parity as PR

pytestmark = pytest.mark.gpu

TOL_RING = 1e-12       # every sampled ring with >= 64 pixels, equator rings included
TOL_AGG = 5e-13        # relative L2 over all sampled pixels of those rings
TOL_POLAR = 1e-12      # rings with < 64 pixels (4 .. 60 pixels each); the oracle runs them in __float128
TOL_ADJ_ROW = 1e-12    # every sampled adjoint row, m <= 10 included


@pytest.mark.parametrize("nside,lmax,spin", [(4096, 8191, 0), (4096, 8191, 2), (2048, 6143, 2), (1024, 3071, 0)])
def test_sampled_parity_at_baseline_sizes(H, O, nside, lmax, spin):
    ncomp = 1 if spin == 0 else 2
    rng = np.random.default_rng(4096 + spin)
    plan = H.Plan(nside, lmax, H.COMM_SELF, ncomp)
    mval, mstart0 = plan.alm_info()
    info = plan.map_info()
    alm = (rng.standard_normal((ncomp, plan.nalm_loc)) + 1j * rng.standard_normal((ncomp, plan.nalm_loc))) * np.sqrt(0.5)
    mp = np.zeros((ncomp, plan.npix_loc))
    plan.alm2map(alm, mp, ncomp)
    loc = {int(r): (int(s), int(n)) for r, s, n in zip(info["rings"], info["rstart0"], info["nphi"])}
    sel = PR.sample_rings(nside, info["rings"])
    got = {int(r): mp[:, loc[int(r)][0]:loc[int(r)][0] + loc[int(r)][1]] for r in sel}
    res = PR.synthesis_parity(nside, lmax, spin, alm, got, sel)
    raw = PR.synthesis_parity(nside, lmax, spin, alm, got, sel, pair_mirror=False)
    print(f"\n   against the southern rings' own doubles (input rounding l*ulp(pi)/2 included): rel L2 "
          f"{raw['rel_l2']:.2e}, worst ring {raw['max_ring']:.2e}")
    print(f"\nalm2map nside={nside} lmax={lmax} spin={spin}: rel L2 {res['rel_l2']:.2e}, worst ring "
          f"{res['max_ring']:.2e}, equator {res['max_equator_ring']:.2e}, polar {res['max_polar_ring']:.2e} "
          f"({res['n_rings']} rings, {res['n_pixels']} px)")
    worst = sorted(zip(res["per_ring"], res["rings"]), reverse=True)[:6]
    print("   worst rings:", [(r, float(f"{e:.2e}")) for e, r in worst])
    fails = []
    if res["rel_l2"] > TOL_AGG:
        fails.append(("aggregate", res["rel_l2"]))
    if res["max_ring"] > TOL_RING:
        fails.append(("ring", worst))
    if res["max_polar_ring"] > TOL_POLAR:
        fails.append(("polar", res["max_polar_ring"]))
    if raw["max_ring"] > TOL_RING + 1.2e-16 * lmax * 2:   # + the input rounding of theta_S
        fails.append(("ring, own doubles", raw["max_ring"]))

    f = rng.standard_normal((ncomp, plan.npix_loc))
    out = np.zeros((ncomp, plan.nalm_loc), dtype=np.complex128)
    plan.adjoint_alm2map(f, out, ncomp)
    msel = [0, 1, 2, 3, 10, 100, 1000, lmax // 2, lmax - 100, lmax - 1, lmax]
    rings_all = [f[:, s:s + n] for s, n in zip(info["rstart0"], info["nphi"])]
    rows = {int(m): out[:, mstart0[m] + m: mstart0[m] + lmax + 1] for m in msel}
    adj = PR.adjoint_parity(nside, lmax, spin, rings_all, info["rings"], rows, msel)
    print("adjoint rows:", {m: float(f"{v:.2e}") for m, v in adj.items()})
    plan.destroy()
    if max(adj.values()) > TOL_ADJ_ROW:
        fails.append(("adjoint rows", adj))
    assert not fails, fails
