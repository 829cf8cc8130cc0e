"""Sampled parity checks of a plan against the extended-precision oracle (long double; __float128 recurrences on
the rings next to the poles) -- TEST INFRASTRUCTURE ONLY.

Used by tests/test_gpu_headline.py and by bench.py's `parity` block (after the timed region).  The oracle
applies the reference backend's m > mlim(theta) ring cut so that the comparison measures the
arithmetic, not the cut (profiles/r01_parity.md explains the difference at lmax = 8191).
"""
from __future__ import annotations

import numpy as np

from . import oracle as O


def rel(a, b) -> float:
    den = float(np.sqrt(np.sum(np.abs(b) ** 2)))
    return float(np.sqrt(np.sum(np.abs(a - b) ** 2))) / (den if den > 0 else 1.0)


def sample_rings(nside: int, own_rings, n_bulk: int = 12, seed: int = 99):
    """1-based ring ids drawn from `own_rings` (the rank's rings): the ones closest to both poles, the equator
    neighbourhood, the cap/belt boundary and a pixel-weighted random sample; north and south."""
    own = np.sort(np.asarray(own_rings, dtype=np.int64))
    nr = 4 * nside - 1
    want = [1, 2, 3, 5, 16, 64, nside, 2 * nside - 2, 2 * nside - 1, 2 * nside]
    rs = np.random.default_rng(seed)
    want += [int(x) for x in np.sqrt(rs.uniform(1, nside ** 2, n_bulk // 2))]
    want += [int(x) for x in rs.uniform(nside, 2 * nside, n_bulk - n_bulk // 2)]
    want = want + [nr + 1 - r for r in want]
    # snap every wanted ring to the nearest ring this rank owns
    idx = np.clip(np.searchsorted(own, want), 0, len(own) - 1)
    idx2 = np.clip(idx - 1, 0, len(own) - 1)
    pick = np.where(np.abs(own[idx] - want) <= np.abs(own[idx2] - want), own[idx], own[idx2])
    return np.unique(pick)


def _parity_sign(lmax):
    """(-1)^(l+m) in the 1-rank local alm order (m-major, l = m..lmax)."""
    return np.concatenate([np.where((np.arange(m, lmax + 1) + m) % 2 == 0, 1.0, -1.0) for m in range(lmax + 1)])


def synthesis_parity(nside, lmax, spin, alm_full, got_rings: dict, rings, pair_mirror: bool = True):
    """alm_full [ncomp, nalm] in the 1-rank local order (m-major, l = m..lmax) ; got_rings[ring] = [ncomp, nphi]
    pixels the implementation produced.  Returns a dict of per-class relative L2 errors against the oracle.

    pair_mirror: evaluate a southern ring at the exact mirror image pi - theta_N of its northern partner's
    double (through lambda(pi - theta) = (-1)^(l+m) lambda(theta), resp. the +-2 exchange for spin 2) instead
    of at its own double theta_S.  Ring-pairing backends (libsharp / ducc, and this library) evaluate both
    rings of a pair from one colatitude; the caller's theta_S is pi - theta_N rounded once more, i.e. off by
    up to ulp(pi)/2 = 2.2e-16, which moves lambda_lm by l * 2.2e-16 = 1.8e-12 at l = 8191 -- input rounding,
    not arithmetic.  pair_mirror = False measures against the southern rings' own doubles."""
    rings = np.asarray(rings, dtype=np.int64)
    nphi, theta, phi0, _ = O.healpix_rings(nside, rings)
    mval = np.arange(lmax + 1, dtype=np.int64)
    mstart0 = O.rr_mstart1(lmax, 1, mval) - 1
    O.set_reference_cut(True)
    try:
        leg = O.alm2leg(alm_full, spin, lmax, mval, mstart0, theta)
        south = rings > 2 * nside
        if pair_mirror and south.any():
            _, th_n, _, _ = O.healpix_rings(nside, 4 * nside - rings[south])
            sg = _parity_sign(lmax)
            almm = np.array(alm_full, dtype=np.complex128) * sg[None, :]
            if spin:
                almm[1] = -almm[1]
            legs = O.alm2leg(almm, spin, lmax, mval, mstart0, th_n)
            if spin:
                legs[1] = -legs[1]
            leg[:, south, :] = legs
    finally:
        O.set_reference_cut(False)
    rs0 = np.concatenate([[0], np.cumsum(nphi)[:-1]]).astype(np.int64)
    ref = O.leg2map(leg, nphi, phi0, rs0, int(nphi.sum())).astype(np.float64)
    per, num, den = [], [], []
    for r, a, n in zip(rings, rs0, nphi):
        g = np.asarray(got_rings[int(r)], dtype=np.float64)
        x = ref[:, a:a + n]
        num.append(float(np.sum((g - x) ** 2)))
        den.append(float(np.sum(x ** 2)))
        per.append(np.sqrt(num[-1] / den[-1]))
    per, num, den = np.array(per), np.array(num), np.array(den)
    n_ = np.minimum(rings, 4 * nside - rings)
    polar = nphi < 64
    equator = np.abs(rings - 2 * nside) <= 8
    bulk = ~polar
    out = {
        "rings": [int(r) for r in rings],
        "per_ring": [float(p) for p in per],
        "rel_l2": float(np.sqrt(num[bulk].sum() / den[bulk].sum())) if bulk.any() else None,
        "max_ring": float(per[bulk].max()) if bulk.any() else None,
        "max_polar_ring": float(per[polar].max()) if polar.any() else None,
        "max_equator_ring": float(per[equator].max()) if equator.any() else None,
        "n_rings": int(len(rings)), "n_pixels": int(nphi.sum()),
    }
    del n_
    return out


def adjoint_parity(nside, lmax, spin, map_full_rings, ring_ids, got_alm_rows: dict, msel, pair_mirror: bool = True):
    """map_full_rings: list of [ncomp, nphi] arrays for ALL rings `ring_ids` (1-based, any order) ;
    got_alm_rows[m] = [ncomp, lmax+1-m] complex rows the implementation produced.  The sampled-m legs are
    computed with a long-double real FFT (scipy's pocketfft, independent of the implementation)."""
    import scipy.fft

    ncomp = 1 if spin == 0 else 2
    ring_ids = np.asarray(ring_ids, dtype=np.int64)
    nphi, theta, phi0, _ = O.healpix_rings(nside, ring_ids)
    msel = np.asarray(sorted(set(int(m) for m in msel)), dtype=np.int64)
    legs = np.zeros((ncomp, len(ring_ids), len(msel)), dtype=np.clongdouble)
    for ir, n in enumerate(nphi):
        n = int(n)
        k = msel % n
        up = k > n // 2
        ang = msel.astype(np.longdouble) * np.longdouble(phi0[ir])
        ph = np.cos(ang) - 1j * np.sin(ang)
        for c in range(ncomp):
            X = scipy.fft.rfft(np.asarray(map_full_rings[ir][c], dtype=np.longdouble))
            v = np.where(up, np.conj(X[np.where(up, n - k, 0)]), X[np.where(up, 0, k)])
            legs[c, ir, :] = v * ph
    ms0 = O.rr_mstart1(lmax, 1, np.arange(lmax + 1))[msel] - 1
    nalm = (lmax + 1) * (lmax + 2) // 2
    O.set_reference_cut(True)
    try:
        south = ring_ids > 2 * nside
        if pair_mirror and south.any():
            # southern rings enter through the mirror image of their northern partner (see synthesis_parity)
            refa = O.leg2alm(legs[:, ~south, :], spin, lmax, msel, ms0, theta[~south], nalm)
            _, th_n, _, _ = O.healpix_rings(nside, 4 * nside - ring_ids[south])
            ls = legs[:, south, :].copy()
            if spin:
                ls[1] = -ls[1]
            rs = O.leg2alm(ls, spin, lmax, msel, ms0, th_n, nalm)
            for k, m in enumerate(msel):
                l = np.arange(int(m), lmax + 1)
                sg = np.where((l + int(m)) % 2 == 0, 1.0, -1.0)
                a = slice(ms0[k] + int(m), ms0[k] + lmax + 1)
                refa[0, a] += sg * rs[0, a]
                if spin:
                    refa[1, a] -= sg * rs[1, a]
            refa = refa.astype(np.complex128)
        else:
            refa = O.leg2alm(legs, spin, lmax, msel, ms0, theta, nalm).astype(np.complex128)
    finally:
        O.set_reference_cut(False)
    rows = {}
    for k, m in enumerate(msel):
        lo = max(int(m), 2 if spin else 0)
        if lo > lmax:
            continue
        ref = refa[:, ms0[k] + lo: ms0[k] + lmax + 1]
        got = np.asarray(got_alm_rows[int(m)])[:, lo - int(m):]
        rows[int(m)] = rel(got, ref)
    return rows
