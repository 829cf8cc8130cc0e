"""ctypes/numpy front-end of the CPU checker (TEST INFRASTRUCTURE ONLY).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product package never does.  See the header of sht_oracle.c for what
pins the oracle ("parity unpinned" at 1e-12: the reference holds no golden vectors).

Layouts follow the reference's Ducc0 seam (src/sht.jl:67-69): Julia column-major arrays, i.e.
numpy arrays indexed [comp, ring, m] (alm-side / map-side leg), [comp, i] (alm), [comp, pix].
Index arrays are 0-based here (Julia's mstart/rstart minus 1).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

i64 = np.int64
LD = np.longdouble
CLD = np.clongdouble


def build(force: bool = False) -> None:
    """Compile liborc.so / libcpuport.so with the committed Makefile."""
    args = ["make", "-s", "-C", _HERE]
    if force:
        args.append("-B")
    subprocess.check_call(args)


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liborc.so")
        if not os.path.exists(path):
            build()
        _LIB = C.CDLL(path)
        _LIB.orc_brute_sylm.restype = C.c_double
        _LIB.orc_brute_sylm.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double]
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def set_reference_cut(on: bool) -> None:
    """Apply (or not) the libsharp/ducc m > mlim(theta) ring-culling rule inside alm2leg / leg2alm, to
    separate an implementation's arithmetic error from the cut the reference's backend makes."""
    lib().orc_set_cut(C.c_int(1 if on else 0))


def _i64(a):
    return np.ascontiguousarray(a, dtype=i64)


# --------------------------------------------------------------------------------------------
# geometry
# --------------------------------------------------------------------------------------------
def healpix_ring(nside: int, ring: int):
    """(nphi, theta, phi0, first_pix0) of 1-based ring `ring` (src/map.jl:177-183 inputs)."""
    nphi = C.c_int64()
    th = C.c_double()
    p0 = C.c_double()
    fp = C.c_int64()
    lib().orc_healpix_ring(C.c_int64(nside), C.c_int64(ring), C.byref(nphi), C.byref(th),
                           C.byref(p0), C.byref(fp))
    return nphi.value, th.value, p0.value, fp.value


def healpix_rings(nside: int, rings):
    """Vectorised healpix_ring over 1-based ring ids -> (nphi[i64], theta, phi0, first_pix[i64])."""
    rings = np.asarray(rings, dtype=i64)
    out = [healpix_ring(nside, int(r)) for r in rings]
    nphi = np.array([o[0] for o in out], dtype=i64)
    theta = np.array([o[1] for o in out], dtype=np.float64)
    phi0 = np.array([o[2] for o in out], dtype=np.float64)
    fp = np.array([o[3] for o in out], dtype=i64)
    return nphi, theta, phi0, fp


# --------------------------------------------------------------------------------------------
# harmonics
# --------------------------------------------------------------------------------------------
def lam(spin: int, lmax: int, m: int, theta: float) -> np.ndarray:
    """{spin}lambda_lm(theta), l = 0..lmax, long double (spin in {0,+2,-2}, m >= 0)."""
    out = np.zeros(lmax + 1, dtype=LD)
    lib().orc_lambda(C.c_int(spin), C.c_int64(lmax), C.c_int64(m), C.c_double(theta), _p(out))
    return out


def lam_quad(spin: int, lmax: int, m: int, theta: float):
    """The same recurrence in __float128, returned as (hi, lo) double arrays; None if the start value leaves the
    quad range.  Only for checking lam() at large l (tests/test_oracle.py)."""
    hi = np.zeros(lmax + 1)
    lo = np.zeros(lmax + 1)
    f = lib().orc_lambda_quad
    f.restype = C.c_int
    rc = f(C.c_int(spin), C.c_int64(lmax), C.c_int64(m), C.c_double(theta), _p(hi), _p(lo))
    return None if rc else (hi, lo)


def brute_sylm(s: int, l: int, m: int, theta: float) -> float:
    return lib().orc_brute_sylm(s, l, m, theta)


# --------------------------------------------------------------------------------------------
# stages (long double in the middle)
# --------------------------------------------------------------------------------------------
def alm2leg(alm, spin, lmax, mval, mstart0, theta) -> np.ndarray:
    """alm [ncomp, nalm] complex128 -> leg [ncomp, nring, nm] clongdouble  (src/sht.jl:85)."""
    alm = np.ascontiguousarray(alm, dtype=np.complex128)
    ncomp = 1 if spin == 0 else 2
    assert alm.ndim == 2 and alm.shape[0] == ncomp
    mval = _i64(mval)
    mstart0 = _i64(mstart0)
    theta = np.ascontiguousarray(theta, dtype=np.float64)
    nm, nring = len(mval), len(theta)
    leg = np.zeros((ncomp, nring, nm), dtype=CLD)
    lib().orc_alm2leg(C.c_int(spin), C.c_int64(lmax), C.c_int64(nm), _p(mval), _p(mstart0),
                      _p(alm), C.c_int64(alm.shape[1]), C.c_int64(nring), _p(theta), _p(leg))
    return leg


def leg2alm(leg, spin, lmax, mval, mstart0, theta, nalm) -> np.ndarray:
    """leg [ncomp, nring, nm] -> alm [ncomp, nalm] clongdouble  (src/sht.jl:178)."""
    ncomp = 1 if spin == 0 else 2
    leg = np.ascontiguousarray(leg, dtype=CLD)
    mval = _i64(mval)
    mstart0 = _i64(mstart0)
    theta = np.ascontiguousarray(theta, dtype=np.float64)
    nm, nring = len(mval), len(theta)
    assert leg.shape == (ncomp, nring, nm)
    out = np.zeros((ncomp, nalm), dtype=CLD)
    lib().orc_leg2alm(C.c_int(spin), C.c_int64(lmax), C.c_int64(nm), _p(mval), _p(mstart0),
                      _p(leg), C.c_int64(nring), _p(theta), _p(out), C.c_int64(nalm))
    return out


def leg2map(leg, nphi, phi0, rstart0, npix, ring_sel=None) -> np.ndarray:
    """leg [ncomp, nring, nm] -> map [ncomp, npix] longdouble  (src/sht.jl:93)."""
    leg = np.ascontiguousarray(leg, dtype=CLD)
    ncomp, nring, nm = leg.shape
    nphi = _i64(nphi)
    rstart0 = _i64(rstart0)
    phi0 = np.ascontiguousarray(phi0, dtype=np.float64)
    out = np.zeros((ncomp, npix), dtype=LD)
    sel = None if ring_sel is None else np.ascontiguousarray(ring_sel, dtype=np.uint8)
    lib().orc_leg2map(C.c_int(ncomp), C.c_int64(nm), C.c_int64(nring), _p(nphi), _p(phi0),
                      _p(rstart0), _p(leg), _p(out), C.c_int64(npix),
                      _p(sel) if sel is not None else None)
    return out


def map2leg(mp, nm, nphi, phi0, rstart0, ring_sel=None) -> np.ndarray:
    """map [ncomp, npix] float64 -> leg [ncomp, nring, nm] clongdouble  (src/sht.jl:169)."""
    mp = np.ascontiguousarray(mp, dtype=np.float64)
    ncomp, npix = mp.shape
    nphi = _i64(nphi)
    rstart0 = _i64(rstart0)
    phi0 = np.ascontiguousarray(phi0, dtype=np.float64)
    nring = len(nphi)
    leg = np.zeros((ncomp, nring, nm), dtype=CLD)
    sel = None if ring_sel is None else np.ascontiguousarray(ring_sel, dtype=np.uint8)
    lib().orc_map2leg(C.c_int(ncomp), C.c_int64(nm), C.c_int64(nring), _p(nphi), _p(phi0),
                      _p(rstart0), _p(mp), C.c_int64(npix), _p(leg),
                      _p(sel) if sel is not None else None)
    return leg


def brute_synth(spin: int, lmax: int, alm_full, theta, phi) -> np.ndarray:
    """float128 brute-force synthesis at arbitrary points. alm_full [ncomp, nalm] in Healpix.jl
    Alm order (m-major).  Returns [ncomp, npts] float64."""
    ncomp = 1 if spin == 0 else 2
    alm_full = np.ascontiguousarray(alm_full, dtype=np.complex128)
    assert alm_full.shape == (ncomp, (lmax + 1) * (lmax + 2) // 2)
    theta = np.ascontiguousarray(theta, dtype=np.float64)
    phi = np.ascontiguousarray(phi, dtype=np.float64)
    out = np.zeros((ncomp, len(theta)), dtype=np.float64)
    lib().orc_brute_synth(C.c_int(spin), C.c_int(lmax), _p(alm_full), C.c_int64(len(theta)),
                          _p(theta), _p(phi), _p(out))
    return out


# --------------------------------------------------------------------------------------------
# Reference-shaped helpers: serial Healpix.jl objects and the RR scatter, restated in numpy.
# (The product has its own C implementation of the bookkeeping; this one is the checker.)
# --------------------------------------------------------------------------------------------
def alm_index(lmax: int, l, m):
    """0-based index into a Healpix.jl Alm.alm vector (m-major), mmax = lmax."""
    return m * (2 * lmax + 1 - m) // 2 + l


def rr_nm(mmax, rank, size):  # src/alm.jl:102-105
    if not rank < size:
        raise ValueError("rank can not exceed communicator size")
    return (mmax + size - rank) // size


def rr_mval(mmax, rank, size):  # src/alm.jl:112-119
    return np.array([rank + i * size for i in range(rr_nm(mmax, rank, size))], dtype=i64)


def rr_mstart1(lmax, stride, mval):  # src/alm.jl:141-151 (1-based, as in Julia)
    idx = 1
    out = []
    for m in mval:
        out.append(stride * (idx - int(m)))
        idx += lmax + 1 - int(m)
    return np.array(out, dtype=i64)


def rr_nrings(eq_idx, rank, size):  # src/map.jl:99-102
    if not rank < size:
        raise ValueError("rank can not exceed communicator size")
    return (eq_idx + size - 1 - rank) // size * 2 - (1 if rank == 0 else 0)


def rr_rindexes(eq_idx, rank, size):  # src/map.jl:113-121 (1-based ring ids)
    n = rr_nrings(eq_idx, rank, size)
    j = 0 if rank == 0 else 1
    out = []
    for i in range(1, n + 1):
        k = (i - j) // 2
        out.append(eq_idx - (-1) ** (i + j) * (rank + k * size))
    return np.array(out, dtype=i64)


def rr_rindexes_tot(eq_idx, size):  # src/map.jl:134-143
    return np.concatenate([rr_rindexes(eq_idx, r, size) for r in range(size)])


def scatter_alm(alm_full, lmax, rank, size):
    """What ScatterAlm! leaves in d_alm (src/alm.jl:159-186): (alm_loc [ncomp, nalm_loc],
    mval, mstart0)."""
    alm_full = np.atleast_2d(alm_full)
    mval = rr_mval(lmax, rank, size)
    if len(mval) == 0:
        raise ValueError("task is empty")
    idx = np.concatenate([alm_index(lmax, np.arange(m, lmax + 1), m) for m in mval])
    mstart0 = rr_mstart1(lmax, 1, mval) - 1
    return np.ascontiguousarray(alm_full[:, idx]), mval, mstart0


def scatter_map_info(nside, rank, size):
    """GeomInfoMPI arrays ScatterMap! builds (src/map.jl:151-197), 0-based rstart."""
    eq = 2 * nside
    rings = rr_rindexes(eq, rank, size)
    if len(rings) == 0:
        raise ValueError("task has no rings")
    nphi, theta, phi0, fp = healpix_rings(nside, rings)
    rstart0 = np.concatenate([[0], np.cumsum(nphi)[:-1]]).astype(i64)
    tot = rr_rindexes_tot(eq, size)
    _, thetatot, _, _ = healpix_rings(nside, tot)
    return dict(rings=rings, nphi=nphi, theta=theta, phi0=phi0, first_pix=fp, rstart0=rstart0,
                thetatot=thetatot, npix_loc=int(nphi.sum()))


def scatter_map(map_full, nside, rank, size):
    """(pixels_loc [ncomp, npix_loc], info) for a RING-ordered full map [ncomp, npix]."""
    map_full = np.atleast_2d(map_full)
    info = scatter_map_info(nside, rank, size)
    pix = np.concatenate([np.arange(f, f + n) for f, n in zip(info["first_pix"], info["nphi"])])
    return np.ascontiguousarray(map_full[:, pix]), info


def full_alm2map(alm_full, spin, nside, lmax, ring_sel_global=None):
    """Serial (1-rank) oracle synthesis: [ncomp, nalm] -> [ncomp, npix] long double, RING order.
    ring_sel_global: optional iterable of 1-based ring ids to restrict the (expensive) sums to."""
    nr = 4 * nside - 1
    rings = np.arange(1, nr + 1) if ring_sel_global is None else np.asarray(ring_sel_global)
    nphi, theta, phi0, fp = healpix_rings(nside, rings)
    mval = np.arange(lmax + 1, dtype=i64)
    mstart0 = rr_mstart1(lmax, 1, mval) - 1
    leg = alm2leg(np.atleast_2d(alm_full), spin, lmax, mval, mstart0, theta)
    return leg2map(leg, nphi, phi0, fp, 12 * nside * nside), (rings, nphi, fp)


def full_adjoint(map_full, spin, nside, lmax):
    """Serial oracle adjoint synthesis: [ncomp, npix] -> [ncomp, nalm] clongdouble."""
    nr = 4 * nside - 1
    rings = np.arange(1, nr + 1)
    nphi, theta, phi0, fp = healpix_rings(nside, rings)
    mval = np.arange(lmax + 1, dtype=i64)
    mstart0 = rr_mstart1(lmax, 1, mval) - 1
    leg = map2leg(np.atleast_2d(map_full), lmax + 1, nphi, phi0, fp)
    return leg2alm(leg, spin, lmax, mval, mstart0, theta, (lmax + 1) * (lmax + 2) // 2)
