"""Minimal serial Healpix.jl objects the distributed API is expressed in (Healpix.jl is an
un-vendored upstream dependency of the reference: Project.toml:10): Alm, HealpixMap (RING order),
PolarizedHealpixMap, Resolution helpers.  Geometry comes from the C library (hpsht_healpix_ring)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


def numberOfAlms(lmax: int, mmax: int | None = None) -> int:
    mmax = lmax if mmax is None else mmax
    return (mmax + 1) * (lmax + 1) - (mmax * (mmax + 1)) // 2


def numOfRings(nside: int) -> int:
    return 4 * nside - 1


def getEquatorIdx(nside: int) -> int:
    return 2 * nside


def nside2npix(nside: int) -> int:
    return 12 * nside * nside


class Alm:
    """Healpix.Alm{ComplexF64}: a_lm for m >= 0, m-major: index(l, m) = m(2 lmax + 1 - m)/2 + l
    (0-based) when mmax == lmax."""

    def __init__(self, lmax: int, mmax: int | None = None, alm=None):
        self.lmax = int(lmax)
        self.mmax = int(lmax if mmax is None else mmax)
        n = numberOfAlms(self.lmax, self.mmax)
        self.alm = np.zeros(n, dtype=np.complex128) if alm is None else np.asarray(alm, dtype=np.complex128)
        if self.alm.shape != (n,):
            raise ValueError("wrong number of alm coefficients")

    def index(self, l, m):
        return m * (2 * self.lmax + 1 - m) // 2 + l

    def each_ell_idx(self, mval):
        """0-based indices of the runs l = m..lmax for each m in mval (Healpix.each_ell_idx)."""
        return np.concatenate([self.index(np.arange(m, self.lmax + 1), m) for m in mval]) \
            if len(mval) else np.zeros(0, dtype=np.int64)


class RingInfo:
    __slots__ = ("ring", "firstPixIdx", "numOfPixels", "colatitude_rad", "phi0")


def getringinfo(nside: int, ring: int) -> RingInfo:
    """Healpix.getringinfo! for 1-based ring; firstPixIdx is 1-based as in Julia."""
    L = _lib.lib()
    nphi = C.c_int64()
    th = C.c_double()
    p0 = C.c_double()
    fp = C.c_int64()
    _lib.check(L.hpsht_healpix_ring(nside, ring, C.byref(nphi), C.byref(th), C.byref(p0), C.byref(fp)))
    r = RingInfo()
    r.ring = ring
    r.firstPixIdx = fp.value + 1
    r.numOfPixels = nphi.value
    r.colatitude_rad = th.value
    r.phi0 = p0.value
    return r


def ring2theta(nside: int) -> np.ndarray:
    return np.array([getringinfo(nside, r).colatitude_rad for r in range(1, 4 * nside)])


class HealpixMap:
    """Healpix.HealpixMap{Float64, RingOrder}."""

    def __init__(self, nside: int, pixels=None):
        self.nside = int(nside)
        n = nside2npix(self.nside)
        self.pixels = np.zeros(n, dtype=np.float64) if pixels is None else np.asarray(pixels, dtype=np.float64)
        if self.pixels.shape != (n,):
            raise ValueError("wrong number of pixels")


class PolarizedHealpixMap:
    """Healpix.PolarizedHealpixMap{Float64, RingOrder}: fields i, q, u."""

    def __init__(self, nside: int, i=None, q=None, u=None):
        self.nside = int(nside)
        self.i = HealpixMap(nside, i)
        self.q = HealpixMap(nside, q)
        self.u = HealpixMap(nside, u)
