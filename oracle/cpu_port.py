"""FP64 CPU port of the hot path (TEST INFRASTRUCTURE / CPU BASELINE ONLY).

Legendre stage: oracle/cpu_port.cpp (OpenMP, same formulation as the CUDA kernels).
Ring-FFT stage: scipy.fft (pocketfft — the same FFT family ducc0 uses) with the alias fold /
phase shift of SURVEY row A8/A9 in numpy.  Used by tests as a second, independent-in-precision
check of the formulation and by bench.py as the `cpu_baseline` / `--impl reference` arm
("kind": "port" — the real reference needs Julia + MPI + Ducc0, none of which exist here).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import scipy.fft

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
i64 = np.int64


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libcpuport.so")
        if not os.path.exists(path):
            from . import oracle as _o
            _o.build()
        _LIB = C.CDLL(path)
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def legendre_steps(spin, lmax, mval, theta):
    mval = np.ascontiguousarray(mval, dtype=i64)
    theta = np.ascontiguousarray(theta, dtype=np.float64)
    c = C.c_double()
    d = C.c_double()
    lib().cpu_legendre_steps(C.c_int(spin), C.c_int64(lmax), C.c_int64(len(mval)), _p(mval),
                             C.c_int64(len(theta)), _p(theta), C.byref(c), C.byref(d))
    return c.value, d.value


def alm2leg(alm, spin, lmax, mval, mstart0, theta):
    """alm [ncomp, nalm] -> leg [ncomp, nring, nm] complex128 (src/sht.jl:85)."""
    alm = np.ascontiguousarray(alm, dtype=np.complex128)
    ncomp = 1 if spin == 0 else 2
    assert alm.shape[0] == ncomp
    mval = np.ascontiguousarray(mval, dtype=i64)
    mstart0 = np.ascontiguousarray(mstart0, dtype=i64)
    theta = np.ascontiguousarray(theta, dtype=np.float64)
    leg = np.zeros((ncomp, len(theta), len(mval)), dtype=np.complex128)
    lib().cpu_alm2leg(C.c_int(spin), C.c_int64(lmax), C.c_int64(len(mval)), _p(mval), _p(mstart0),
                      _p(alm), C.c_int64(alm.shape[1]), C.c_int64(len(theta)), _p(theta), _p(leg))
    return leg


def leg2alm(leg, spin, lmax, mval, mstart0, theta, nalm):
    """leg [ncomp, nring, nm] -> alm [ncomp, nalm] complex128 (src/sht.jl:178)."""
    leg = np.ascontiguousarray(leg, dtype=np.complex128)
    ncomp = 1 if spin == 0 else 2
    mval = np.ascontiguousarray(mval, dtype=i64)
    mstart0 = np.ascontiguousarray(mstart0, dtype=i64)
    theta = np.ascontiguousarray(theta, dtype=np.float64)
    assert leg.shape == (ncomp, len(theta), len(mval))
    alm = np.zeros((ncomp, nalm), dtype=np.complex128)
    lib().cpu_leg2alm(C.c_int(spin), C.c_int64(lmax), C.c_int64(len(mval)), _p(mval), _p(mstart0),
                      _p(leg), C.c_int64(len(theta)), _p(theta), _p(alm), C.c_int64(nalm))
    return alm


def _phase(nm, phi0):
    m = np.arange(nm, dtype=np.float64)
    return np.exp(1j * m * phi0)


def fold(leg_row, phi0, n):
    """half-complex spectrum of one ring from its leg row (SURVEY A8 fold rule)."""
    nm = len(leg_row)
    t = leg_row * _phase(nm, phi0)
    hc = np.zeros(n // 2 + 1, dtype=np.complex128)
    hc[0] = t[0].real
    m = np.arange(1, nm)
    i1 = m % n
    i2 = (n - i1) % n
    s1 = i1 <= n // 2
    np.add.at(hc, i1[s1], t[1:][s1])
    s2 = i2 <= n // 2
    np.add.at(hc, i2[s2], np.conj(t[1:][s2]))
    return hc


def leg2map(leg, nphi, phi0, rstart0, npix):
    """leg [ncomp, nring, nm] -> map [ncomp, npix] float64 (src/sht.jl:93)."""
    ncomp, nring, nm = leg.shape
    out = np.zeros((ncomp, npix), dtype=np.float64)
    for c in range(ncomp):
        for ir in range(nring):
            n = int(nphi[ir])
            hc = fold(leg[c, ir], phi0[ir], n)
            out[c, rstart0[ir]:rstart0[ir] + n] = scipy.fft.irfft(hc, n) * n
    return out


def map2leg(mp, nm, nphi, phi0, rstart0):
    """map [ncomp, npix] -> leg [ncomp, nring, nm] complex128 (src/sht.jl:169)."""
    ncomp = mp.shape[0]
    nring = len(nphi)
    leg = np.zeros((ncomp, nring, nm), dtype=np.complex128)
    m = np.arange(nm)
    for c in range(ncomp):
        for ir in range(nring):
            n = int(nphi[ir])
            X = scipy.fft.rfft(mp[c, rstart0[ir]:rstart0[ir] + n])
            k = m % n
            up = k > n // 2
            v = np.where(up, np.conj(X[np.where(up, n - k, 0)]), X[np.where(up, 0, k)])
            leg[c, ir] = v * np.conj(_phase(nm, phi0[ir]))
    return leg
