"""FP64 CPU port of the hot path (TEST INFRASTRUCTURE / CPU BASELINE ONLY).

Legendre stage: oracle/cpu_port.cpp (OpenMP, same formulation as the CUDA kernels).
Ring stage: oracle/cpu_port.cpp too (alias fold / phase shift of SURVEY rows A8/A9 + a real FFT per ring through a
half-length mixed-radix / Bluestein complex transform, OpenMP over rings, plans kept between calls).  Used by tests as a second, independent-in-precision
check of the formulation and by bench.py as the `cpu_baseline` / `--impl reference` arm
("kind": "port" — the real reference needs Julia + MPI + Ducc0, none of which exist here).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

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


def max_threads():
    """OpenMP threads the port will actually use (honours OMP_NUM_THREADS / set_threads)."""
    f = lib().cpu_max_threads
    f.restype = C.c_int
    return int(f())


def set_dd_start(on):
    """Experiment switch: start values sin(theta)^m in double-double arithmetic (off = what the kernels do)."""
    lib().cpu_set_dd_start(C.c_int(1 if on else 0))


def set_threads(n):
    lib().cpu_set_threads(C.c_int(int(n)))


def prepare_rings(nphi):
    """Build and keep the FFT plans of these (even) ring lengths: plan construction stays out of timed regions."""
    nphi = np.ascontiguousarray(nphi, dtype=i64)
    f = lib().cpu_prepare_rings
    f.restype = C.c_int
    if f(C.c_int64(len(nphi)), _p(nphi)):
        raise ValueError("ring lengths must be positive and even")


def leg2map(leg, nphi, phi0, rstart0, npix):
    """leg [ncomp, nring, nm] -> map [ncomp, npix] float64 (src/sht.jl:93); OpenMP over rings."""
    leg = np.ascontiguousarray(leg, dtype=np.complex128)
    ncomp, nring, nm = leg.shape
    nphi = np.ascontiguousarray(nphi, dtype=i64)
    phi0 = np.ascontiguousarray(phi0, dtype=np.float64)
    rstart0 = np.ascontiguousarray(rstart0, dtype=i64)
    out = np.zeros((ncomp, npix), dtype=np.float64)
    lib().cpu_leg2map(C.c_int(ncomp), C.c_int64(nm), C.c_int64(nring), _p(nphi), _p(phi0), _p(rstart0),
                      _p(leg), _p(out), C.c_int64(npix))
    return out


def map2leg(mp, nm, nphi, phi0, rstart0):
    """map [ncomp, npix] -> leg [ncomp, nring, nm] complex128 (src/sht.jl:169); OpenMP over rings."""
    mp = np.ascontiguousarray(mp, dtype=np.float64)
    ncomp = mp.shape[0]
    nphi = np.ascontiguousarray(nphi, dtype=i64)
    phi0 = np.ascontiguousarray(phi0, dtype=np.float64)
    rstart0 = np.ascontiguousarray(rstart0, dtype=i64)
    nring = len(nphi)
    leg = np.zeros((ncomp, nring, nm), dtype=np.complex128)
    lib().cpu_map2leg(C.c_int(ncomp), C.c_int64(nm), C.c_int64(nring), _p(nphi), _p(phi0), _p(rstart0),
                      _p(mp), C.c_int64(mp.shape[1]), _p(leg))
    return leg
