"""DAlm / AlmInfoMPI and the round-robin alm bookkeeping (reference src/alm.jl).

Index conventions follow Julia where the reference stores them: ``info.mstart`` is the 1-BASED
offset of the hypothetical l=0 coefficient (src/alm.jl:141-151), exactly what the reference hands
to Ducc0.jl; the binding subtracts 1 at the C boundary like Ducc0.jl's glue does.
``DAlm.alm`` has Julia's shape ``(nalm_loc, ncomp)`` in column-major order (numpy order='F')."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .comm import COMM_SELF, Comm
from .healpix import Alm
from .strategy import RR, Strategy


# ---- RR bookkeeping: thin wrappers over the integer-exact C restatement -----------------------
def get_nm_RR(global_mmax: int, task_rank: int, c_size: int) -> int:
    """src/alm.jl:102-105"""
    out = C.c_int64()
    _lib.check(_lib.lib().hpsht_rr_nm(global_mmax, task_rank, c_size, C.byref(out)))
    return out.value


def get_mval_RR(global_mmax: int, task_rank: int, c_size: int) -> np.ndarray:
    """src/alm.jl:112-119"""
    nm = get_nm_RR(global_mmax, task_rank, c_size)
    out = np.zeros(nm, dtype=np.int64)
    _lib.check(_lib.lib().hpsht_rr_mval(global_mmax, task_rank, c_size, _lib.ptr(out)))
    return out


def get_m_tasks_RR(mmax: int, c_size: int) -> np.ndarray:
    """src/alm.jl:127-133"""
    out = np.zeros(mmax + 1, dtype=np.int64)
    _lib.check(_lib.lib().hpsht_rr_m_tasks(mmax, c_size, _lib.ptr(out)))
    return out


def make_mstart_complex(lmax: int, stride: int, mval) -> np.ndarray:
    """1-based mstart as in Julia (src/alm.jl:141-151)."""
    mval = _lib.as_i64(mval)
    out = np.zeros(len(mval), dtype=np.int64)
    _lib.check(_lib.lib().hpsht_rr_mstart(lmax, stride, len(mval), _lib.ptr(mval), _lib.ptr(out)))
    return out + stride  # the C side is 0-based


class AlmInfoMPI:
    """src/alm.jl:30-45"""

    def __init__(self, comm: Comm = COMM_SELF, lmax=0, mmax=0, maxnm=0, mval=None, mstart=None):
        self.comm = comm
        self.lmax = lmax
        self.mmax = mmax
        self.maxnm = maxnm
        self.mval = np.zeros(0, dtype=np.int64) if mval is None else _lib.as_i64(mval)
        self.mstart = np.zeros(0, dtype=np.int64) if mstart is None else _lib.as_i64(mstart)


class DAlm:
    """DAlm{S,T} (src/alm.jl:75-93): ``alm`` is (nalm_loc, ncomp) complex128, column-major.  It may
    also be a device array (anything with .data_ptr(), e.g. a torch CUDA tensor of shape
    (ncomp, nalm_loc) complex128) for device-resident use."""

    def __init__(self, comm: Comm = COMM_SELF, strategy: type = RR, alm=None, info: AlmInfoMPI | None = None):
        if not (isinstance(strategy, type) and issubclass(strategy, Strategy)):
            raise TypeError("strategy must be a Strategy subclass")
        self.strategy = strategy
        self.alm = np.zeros((0, 1), dtype=np.complex128, order="F") if alm is None else alm
        self.info = AlmInfoMPI(comm) if info is None else info

    @property
    def ncomp(self) -> int:
        a = self.alm
        if isinstance(a, np.ndarray):
            return a.shape[1]
        return int(a.shape[0])  # device arrays are (ncomp, nalm)

    def size(self):
        return self.alm.shape


def ScatterAlm_(alm, d_alm: DAlm) -> None:
    """ScatterAlm! (src/alm.jl:159-186): every rank holds the full Alm (or [E, B] list) and slices
    its own m's."""
    if d_alm.strategy is not RR:
        raise NotImplementedError("only the RR strategy exists")
    alms = list(alm) if isinstance(alm, (list, tuple)) else [alm]
    comm = d_alm.info.comm
    c_rank, c_size = comm.rank(), comm.size()
    a0 = alms[0]
    mval = get_mval_RR(a0.mmax, c_rank, c_size)
    if len(mval) == 0:
        raise _lib.DomainError(_lib.HPSHT_ERR_EMPTY_RANK, f"{c_rank}-th task is empty.")
    idx = a0.each_ell_idx(mval)
    cols = [a.alm[idx] for a in alms[:2]]
    d_alm.alm = np.asfortranarray(np.stack(cols, axis=1))
    d_alm.info.lmax = a0.lmax
    d_alm.info.mmax = a0.mmax
    d_alm.info.maxnm = get_nm_RR(a0.mmax, 0, c_size)
    d_alm.info.mval = mval
    d_alm.info.mstart = make_mstart_complex(a0.lmax, 1, mval)


def Scatter_alm_(in_alm, out_d_alm: DAlm, out_d_pol_alm: DAlm | None = None, comm: Comm | None = None,
                 root: int = 0, clear: bool = False) -> None:
    """MPI.Scatter!(in_alm, out_d_alm[, out_d_pol_alm][, comm]; root, clear) (src/alm.jl:219-307).
    ``in_alm`` may be None on non-root ranks.  With two outputs, in_alm = [T, E, B] and T goes to
    the first (spin-0) DAlm, (E, B) to the second."""
    if comm is not None:
        out_d_alm.info.comm = comm
        if out_d_pol_alm is not None:
            out_d_pol_alm.info.comm = comm
    c = out_d_alm.info.comm
    if c.rank() == root and in_alm is None:
        raise _lib.DomainError(_lib.HPSHT_ERR_ARG, "Input alm on root task can not be `nothing`.")
    in_alm = c.bcast(in_alm, root)
    if out_d_pol_alm is None:
        ScatterAlm_(in_alm, out_d_alm)
    else:
        if len(in_alm) < 3:
            raise _lib.DomainError(_lib.HPSHT_ERR_ARG, "Input alm vector must have at least three components.")
        ScatterAlm_(in_alm[0], out_d_alm)
        ScatterAlm_(in_alm[1:3], out_d_pol_alm)


def Gather_alm_(in_d_alm: DAlm, out_alm: Alm | None, comp: int = 1, root: int = 0, allgather: bool = False):
    """MPI.Gather!/Allgather!(in_d_alm, out_alm, comp; root) (src/alm.jl:313-595).  ``comp`` is
    1-based as in Julia.  One exchange instead of the reference's maxnm rounds; the result (every
    a_lm placed at its m-major position) is identical."""
    if in_d_alm.ncomp < comp:
        raise _lib.DomainError(_lib.HPSHT_ERR_ARG, "not enough components in DAlm")
    comm = in_d_alm.info.comm
    info = in_d_alm.info
    piece = (info.mval, np.ascontiguousarray(in_d_alm.alm[:, comp - 1]))
    parts = comm.allgather(piece) if allgather else comm.gather(piece, root)
    if parts is None:
        return
    if out_alm is None:
        raise _lib.DomainError(_lib.HPSHT_ERR_ARG, "output alm can not be `nothing` on a receiving task")
    if out_alm.lmax != info.lmax:
        raise _lib.DomainError(_lib.HPSHT_ERR_ARG, "lmax must match")
    for mval, vals in parts:
        out_alm.alm[out_alm.each_ell_idx(mval)] = vals


def localdot(a: DAlm, b: DAlm, comp1: int = 1, comp2: int = 1) -> float:
    """src/alm.jl:621-646: weight 1 for m = 0 (real parts only), 2 for m > 0."""
    if a.info.lmax != b.info.lmax or not np.array_equal(a.info.mval, b.info.mval) or \
            not np.array_equal(a.info.mstart, b.info.mstart):
        raise _lib.DomainError(_lib.HPSHT_ERR_ARG, "lmax, mval and mstart must match")
    lmax = a.info.lmax
    res0 = 0.0
    rest = 0.0
    x, y = a.alm[:, comp1 - 1], b.alm[:, comp2 - 1]
    for mi, m in enumerate(a.info.mval):
        i1 = a.info.mstart[mi] + m - 1
        i2 = i1 + lmax - m + 1
        if m == 0:
            res0 += float(np.dot(x[i1:i2].real, y[i1:i2].real))
        else:
            rest += float(np.dot(x[i1:i2].real, y[i1:i2].real) + np.dot(x[i1:i2].imag, y[i1:i2].imag))
    return res0 + 2 * rest


def dot(a: DAlm, b: DAlm, comp1: int = 1, comp2: int = 1) -> float:
    """src/alm.jl:655-660"""
    if a.info.comm != b.info.comm:
        raise _lib.DomainError(_lib.HPSHT_ERR_ARG, "Communicators must match")
    return float(a.info.comm.allreduce_sum(localdot(a, b, comp1, comp2)))
