"""DMap / GeomInfoMPI and the round-robin ring bookkeeping (reference src/map.jl).

``info.rings`` are 1-based global ring ids and ``info.rstart`` 1-based pixel offsets, as in Julia
(src/map.jl:19-37); ``DMap.pixels`` has shape (npix_loc, ncomp), column-major."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .comm import COMM_SELF, Comm
from .healpix import HealpixMap, PolarizedHealpixMap, getEquatorIdx, getringinfo, numOfRings, ring2theta
from .strategy import RR, Strategy


def get_nrings_RR(eq_idx: int, task_rank: int, c_size: int) -> int:
    """src/map.jl:99-102"""
    out = C.c_int64()
    _lib.check(_lib.lib().hpsht_rr_nrings(eq_idx, task_rank, c_size, C.byref(out)))
    return out.value


def get_rindexes_RR(nside_or_nrings: int, *args) -> np.ndarray:
    """get_rindexes_RR(nside, t_rank, c_size) or get_rindexes_RR(local_nrings, eq_idx, t_rank, c_size)
    (src/map.jl:113-128): 1-based ring ids, equator outwards, north before south."""
    if len(args) == 2:
        eq_idx, (t_rank, c_size) = getEquatorIdx(nside_or_nrings), args
    else:
        eq_idx, t_rank, c_size = args
    n = get_nrings_RR(eq_idx, t_rank, c_size)
    out = np.zeros(n, dtype=np.int64)
    _lib.check(_lib.lib().hpsht_rr_rindexes(eq_idx, t_rank, c_size, _lib.ptr(out)))
    return out


def get_rindexes_tot_RR(eq_index: int, c_size: int) -> np.ndarray:
    """src/map.jl:134-143"""
    out = np.zeros(2 * eq_index - 1, dtype=np.int64)
    _lib.check(_lib.lib().hpsht_rr_rindexes_tot(eq_index, c_size, _lib.ptr(out)))
    return out


class GeomInfoMPI:
    """src/map.jl:19-37"""

    def __init__(self, comm: Comm = COMM_SELF):
        self.comm = comm
        self.nside = 0
        self.maxnr = 0
        self.thetatot = np.zeros(0)
        self.rings = np.zeros(0, dtype=np.int64)
        self.rstart = np.zeros(0, dtype=np.int64)
        self.nphi = np.zeros(0, dtype=np.int64)
        self.theta = np.zeros(0)
        self.phi0 = np.zeros(0)


class DMap:
    """DMap{S,T} (src/map.jl:73-90).  ``pixels`` may also be a device array of shape
    (ncomp, npix_loc) float64 with .data_ptr()."""

    def __init__(self, comm: Comm = COMM_SELF, strategy: type = RR, pixels=None, info: GeomInfoMPI | None = None):
        if not (isinstance(strategy, type) and issubclass(strategy, Strategy)):
            raise TypeError("strategy must be a Strategy subclass")
        self.strategy = strategy
        self.pixels = np.zeros((0, 1), dtype=np.float64, order="F") if pixels is None else pixels
        self.info = GeomInfoMPI(comm) if info is None else info

    @property
    def ncomp(self) -> int:
        p = self.pixels
        if isinstance(p, np.ndarray):
            return p.shape[1]
        return int(p.shape[0])

    def size(self):
        return self.pixels.shape


def _fill_info(info: GeomInfoMPI, nside: int) -> None:
    comm = info.comm
    c_rank, c_size = comm.rank(), comm.size()
    eq = getEquatorIdx(nside)
    nrings = get_nrings_RR(eq, c_rank, c_size)
    if nrings == 0:
        raise _lib.DomainError(_lib.HPSHT_ERR_EMPTY_RANK, f"{c_rank}-th MPI task has no rings.")
    rings = get_rindexes_RR(nrings, eq, c_rank, c_size)
    ri = [getringinfo(nside, int(r)) for r in rings]
    info.nside = nside
    info.maxnr = get_nrings_RR(eq, 0, c_size) + 1
    info.rings = rings
    info.nphi = np.array([r.numOfPixels for r in ri], dtype=np.int64)
    info.rstart = np.concatenate([[1], 1 + np.cumsum(info.nphi)[:-1]]).astype(np.int64)
    info.theta = np.array([r.colatitude_rad for r in ri])
    info.phi0 = np.array([r.phi0 for r in ri])
    info.thetatot = ring2theta(nside)[get_rindexes_tot_RR(eq, c_size) - 1]
    info._first_pix0 = np.array([r.firstPixIdx - 1 for r in ri], dtype=np.int64)


def _local_pixel_index(info: GeomInfoMPI) -> np.ndarray:
    return np.concatenate([np.arange(f, f + n) for f, n in zip(info._first_pix0, info.nphi)])


def ScatterMap_(m, d_map: DMap) -> None:
    """ScatterMap! (src/map.jl:151-212): local rings in RR order, Q then U for polarised maps."""
    if d_map.strategy is not RR:
        raise NotImplementedError("only the RR strategy exists")
    if isinstance(m, PolarizedHealpixMap):
        nside, comps = m.nside, [m.q.pixels, m.u.pixels]
    else:
        nside, comps = m.nside, [m.pixels]
    _fill_info(d_map.info, nside)
    idx = _local_pixel_index(d_map.info)
    d_map.pixels = np.asfortranarray(np.stack([c[idx] for c in comps], axis=1))


def Scatter_map_(in_map, out_d_map: DMap, out_d_pol_map: DMap | None = None, comm: Comm | None = None,
                 root: int = 0, clear: bool = False) -> None:
    """MPI.Scatter!(in_map, out_d_map[, out_d_pol_map][, comm]; root, clear) (src/map.jl:245-330).
    With two outputs, a PolarizedHealpixMap's I goes to the first (spin-0) DMap and (Q, U) to the
    second."""
    if comm is not None:
        out_d_map.info.comm = comm
        if out_d_pol_map is not None:
            out_d_pol_map.info.comm = comm
    c = out_d_map.info.comm
    if c.rank() == root and in_map is None:
        raise _lib.DomainError(_lib.HPSHT_ERR_ARG, "Input map on root task can not be `nothing`.")
    in_map = c.bcast(in_map, root)
    if out_d_pol_map is None:
        ScatterMap_(in_map, out_d_map)
    else:
        ScatterMap_(in_map.i, out_d_map)
        ScatterMap_(in_map, out_d_pol_map)


def Gather_map_(in_d_map: DMap, out_map: HealpixMap | None, comp: int = 1, root: int = 0,
                allgather: bool = False) -> None:
    """MPI.Gather!/Allgather!(in_d_map, out_map, comp; root) (src/map.jl:337-559); comp 1-based."""
    if in_d_map.ncomp < comp:
        raise _lib.DomainError(_lib.HPSHT_ERR_ARG, "not enough components in DMap")
    comm = in_d_map.info.comm
    info = in_d_map.info
    piece = (_local_pixel_index(info), np.ascontiguousarray(in_d_map.pixels[:, comp - 1]))
    parts = comm.allgather(piece) if allgather else comm.gather(piece, root)
    if parts is None:
        return
    if out_map is None:
        raise _lib.DomainError(_lib.HPSHT_ERR_ARG, "output map can not be `nothing` on a receiving task")
    if out_map.nside != info.nside:
        raise _lib.DomainError(_lib.HPSHT_ERR_ARG, "nside must match")
    for idx, vals in parts:
        out_map.pixels[idx] = vals


def Scatter_array(arr, nside: int, comm: Comm, strategy: type = RR, root: int = 0) -> np.ndarray:
    """MPI.Scatter(arr, nside, comm; strategy, root) (src/tools.jl:41-58): ring-wise slice of a
    pixel-space vector, same pixel order as DMap.pixels."""
    arr = comm.bcast(arr, root)
    info = GeomInfoMPI(comm)
    _fill_info(info, nside)
    return np.asarray(arr)[_local_pixel_index(info)]
