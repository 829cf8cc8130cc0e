"""alm2map! / adjoint_alm2map! on DAlm{RR} / DMap{RR} (reference src/sht.jl), with the Ducc0 calls
replaced by the C-ABI CUDA library and MPI.Alltoallv! by NCCL.

Two forms, as in the reference:
  * 4-argument form with caller-provided leg buffers (src/sht.jl:75-94, 165-179): stage by stage —
    hpsht_alm2leg -> communicate_alm2map_ -> hpsht_leg2map — on host arrays; the library stages
    them through the GPU.  This is the literal drop-in of the reference code path.
  * 2-argument form (src/sht.jl:96-100, 181-185): the fused plan (hpsht_alm2map /
    hpsht_adjoint_alm2map) which owns the leg buffers and the NCCL exchange; data may be host
    (numpy) or device resident.
Julia's trailing ``!`` is spelled as a trailing underscore."""
from __future__ import annotations

import ctypes as C
import hashlib
import weakref

import numpy as np

from . import _lib
from .alm import DAlm, get_mval_RR, get_nm_RR
from .comm import Comm
from .healpix import numOfRings
from .map import DMap, get_nrings_RR
from .strategy import RR


# --------------------------------------------------------------------------------------------
# leg redistribution, reference-faithful (used by the 4-argument form and by the CPU tests)
# --------------------------------------------------------------------------------------------
def communicate_alm2map_(in_leg: np.ndarray, out_leg: np.ndarray, comm: Comm, strategy=RR) -> None:
    """communicate_alm2map! (src/sht.jl:5-44): in_leg (loc_nm, tot_nr, ncomp) -> out_leg
    (tot_nm, loc_nr, ncomp), both column-major complex128."""
    c_size = comm.size()
    tot_nm, loc_nr, ncomp_out = out_leg.shape
    loc_nm, tot_nr, ncomp_in = in_leg.shape
    if ncomp_in != ncomp_out:
        raise _lib.DomainError(_lib.HPSHT_ERR_ARG, "ncomp's of leg coefficinets must match")
    eq_index = (tot_nr + 1) // 2
    tot_mmax = tot_nm - 1
    send_counts = [loc_nm * get_nrings_RR(eq_index, t, c_size) for t in range(c_size)]
    rec_counts = [loc_nr * get_nm_RR(tot_mmax, t, c_size) for t in range(c_size)]
    cumrecs = np.concatenate([[0], np.cumsum(rec_counts)[:-1]])
    rec_array = np.empty(tot_nm * loc_nr, dtype=np.complex128)
    for comp in range(ncomp_in):
        # rings are stored grouped by destination rank: the send buffer is in_leg[:, :, comp] as is
        send = np.ascontiguousarray(in_leg[:, :, comp].reshape(-1, order="F"))
        comm.alltoallv(send, send_counts, rec_array, rec_counts)
        # unpack (src/sht.jl:35-42): source s's block is (its m's ascending, fastest) x (my rings)
        for s in range(c_size):
            nm_s = get_nm_RR(tot_mmax, s, c_size)
            blk = rec_array[cumrecs[s]:cumrecs[s] + rec_counts[s]].reshape((nm_s, loc_nr), order="F")
            out_leg[s::c_size, :, comp] = blk


def communicate_map2alm_(in_leg: np.ndarray, out_leg: np.ndarray, comm: Comm, strategy=RR) -> None:
    """communicate_map2alm! (src/sht.jl:106-135): in_leg (tot_nm, loc_nr, ncomp) -> out_leg
    (loc_nm, tot_nr, ncomp)."""
    c_size = comm.size()
    tot_nm, loc_nr, ncomp_in = in_leg.shape
    loc_nm, tot_nr, ncomp_out = out_leg.shape
    if ncomp_in != ncomp_out:
        raise _lib.DomainError(_lib.HPSHT_ERR_ARG, "ncomp's of leg coefficinets must match")
    eq_index = (tot_nr + 1) // 2
    tot_mmax = tot_nm - 1
    send_counts = [get_nm_RR(tot_mmax, t, c_size) * loc_nr for t in range(c_size)]
    rec_counts = [loc_nm * get_nrings_RR(eq_index, t, c_size) for t in range(c_size)]
    recv = np.empty(loc_nm * tot_nr, dtype=np.complex128)
    for comp in range(ncomp_in):
        # pack (src/sht.jl:120-130): rows mval_RR(t) of in_leg[:, :, comp], m fastest then ring
        send = np.concatenate([in_leg[t::c_size, :, comp].reshape(-1, order="F") for t in range(c_size)])
        comm.alltoallv(send, send_counts, recv, rec_counts)
        out_leg[:, :, comp] = recv.reshape((loc_nm, tot_nr), order="F")


# --------------------------------------------------------------------------------------------
# fused plan cache
# --------------------------------------------------------------------------------------------
class Plan:
    """Owner of one hpsht_plan (one per (nside, lmax, communicator, max_ncomp) and process)."""

    def __init__(self, nside: int, lmax: int, comm: Comm, max_ncomp: int = 2, device: int = -1,
                 thetatot=None, phi0=None):
        L = _lib.lib()
        self.nside, self.lmax, self.comm = nside, lmax, comm
        rank, size = comm.rank(), comm.size()
        uid = None
        if size > 1:
            buf = (C.c_ubyte * _lib.HPSHT_NCCL_ID_BYTES)()
            if rank == 0:
                _lib.check(L.hpsht_nccl_unique_id(C.addressof(buf)))
            uid = comm.bcast(bytes(buf), 0)
            uid = (C.c_ubyte * _lib.HPSHT_NCCL_ID_BYTES).from_buffer_copy(uid)
        h = C.c_void_p()
        _lib.check(L.hpsht_plan_create(C.byref(h), nside, lmax, rank, size,
                                       C.addressof(uid) if uid is not None else None, device, max_ncomp))
        self.handle = h
        self._fin = weakref.finalize(self, L.hpsht_plan_destroy, h)
        if thetatot is not None or phi0 is not None:
            tt = None if thetatot is None else _lib.as_f64(thetatot)
            p0 = None if phi0 is None else _lib.as_f64(phi0)
            _lib.check(L.hpsht_plan_set_geometry(h, _lib.ptr(tt), _lib.ptr(p0)))
        s = [C.c_int64() for _ in range(5)]
        _lib.check(L.hpsht_plan_sizes(h, *[C.byref(x) for x in s]))
        self.nm_loc, self.nalm_loc, self.nring_loc, self.npix_loc, self.nring_tot = [x.value for x in s]

    def alm_info(self):
        mval = np.zeros(self.nm_loc, dtype=np.int64)
        mstart0 = np.zeros(self.nm_loc, dtype=np.int64)
        _lib.check(_lib.lib().hpsht_plan_alm_info(self.handle, _lib.ptr(mval), _lib.ptr(mstart0)))
        return mval, mstart0

    def map_info(self):
        n, nt = self.nring_loc, self.nring_tot
        rings, rstart0, nphi = (np.zeros(n, dtype=np.int64) for _ in range(3))
        theta, phi0 = np.zeros(n), np.zeros(n)
        thetatot = np.zeros(nt)
        _lib.check(_lib.lib().hpsht_plan_map_info(self.handle, _lib.ptr(rings), _lib.ptr(rstart0),
                                                  _lib.ptr(nphi), _lib.ptr(theta), _lib.ptr(phi0),
                                                  _lib.ptr(thetatot)))
        return dict(rings=rings, rstart0=rstart0, nphi=nphi, theta=theta, phi0=phi0, thetatot=thetatot)

    def alm2map(self, alm, mp, ncomp: int, nthreads: int = 0) -> None:
        _lib.check(_lib.lib().hpsht_alm2map(self.handle, _lib.ptr(alm), _lib.ptr(mp), ncomp, nthreads))

    def adjoint_alm2map(self, mp, alm, ncomp: int, nthreads: int = 0) -> None:
        _lib.check(_lib.lib().hpsht_adjoint_alm2map(self.handle, _lib.ptr(mp), _lib.ptr(alm), ncomp, nthreads))

    def alm2map_async(self, d_alm, d_map, ncomp: int, cuda_stream: int = 0) -> None:
        """Stream-ordered alm2map! on device-resident shards: ordered after the work queued on ``cuda_stream``
        (a raw cudaStream_t, e.g. ``torch.cuda.current_stream().cuda_stream``); does not block the host."""
        _lib.check(_lib.lib().hpsht_alm2map_async(self.handle, _lib.ptr(d_alm), _lib.ptr(d_map), ncomp,
                                                  C.c_void_p(cuda_stream)))

    def adjoint_alm2map_async(self, d_map, d_alm, ncomp: int, cuda_stream: int = 0) -> None:
        _lib.check(_lib.lib().hpsht_adjoint_alm2map_async(self.handle, _lib.ptr(d_map), _lib.ptr(d_alm), ncomp,
                                                          C.c_void_p(cuda_stream)))

    def synchronize(self) -> None:
        _lib.check(_lib.lib().hpsht_plan_synchronize(self.handle))

    def set_dev_blocks(self, on: bool) -> None:
        """P > 1: ring-block pipelining of device-pointer calls (same value on every rank)."""
        _lib.check(_lib.lib().hpsht_plan_set_dev_blocks(self.handle, 1 if on else 0))

    # ---- shards on the device: numpy arrays or torch CUDA tensors alike ----
    def scatter_alm(self, global_alm, alm, ncomp: int, root: int = 0) -> None:
        """MPI.Scatter!(alm, d_alm; root) (src/alm.jl:219-307): ``global_alm`` [ncomp][nalm_tot] complex is read
        on ``root`` only (None elsewhere); ``alm`` receives this rank's shard."""
        _lib.check(_lib.lib().hpsht_plan_scatter_alm(self.handle, _lib.ptr(global_alm) if global_alm is not None
                                                     else None, _lib.ptr(alm), ncomp, root))

    def gather_alm(self, alm, global_alm, ncomp: int, root: int = 0) -> None:
        """MPI.Gather! (root >= 0) / MPI.Allgather! (root = -1) (src/alm.jl:313-595)."""
        _lib.check(_lib.lib().hpsht_plan_gather_alm(self.handle, _lib.ptr(alm), _lib.ptr(global_alm)
                                                    if global_alm is not None else None, ncomp, root))

    def scatter_map(self, global_map, mp, ncomp: int, root: int = 0) -> None:
        """MPI.Scatter!(map, d_map; root) (src/map.jl:245-330)."""
        _lib.check(_lib.lib().hpsht_plan_scatter_map(self.handle, _lib.ptr(global_map) if global_map is not None
                                                     else None, _lib.ptr(mp), ncomp, root))

    def gather_map(self, mp, global_map, ncomp: int, root: int = 0) -> None:
        """MPI.Gather! / MPI.Allgather! (src/map.jl:337-559)."""
        _lib.check(_lib.lib().hpsht_plan_gather_map(self.handle, _lib.ptr(mp), _lib.ptr(global_map)
                                                    if global_map is not None else None, ncomp, root))

    def dot(self, alm1, alm2) -> float:
        """dot(a, b) of one component each (src/alm.jl:621-660), summed over the ranks."""
        r = C.c_double()
        _lib.check(_lib.lib().hpsht_plan_dot(self.handle, _lib.ptr(alm1), _lib.ptr(alm2), C.byref(r)))
        return r.value

    def almxfl(self, alm, fl, ncomp: int) -> None:
        """almxfl!(d_alm, fl) (src/alm.jl:693-713), in place.  Like the reference, column c of ``alm`` ([ncomp][nalm_loc])
        is scaled by column c of ``fl`` for the first min(ncomp, ncol(fl)) columns only: a 1-D ``fl`` touches the first
        component alone, ``fl`` of shape [ncol, n] scales ncol components, each by its own ``fl``, zero-extended to lmax + 1."""
        fl = np.atleast_2d(np.ascontiguousarray(fl, dtype=np.float64))
        base = _lib.ptr(alm)
        for c in range(min(ncomp, fl.shape[0])):
            flc = np.ascontiguousarray(fl[c])
            _lib.check(_lib.lib().hpsht_plan_almxfl(self.handle, base + c * self.nalm_loc * 16, _lib.ptr(flc), len(flc), 1))

    def alm2cl(self, alm1, alm2=None) -> np.ndarray:
        """alm2cl(a[, b]) of one component each (src/cl.jl:5-38), summed over the ranks."""
        cl = np.zeros(self.lmax + 1)
        _lib.check(_lib.lib().hpsht_plan_alm2cl(self.handle, _lib.ptr(alm1), _lib.ptr(alm1 if alm2 is None else alm2),
                                                _lib.ptr(cl)))
        return cl

    def timings(self) -> dict:
        ms = (C.c_double * _lib.HPSHT_NTIMES)()
        _lib.check(_lib.lib().hpsht_plan_timings(self.handle, ms))
        keys = ["total", "legendre_kernel", "legendre", "exchange", "ringfft", "h2d", "d2h", "_"]
        return dict(zip(keys, list(ms)))

    def launches(self) -> int:
        n = C.c_int64()
        _lib.check(_lib.lib().hpsht_plan_launches(self.handle, C.byref(n)))
        return n.value

    def legendre_steps(self, spin: int):
        c, d = C.c_double(), C.c_double()
        _lib.check(_lib.lib().hpsht_plan_legendre_steps(self.handle, spin, C.byref(c), C.byref(d)))
        return c.value, d.value

    def destroy(self) -> None:
        self._fin()


_PLANS: dict = {}


def _geom_key(thetatot, phi0):
    """The geometry a plan is built with is part of its identity: a DMap with another thetatot / phi0 must not
    pick up a plan made for the standard HEALPix rings."""
    if thetatot is None and phi0 is None:
        return None
    h = hashlib.blake2b(digest_size=16)
    for a in (thetatot, phi0):
        h.update(b"-" if a is None else np.ascontiguousarray(a, dtype=np.float64).tobytes())
    return h.hexdigest()


def get_plan(nside: int, lmax: int, comm: Comm, thetatot=None, phi0=None) -> Plan:
    key = (nside, lmax, comm, _geom_key(thetatot, phi0))
    p = _PLANS.get(key)
    if p is None:
        p = Plan(nside, lmax, comm, 2, -1, thetatot, phi0)
        _PLANS[key] = p
    return p


def clear_plans() -> None:
    for p in _PLANS.values():
        p.destroy()
    _PLANS.clear()


# --------------------------------------------------------------------------------------------
# the public transforms
# --------------------------------------------------------------------------------------------
def _require_colmajor(a, what: str) -> None:
    """Host arrays must have Julia's memory layout (column-major) — a C-ordered (n, ncomp) numpy array
    would silently interleave the components."""
    if isinstance(a, np.ndarray) and a.ndim >= 2 and a.shape[-1] > 1 and not a.flags.f_contiguous:
        raise _lib.DomainError(_lib.HPSHT_ERR_ARG, f"{what} must be column-major (order='F')")


def _checks(d_alm: DAlm, d_map: DMap):
    _require_colmajor(d_alm.alm, "DAlm.alm")
    _require_colmajor(d_map.pixels, "DMap.pixels")
    if d_alm.ncomp != d_map.ncomp:
        raise _lib.DomainError(_lib.HPSHT_ERR_ARG, "Number of components must match.")
    if d_alm.info.comm != d_map.info.comm:
        raise _lib.DomainError(_lib.HPSHT_ERR_ARG, "Communicators must match")
    return d_alm.info.comm


def _check_plan_matches(plan: Plan, d_alm: DAlm, d_map: DMap) -> None:
    """The C ABI takes raw pointers: everything it will assume about the two buffers is checked here
    (descriptor arrays, element type, extent), so a mismatch raises instead of reading or writing out of bounds."""
    for name, a, dt, n in (("DAlm.alm", d_alm.alm, np.complex128, plan.nalm_loc), ("DMap.pixels", d_map.pixels, np.float64, plan.npix_loc)):
        if isinstance(a, np.ndarray) and (a.dtype != dt or a.shape[0] != n):
            raise _lib.DomainError(_lib.HPSHT_ERR_ARG, f"{name} must be {np.dtype(dt).name} with {n} rows, got "
                                                       f"{a.dtype.name} with {a.shape[0]}")
    if d_alm.info.mmax != d_alm.info.lmax:
        raise _lib.DomainError(_lib.HPSHT_ERR_ARG, "mmax must equal lmax (src/alm.jl:171)")
    mval, mstart0 = plan.alm_info()
    if not (np.array_equal(mval, d_alm.info.mval) and np.array_equal(mstart0 + 1, d_alm.info.mstart)):
        raise _lib.DomainError(_lib.HPSHT_ERR_ARG, "DAlm.info does not describe an RR shard of this communicator")
    mi = plan.map_info()
    if not (np.array_equal(mi["rings"], d_map.info.rings) and np.array_equal(mi["rstart0"] + 1, d_map.info.rstart)
            and np.array_equal(mi["nphi"], d_map.info.nphi)):
        raise _lib.DomainError(_lib.HPSHT_ERR_ARG, "DMap.info does not describe an RR shard of this communicator")


def alm2map_(d_alm: DAlm, d_map: DMap, aux_in_leg=None, aux_out_leg=None, nthreads: int = 0) -> None:
    """alm2map!(d_alm, d_map[, aux_in_leg, aux_out_leg]; nthreads) (src/sht.jl:75-100)."""
    comm = _checks(d_alm, d_map)
    ncomp = d_alm.ncomp
    if (aux_in_leg is None) != (aux_out_leg is None):
        raise _lib.DomainError(_lib.HPSHT_ERR_ARG, "pass both auxiliary leg arrays or neither")
    if aux_in_leg is None:
        plan = get_plan(d_map.info.nside, d_alm.info.lmax, comm, d_map.info.thetatot, d_map.info.phi0)
        _check_plan_matches(plan, d_alm, d_map)
        plan.alm2map(d_alm.alm, d_map.pixels, ncomp, nthreads)
        return
    L = _lib.lib()
    ai, mi = d_alm.info, d_map.info
    spin = 0 if ncomp == 1 else 2
    mval = _lib.as_i64(ai.mval)
    mstart0 = _lib.as_i64(ai.mstart) - 1
    thetatot = _lib.as_f64(mi.thetatot)
    nm, nring_tot = len(mval), len(thetatot)
    assert aux_in_leg.shape == (nm, nring_tot, ncomp) and aux_in_leg.flags.f_contiguous
    _lib.check(L.hpsht_alm2leg(_lib.ptr(d_alm.alm), d_alm.alm.shape[0], _lib.ptr(aux_in_leg), spin, ai.lmax,
                               nm, _lib.ptr(mval), _lib.ptr(mstart0), 1, nring_tot, _lib.ptr(thetatot), nthreads))
    if comm.size() == 1:
        aux_out_leg = aux_in_leg  # src/sht.jl:87-88: the caller's aux_out_leg stays untouched
    else:
        communicate_alm2map_(aux_in_leg, aux_out_leg, comm, d_alm.strategy)
    nphi = _lib.as_i64(mi.nphi)
    phi0 = _lib.as_f64(mi.phi0)
    rstart0 = _lib.as_i64(mi.rstart) - 1
    assert aux_out_leg.shape[1] == len(nphi) and aux_out_leg.flags.f_contiguous
    _lib.check(L.hpsht_leg2map(_lib.ptr(aux_out_leg), _lib.ptr(d_map.pixels), d_map.pixels.shape[0], ncomp,
                               aux_out_leg.shape[0], len(nphi), _lib.ptr(nphi), _lib.ptr(phi0),
                               _lib.ptr(rstart0), 1, nthreads))


def adjoint_alm2map_(d_map: DMap, d_alm: DAlm, aux_in_leg=None, aux_out_leg=None, nthreads: int = 0) -> None:
    """adjoint_alm2map!(d_map, d_alm[, aux_in_leg, aux_out_leg]; nthreads) (src/sht.jl:165-185):
    Y^T without quadrature weights.  Here aux_in_leg is the MAP-side (tot_nm, loc_nr, ncomp) array
    and aux_out_leg the ALM-side (loc_nm, tot_nr, ncomp) one (src/sht.jl:182-183)."""
    comm = _checks(d_alm, d_map)
    ncomp = d_alm.ncomp
    if (aux_in_leg is None) != (aux_out_leg is None):
        raise _lib.DomainError(_lib.HPSHT_ERR_ARG, "pass both auxiliary leg arrays or neither")
    if aux_in_leg is None:
        plan = get_plan(d_map.info.nside, d_alm.info.lmax, comm, d_map.info.thetatot, d_map.info.phi0)
        _check_plan_matches(plan, d_alm, d_map)
        plan.adjoint_alm2map(d_map.pixels, d_alm.alm, ncomp, nthreads)
        return
    L = _lib.lib()
    ai, mi = d_alm.info, d_map.info
    spin = 0 if ncomp == 1 else 2
    nphi = _lib.as_i64(mi.nphi)
    phi0 = _lib.as_f64(mi.phi0)
    rstart0 = _lib.as_i64(mi.rstart) - 1
    assert aux_in_leg.shape[1:] == (len(nphi), ncomp) and aux_in_leg.flags.f_contiguous
    _lib.check(L.hpsht_map2leg(_lib.ptr(d_map.pixels), d_map.pixels.shape[0], _lib.ptr(aux_in_leg), ncomp,
                               aux_in_leg.shape[0], len(nphi), _lib.ptr(nphi), _lib.ptr(phi0),
                               _lib.ptr(rstart0), 1, nthreads))
    if comm.size() == 1:
        aux_out_leg = aux_in_leg
    else:
        communicate_map2alm_(aux_in_leg, aux_out_leg, comm, d_alm.strategy)
    mval = _lib.as_i64(ai.mval)
    mstart0 = _lib.as_i64(ai.mstart) - 1
    thetatot = _lib.as_f64(mi.thetatot)
    assert aux_out_leg.shape == (len(mval), len(thetatot), ncomp) and aux_out_leg.flags.f_contiguous
    _lib.check(L.hpsht_leg2alm(_lib.ptr(aux_out_leg), _lib.ptr(d_alm.alm), d_alm.alm.shape[0], spin, ai.lmax,
                               len(mval), _lib.ptr(mval), _lib.ptr(mstart0), 1, len(thetatot),
                               _lib.ptr(thetatot), nthreads))


def alloc_leg_buffers(d_alm: DAlm, d_map: DMap):
    """The two arrays the reference's allocating wrappers create (src/sht.jl:97-98):
    (alm-side (loc_nm, tot_nr, ncomp), map-side (tot_nm, loc_nr, ncomp))."""
    ncomp = d_alm.ncomp
    a = np.zeros((len(d_alm.info.mval), numOfRings(d_map.info.nside), ncomp), dtype=np.complex128, order="F")
    b = np.zeros((d_alm.info.mmax + 1, len(d_map.info.rings), ncomp), dtype=np.complex128, order="F")
    return a, b
