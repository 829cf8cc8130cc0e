"""healpixmpi.jl_b200 — B200-native drop-in for HealpixMPI.jl's distributed spherical-harmonic
transform hot path (alm2map! / adjoint_alm2map! on DAlm{RR} / DMap{RR}, spin 0 and spin 2).

The compute lives in libhpsht_b200.so (hand-written sm_100a CUDA behind the C ABI declared in
include/hpsht_b200.h); this package is the host-side mirror of the reference's Julia interface
(same names, argument meaning and error behaviour; Julia's ``f!`` is spelled ``f_``).  There is no
CPU fallback: without the built library / a CUDA device the transforms raise."""
from . import _lib
from ._lib import DomainError, HpshtError, build
from .alm import (AlmInfoMPI, DAlm, Gather_alm_, Scatter_alm_, dot, get_m_tasks_RR, get_mval_RR, get_nm_RR,
                  localdot, make_mstart_complex)
from .comm import COMM_SELF, Comm, TorchComm, comm_world
from .healpix import (Alm, HealpixMap, PolarizedHealpixMap, getEquatorIdx, getringinfo, numberOfAlms,
                      numOfRings, nside2npix, ring2theta)
from .map import (DMap, GeomInfoMPI, Gather_map_, Scatter_array, Scatter_map_, get_nrings_RR,
                  get_rindexes_RR, get_rindexes_tot_RR)
from .sht import (Plan, adjoint_alm2map_, alloc_leg_buffers, alm2map_, clear_plans, communicate_alm2map_,
                  communicate_map2alm_, get_plan)
from .strategy import RR, Strategy


def Scatter_(obj, *outs, **kw):
    """MPI.Scatter! overload set (src/alm.jl:219-307, src/map.jl:245-330): dispatches on the output
    type like the Julia methods do."""
    if not outs:
        raise TypeError("Scatter_ needs at least one distributed output")
    if isinstance(outs[0], DAlm):
        return Scatter_alm_(obj, *outs, **kw)
    return Scatter_map_(obj, *outs, **kw)


def Gather_(d_obj, out, comp: int = 1, **kw):
    """MPI.Gather! overload set (src/alm.jl:411-486, src/map.jl:429-471)."""
    if isinstance(d_obj, DAlm):
        return Gather_alm_(d_obj, out, comp, **kw)
    return Gather_map_(d_obj, out, comp, **kw)


def Allgather_(d_obj, out, comp: int = 1, **kw):
    """MPI.Allgather! overload set (src/alm.jl:550-595, src/map.jl:531-559)."""
    return Gather_(d_obj, out, comp, allgather=True, **kw)
