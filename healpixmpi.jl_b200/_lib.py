"""ctypes binding of libhpsht_b200.so (include/hpsht_b200.h).  This is the Python analogue of the
Julia ``ccall`` shim in healpixmpi.jl_b200/julia/HealpixMPIB200.jl; every symbol the header
declares is bound here with explicit argtypes."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# HPSHT_B200_LIB: another build of the same library (kernel experiments, tools/build_variant.sh)
LIB_PATH = os.environ.get("HPSHT_B200_LIB") or os.path.join(_HERE, "libhpsht_b200.so")

i64 = C.c_int64
i64p = C.POINTER(C.c_int64)
f64p = C.POINTER(C.c_double)
vp = C.c_void_p

HPSHT_OK = 0
HPSHT_ERR_ARG = 1
HPSHT_ERR_CUDA = 2
HPSHT_ERR_NCCL = 3
HPSHT_ERR_UNSUPPORTED = 4
HPSHT_ERR_EMPTY_RANK = 5
HPSHT_ERR_INTERNAL = 6
HPSHT_NTIMES = 8
HPSHT_NCCL_ID_BYTES = 128

# name -> (restype, argtypes); kept in one table so tests can check that the library exports
# exactly what include/hpsht_b200.h declares.
SIGNATURES = {
    "hpsht_version": (C.c_int, []),
    "hpsht_last_error": (C.c_char_p, []),
    "hpsht_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "hpsht_rr_nm": (C.c_int, [i64, C.c_int, C.c_int, i64p]),
    "hpsht_rr_mval": (C.c_int, [i64, C.c_int, C.c_int, vp]),
    "hpsht_rr_m_tasks": (C.c_int, [i64, C.c_int, vp]),
    "hpsht_rr_mstart": (C.c_int, [i64, i64, i64, vp, vp]),
    "hpsht_rr_nrings": (C.c_int, [i64, C.c_int, C.c_int, i64p]),
    "hpsht_rr_rindexes": (C.c_int, [i64, C.c_int, C.c_int, vp]),
    "hpsht_rr_rindexes_tot": (C.c_int, [i64, C.c_int, vp]),
    "hpsht_rr_a2a_counts": (C.c_int, [i64, i64, C.c_int, C.c_int, vp, vp]),
    "hpsht_healpix_ring": (C.c_int, [i64, i64, i64p, f64p, f64p, i64p]),
    "hpsht_alm2leg": (C.c_int, [vp, i64, vp, C.c_int, i64, i64, vp, vp, i64, i64, vp, C.c_int]),
    "hpsht_leg2alm": (C.c_int, [vp, vp, i64, C.c_int, i64, i64, vp, vp, i64, i64, vp, C.c_int]),
    "hpsht_leg2map": (C.c_int, [vp, vp, i64, C.c_int, i64, i64, vp, vp, vp, i64, C.c_int]),
    "hpsht_map2leg": (C.c_int, [vp, i64, vp, C.c_int, i64, i64, vp, vp, vp, i64, C.c_int]),
    "hpsht_nccl_unique_id": (C.c_int, [vp]),
    "hpsht_plan_create": (C.c_int, [C.POINTER(vp), i64, i64, C.c_int, C.c_int, vp, C.c_int, C.c_int]),
    "hpsht_plan_set_geometry": (C.c_int, [vp, vp, vp]),
    "hpsht_plan_destroy": (C.c_int, [vp]),
    "hpsht_plan_sizes": (C.c_int, [vp, i64p, i64p, i64p, i64p, i64p]),
    "hpsht_plan_alm_info": (C.c_int, [vp, vp, vp]),
    "hpsht_plan_map_info": (C.c_int, [vp, vp, vp, vp, vp, vp, vp]),
    "hpsht_alm2map": (C.c_int, [vp, vp, vp, C.c_int, C.c_int]),
    "hpsht_adjoint_alm2map": (C.c_int, [vp, vp, vp, C.c_int, C.c_int]),
    "hpsht_alm2map_async": (C.c_int, [vp, vp, vp, C.c_int, vp]),
    "hpsht_adjoint_alm2map_async": (C.c_int, [vp, vp, vp, C.c_int, vp]),
    "hpsht_plan_synchronize": (C.c_int, [vp]),
    "hpsht_plan_set_dev_blocks": (C.c_int, [vp, C.c_int]),
    "hpsht_ring_block_ranges": (C.c_int, [i64, i64, C.c_int, C.POINTER(C.c_int), vp, vp, C.c_int]),
    "hpsht_plan_scatter_alm": (C.c_int, [vp, vp, vp, C.c_int, C.c_int]),
    "hpsht_plan_gather_alm": (C.c_int, [vp, vp, vp, C.c_int, C.c_int]),
    "hpsht_plan_scatter_map": (C.c_int, [vp, vp, vp, C.c_int, C.c_int]),
    "hpsht_plan_gather_map": (C.c_int, [vp, vp, vp, C.c_int, C.c_int]),
    "hpsht_plan_dot": (C.c_int, [vp, vp, vp, f64p]),
    "hpsht_plan_almxfl": (C.c_int, [vp, vp, vp, i64, C.c_int]),
    "hpsht_plan_alm2cl": (C.c_int, [vp, vp, vp, vp]),
    "hpsht_plan_timings": (C.c_int, [vp, f64p]),
    "hpsht_plan_launches": (C.c_int, [vp, i64p]),
    "hpsht_plan_legendre_steps": (C.c_int, [vp, C.c_int, f64p, f64p]),
    "hpsht_device_malloc": (C.c_int, [C.POINTER(vp), C.c_size_t]),
    "hpsht_device_free": (C.c_int, [vp]),
    "hpsht_host_malloc": (C.c_int, [C.POINTER(vp), C.c_size_t]),
    "hpsht_host_free": (C.c_int, [vp]),
    "hpsht_memcpy_h2d": (C.c_int, [vp, vp, C.c_size_t]),
    "hpsht_memcpy_d2h": (C.c_int, [vp, vp, C.c_size_t]),
    "hpsht_device_synchronize": (C.c_int, []),
    "hpsht_measure_fp64_peak": (C.c_int, [f64p, f64p]),
}


class HpshtError(RuntimeError):
    """Non-zero status from the C library (the Julia shim raises DomainError / ErrorException)."""

    def __init__(self, code: int, msg: str):
        super().__init__(f"hpsht error {code}: {msg}")
        self.code = code


class DomainError(HpshtError, ValueError):
    """Argument errors — what the reference throws as DomainError (src/sht.jl:76-77 etc.)."""


_LIB = None


def build(force: bool = False) -> None:
    """Compile the CUDA library in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    args = ["make", "-s", "-C", os.path.join(_HERE, "csrc"), "-j4"]
    if force:
        args.append("-B")
    subprocess.check_call(args)


def lib() -> C.CDLL:
    """The loaded library.  Fails loudly when the extension is missing: there is no fallback."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with __graft_entry__.build() "
                "(make -C healpixmpi.jl_b200/csrc); there is no CPU fallback.")
        L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        for name, (res, args) in SIGNATURES.items():
            f = getattr(L, name)
            f.restype = res
            f.argtypes = args
        _LIB = L
    return _LIB


def check(status: int) -> None:
    if status != HPSHT_OK:
        msg = lib().hpsht_last_error().decode("utf-8", "replace")
        if status in (HPSHT_ERR_ARG, HPSHT_ERR_EMPTY_RANK):
            raise DomainError(status, msg)
        raise HpshtError(status, msg)


def ptr(a):
    """Raw address of a numpy array or of anything with .data_ptr() (torch tensors, device
    arrays).  None -> NULL."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    if hasattr(a, "data_ptr"):
        return a.data_ptr()
    if isinstance(a, int):
        return a
    raise TypeError(f"cannot take the address of {type(a)}")


def as_i64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.int64)


def as_f64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float64)
