"""Per-transform timing of host-pointer vs device-pointer calls on one GPU (who pays for PCIe).

usage: python tools/e2e_breakdown.py [nside] [lmax]      (HPSHT_HOST_BLOCKS selects the blocking)
"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import healpixmpi_jl_b200 as Hm  # noqa: E402
from healpixmpi_jl_b200 import _lib  # noqa: E402

nside = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
lmax = int(sys.argv[2]) if len(sys.argv) > 2 else 2 * nside - 1
plan = Hm.Plan(nside, lmax, Hm.COMM_SELF, 2)
rng = np.random.default_rng(0)


def pinned(shape, dtype):
    t = torch.empty(shape, dtype=dtype).pin_memory()
    return t


for ncomp in (1, 2):
    alm_h = pinned((ncomp, plan.nalm_loc, 2), torch.float64)
    alm_h.normal_()
    map_h = pinned((ncomp, plan.npix_loc), torch.float64)
    alm_d = alm_h.cuda()
    map_d = torch.empty((ncomp, plan.npix_loc), dtype=torch.float64, device="cuda")
    for name, fn, a, b in (("alm2map host", plan.alm2map, alm_h, map_h), ("alm2map dev ", plan.alm2map, alm_d, map_d),
                           ("adjoint host", plan.adjoint_alm2map, map_h, alm_h), ("adjoint dev ", plan.adjoint_alm2map, map_d, alm_d)):
        for it in range(3):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            fn(a, b, ncomp)
            torch.cuda.synchronize()
            wall = (time.perf_counter() - t0) * 1e3
        t = plan.timings()
        print(f"ncomp {ncomp} {name}: wall {wall:7.1f} ms  total {t['total']:7.1f}  legendre {t['legendre']:7.1f} "
              f"(kernel {t['legendre_kernel']:7.1f})  fft {t['ringfft']:6.1f}  h2d {t['h2d']:6.1f}  d2h {t['d2h']:6.1f}  "
              f"launches {plan.launches()}", flush=True)
