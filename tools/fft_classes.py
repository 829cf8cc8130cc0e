"""Ring-FFT stage time by ring class at nside 4096 (stage-level C ABI, device pointers):
   python tools/fft_classes.py [nside] [lmax]"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import healpixmpi_jl_b200 as H
from healpixmpi_jl_b200 import _lib

nside = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
lmax = int(sys.argv[2]) if len(sys.argv) > 2 else 2 * nside - 1
nm = lmax + 1
L = _lib.lib()

def geom(rings):
    out = [H.healpix_ring(nside, int(r)) if hasattr(H, "healpix_ring") else None for r in rings]
    return out

import ctypes as C
def ring(r):
    nphi = C.c_int64(); th = C.c_double(); p0 = C.c_double(); fp = C.c_int64()
    _lib.check(L.hpsht_healpix_ring(nside, int(r), C.byref(nphi), C.byref(th), C.byref(p0), C.byref(fp)))
    return nphi.value, p0.value

classes = {
    "belt (r = nside)": [r for r in range(nside, 3 * nside + 1)],
    "caps nside/2 < r < nside": [r for r in range(nside // 2 + 1, nside)] + [4 * nside - r for r in range(nside // 2 + 1, nside)],
    "caps nside/4 < r <= nside/2": [r for r in range(nside // 4 + 1, nside // 2 + 1)] + [4 * nside - r for r in range(nside // 4 + 1, nside // 2 + 1)],
    "caps r <= nside/4": [r for r in range(1, nside // 4 + 1)] + [4 * nside - r for r in range(1, nside // 4 + 1)],
}
tot = {"leg2map": 0.0, "map2leg": 0.0}
for name, rings in classes.items():
    g = [ring(r) for r in rings]
    nphi = np.array([x[0] for x in g], dtype=np.int64)
    phi0 = np.array([x[1] for x in g])
    rs = np.concatenate([[0], np.cumsum(nphi)[:-1]]).astype(np.int64)
    npix = int(nphi.sum())
    leg = torch.randn((1, len(rings), nm, 2), dtype=torch.float64, device="cuda")
    mp = torch.zeros((1, npix), dtype=torch.float64, device="cuda")
    res = {}
    for what in ("leg2map", "map2leg"):
        ts = []
        for _ in range(4):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            if what == "leg2map":
                _lib.check(L.hpsht_leg2map(_lib.ptr(leg), _lib.ptr(mp), npix, 1, nm, len(rings), _lib.ptr(nphi), _lib.ptr(phi0), _lib.ptr(rs), 1, 0))
            else:
                _lib.check(L.hpsht_map2leg(_lib.ptr(mp), npix, _lib.ptr(leg), 1, nm, len(rings), _lib.ptr(nphi), _lib.ptr(phi0), _lib.ptr(rs), 1, 0))
            torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
        res[what] = min(ts); tot[what] += min(ts)
    by = 16.0 * nm * len(rings) + 8.0 * npix
    print(f"{name:32s} {len(rings):6d} rings  leg2map {res['leg2map']:.3f} ms ({by/res['leg2map']/1e6:.0f} GB/s)  map2leg {res['map2leg']:.3f} ms ({by/res['map2leg']/1e6:.0f} GB/s)", flush=True)
print("sum", tot)
