"""Cycle accounting of k_leg2map by phase (needs the -DHPSHT_FFT_PROFILE build: tools/build_variant.sh):
   HPSHT_B200_LIB=healpixmpi.jl_b200/variants/libhpsht_b200_fftprof.so python tools/fft_profile.py"""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import healpixmpi_jl_b200 as H
from healpixmpi_jl_b200 import _lib
nside, lmax = 4096, 8191
plan = H.Plan(nside, lmax, H.COMM_SELF, 1)
a = torch.randn((1, plan.nalm_loc, 2), dtype=torch.float64, device="cuda")
m = torch.zeros((1, plan.npix_loc), dtype=torch.float64, device="cuda")
plan.alm2map(a, m, 1)
L = C.CDLL(_lib.LIB_PATH)
buf = (C.c_ulonglong * 64)()
L.hpsht_debug_fft_profile(buf, 1)
plan.alm2map(a, m, 1)
L.hpsht_debug_fft_profile(buf, 0)
names = ["root tables", "fold", "X->Zc->u|v", "FFT / Bluestein", "pixel store"]
for c, cn in enumerate(["power-of-two rings", "Bluestein M = 8192", "Bluestein M <= 4096", "power-of-two rings, half-ring CTAs"]):
    v = np.array(buf[c * 16:c * 16 + 5], dtype=float)
    if v.sum() == 0:
        continue
    print(cn, "total Mcycles %.1f:" % (v.sum() / 1e6), ", ".join(f"{n} {100 * x / v.sum():.0f}%" for n, x in zip(names, v)))
print("ring fft stage ms", plan.timings()["ringfft"])
L.hpsht_debug_fft_profile(buf, 1)
plan.adjoint_alm2map(m, a, 1)
L.hpsht_debug_fft_profile(buf, 0)
v = np.array(buf[3 * 16 + 8:3 * 16 + 12], dtype=float)
if v.sum() > 0:
    print("map2leg half-ring CTAs total Mcycles %.1f:" % (v.sum() / 1e6), ", ".join(f"{n} {100 * x / v.sum():.0f}%" for n, x in zip(["root tables", "pixel load", "FFT", "Z / X / un-alias / leg store"], v)))
print("ring fft stage ms (adjoint)", plan.timings()["ringfft"])
