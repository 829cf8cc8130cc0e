"""Ring-FFT stage time (ms) of device-resident transforms: python tools/fft_time.py [nside] [lmax]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import healpixmpi_jl_b200 as Hm  # noqa: E402

nside = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
lmax = int(sys.argv[2]) if len(sys.argv) > 2 else 2 * nside - 1
plan = Hm.Plan(nside, lmax, Hm.COMM_SELF, 1)
a = torch.randn((1, plan.nalm_loc, 2), dtype=torch.float64, device="cuda")
m = torch.randn((1, plan.npix_loc), dtype=torch.float64, device="cuda")
for name, fn, x, y in (("alm2map", plan.alm2map, a, m), ("adjoint", plan.adjoint_alm2map, m, a)):
    ts = []
    for _ in range(4):
        fn(x, y, 1)
        ts.append(plan.timings()["ringfft"])
    print(f"{name}: ring FFT stage {min(ts):.3f} ms", flush=True)
