"""Run one transform a few times on device-resident random shards (for ncu captures and quick timings).
   python tools/run_one.py NSIDE LMAX NCOMP a2m|adj [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import healpixmpi_jl_b200 as H

nside, lmax, ncomp, d = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 3
plan = H.Plan(nside, lmax, H.COMM_SELF, ncomp)
g = torch.Generator(device="cuda").manual_seed(1)
a = torch.randn((ncomp, plan.nalm_loc, 2), dtype=torch.float64, device="cuda", generator=g)
f = torch.randn((ncomp, plan.npix_loc), dtype=torch.float64, device="cuda", generator=g)
for i in range(reps):
    if d == "a2m":
        plan.alm2map(a, f, ncomp)
    else:
        plan.adjoint_alm2map(f, a, ncomp)
    t = plan.timings()
    print(f"{d} ncomp={ncomp} nside={nside}: total {t['total']:.2f} ms, legendre kernels {t['legendre_kernel']:.2f} ms, "
          f"ring fft {t['ringfft']:.2f} ms, launches {plan.launches()}", flush=True)
plan.destroy()
