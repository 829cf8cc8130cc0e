"""BASELINE config 5: spin-0 adjoint_alm2map! at nside=8192, lmax=16383 on one GPU (long-ring FFT path):
timing and the adjointness identity <f, Y a> = <Y^T f, a>_w (size-independent parity property)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import healpixmpi_jl_b200 as H

nside = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
lmax = int(sys.argv[2]) if len(sys.argv) > 2 else 2 * nside - 1
t0 = time.time()
plan = H.Plan(nside, lmax, H.COMM_SELF, 1)
g = torch.Generator(device="cuda").manual_seed(1)
f = torch.randn((1, plan.npix_loc), dtype=torch.float64, device="cuda", generator=g)
a = torch.randn((1, plan.nalm_loc, 2), dtype=torch.float64, device="cuda", generator=g)
a[0, :lmax + 1, 1] = 0.0  # real a_l0
out_a = torch.empty_like(a)
out_m = torch.empty_like(f)
for it in range(2):
    plan.adjoint_alm2map(f, out_a, 1)
    ta = plan.timings()
    plan.alm2map(a, out_m, 1)
    ts = plan.timings()
print(f"nside={nside} lmax={lmax}: setup+2 runs {time.time()-t0:.1f}s")
print("adjoint  ms:", {k: round(v, 2) for k, v in ta.items() if k != '_'})
print("alm2map  ms:", {k: round(v, 2) for k, v in ts.items() if k != '_'})
w = torch.ones(plan.nalm_loc, dtype=torch.float64, device="cuda") * 2.0
w[:lmax + 1] = 1.0
lhs = float((f * out_m).sum())
rhs = float((w * (out_a[0, :, 0] * a[0, :, 0] + out_a[0, :, 1] * a[0, :, 1])).sum())
nf, nm_ = float(f.norm()), float(out_m.norm())
print(f"adjointness: lhs {lhs:.12e} rhs {rhs:.12e} |diff|/(|f||Ya|) = {abs(lhs-rhs)/(nf*nm_):.2e}")
sc, sd = plan.legendre_steps(0)
print(f"Legendre kernels: adjoint {6*sc/ta['legendre_kernel']/1e9:.2f} TFLOP/s, synthesis {6*sc/ts['legendre_kernel']/1e9:.2f} TFLOP/s (culled steps {sc:.4e})")
print("max mem GB", torch.cuda.max_memory_allocated() / 1e9, "(torch only)")
