"""nside 8192 / lmax 16383 (long-ring FFT path): host-pointer calls (ring blocks, HPSHT_HOST_BLOCKS) against
device-pointer calls.  usage: [HPSHT_HOST_BLOCKS=n] python tools/check_nside8192.py [nside lmax]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import healpixmpi_jl_b200 as Hm  # noqa: E402

nside = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
lmax = int(sys.argv[2]) if len(sys.argv) > 2 else 2 * nside - 1
plan = Hm.Plan(nside, lmax, Hm.COMM_SELF, 1)
REPS = int(os.environ.get("CHECK_REPS", "2"))
g = torch.Generator(device="cuda").manual_seed(1)
if os.environ.get("CHECK_EARLY_ALLOC"):
    _early = [torch.empty((1, plan.npix_loc), dtype=torch.float64, device="cuda"),
              torch.empty((1, plan.nalm_loc, 2), dtype=torch.float64, device="cuda")]
    del _early  # returns to torch's cache: the later f / o_dev reuse these blocks
a = torch.randn((1, plan.nalm_loc, 2), dtype=torch.float64, device="cuda", generator=g)
m_dev = torch.zeros((1, plan.npix_loc), dtype=torch.float64, device="cuda")
for _ in range(REPS):
    plan.alm2map(a, m_dev, 1)
print("a2m dev ", plan.launches(), {k: round(v, 1) for k, v in plan.timings().items()})
PIN = os.environ.get("CHECK_PINNED", "0") == "1"
a_h = a.cpu().pin_memory().numpy() if PIN else a.cpu().numpy()
m_h = torch.zeros((1, plan.npix_loc), dtype=torch.float64).pin_memory().numpy() if PIN else np.zeros((1, plan.npix_loc))
for _ in range(REPS):
    plan.alm2map(a_h, m_h, 1)
print("a2m host", plan.launches(), {k: round(v, 1) for k, v in plan.timings().items()})
print("synthesis host == device:", np.array_equal(m_h, m_dev.cpu().numpy()))
f = torch.randn((1, plan.npix_loc), dtype=torch.float64, device="cuda", generator=g)
o_dev = torch.zeros_like(a)
for _ in range(REPS):
    plan.adjoint_alm2map(f, o_dev, 1)
print("adj dev ", plan.launches(), {k: round(v, 1) for k, v in plan.timings().items()})
o_h = torch.zeros((1, plan.nalm_loc, 2), dtype=torch.float64).pin_memory().numpy() if PIN else np.zeros((1, plan.nalm_loc, 2))
f_h = f.cpu().pin_memory().numpy() if PIN else f.cpu().numpy()
for _ in range(REPS):
    plan.adjoint_alm2map(f_h, o_h, 1)
print("adj host", plan.launches(), {k: round(v, 1) for k, v in plan.timings().items()})
d = torch.from_numpy(o_h).cuda() - o_dev
print("adjoint host vs device rel", float(torch.linalg.norm(d) / torch.linalg.norm(o_dev)))
mval, mstart0 = plan.alm_info()
dm = d[0].norm(dim=1).cpu().numpy()
bad = np.nonzero(dm > 1e-9 * float(o_dev.abs().max()))[0]
if len(bad):
    idx_m = np.searchsorted(mstart0 + mval, bad, side="right") - 1
    print("bad elements", len(bad), "m range", mval[idx_m].min(), mval[idx_m].max(), "first l", (bad - mstart0[idx_m])[:5])
