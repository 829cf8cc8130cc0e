"""Device-resident timing of the shard operations at a given size on one GPU:
algorithmic bytes / wall time (the calls are synchronous; host-side descriptor building is included)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import healpixmpi_jl_b200 as Hm  # noqa: E402

nside = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
lmax = int(sys.argv[2]) if len(sys.argv) > 2 else 2 * nside - 1
plan = Hm.Plan(nside, lmax, Hm.COMM_SELF, 1)
nalm, npix = Hm.numberOfAlms(lmax), 12 * nside * nside
g_alm = torch.randn((1, nalm, 2), dtype=torch.float64, device="cuda")
g_map = torch.randn((1, npix), dtype=torch.float64, device="cuda")
a = torch.zeros((1, plan.nalm_loc, 2), dtype=torch.float64, device="cuda")
b = torch.randn((1, plan.nalm_loc, 2), dtype=torch.float64, device="cuda")
m = torch.zeros((1, plan.npix_loc), dtype=torch.float64, device="cuda")
fl = np.linspace(1.0, 2.0, lmax + 1)


def timeit(name, fn, nbytes, reps=5):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    t = min(ts)
    print(f"{name:14s} {1e3 * t:8.3f} ms  {nbytes / t / 1e9:8.1f} GB/s (algorithmic {nbytes / 1e9:.2f} GB)", flush=True)


timeit("scatter_alm", lambda: plan.scatter_alm(g_alm, a, 1, 0), 2 * 16 * nalm)
timeit("gather_alm", lambda: plan.gather_alm(a, g_alm, 1, 0), 2 * 16 * nalm)
timeit("scatter_map", lambda: plan.scatter_map(g_map, m, 1, 0), 2 * 8 * npix)
timeit("gather_map", lambda: plan.gather_map(m, g_map, 1, -1), 2 * 8 * npix)
timeit("dot", lambda: plan.dot(a[0], b[0]), 2 * 16 * nalm)
timeit("almxfl", lambda: plan.almxfl(a, fl, 1), 2 * 16 * nalm)
timeit("alm2cl", lambda: plan.alm2cl(a[0], b[0]), 2 * 16 * nalm)
