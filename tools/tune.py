"""Time kernel variants (env-selected) of the Legendre stage on device-resident data.
spec = <s|a><spin>:<cfg index>:<producer mode 1|2>   (s = synthesis, a = adjoint)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import healpixmpi_jl_b200 as H

nside, lmax = int(sys.argv[1]), int(sys.argv[2])
rng = np.random.default_rng(0)
for spec in sys.argv[3:]:
    parts = spec.split(":")
    kind = parts[0]
    adj = kind[0] == "a"
    spin = int(kind[1])
    if kind.endswith("w"):  # single-warp adjoint kernels: a0w:<R>
        os.environ["HPSHT_ADJ_WARP"] = "1"
        os.environ["HPSHT_A%dW_R" % spin] = parts[1]
        os.environ["HPSHT_ADJ_NW"] = parts[2] if len(parts) > 2 else "1"   # a2w:<R>:<warps per CTA>
    else:
        if adj:
            os.environ["HPSHT_ADJ_WARP"] = "0"
        os.environ["HPSHT_%s%d_CFG" % ("A" if adj else "S", spin)] = parts[1]
        os.environ["HPSHT_%s_PROD" % ("ADJ" if adj else "SYN")] = parts[2]
    ncomp = 1 if spin == 0 else 2
    plan = H.Plan(nside, lmax, H.COMM_SELF, ncomp)
    alm = torch.from_numpy((rng.standard_normal((ncomp, plan.nalm_loc, 2)))).cuda()
    mp = torch.from_numpy(rng.standard_normal((ncomp, plan.npix_loc))).cuda()
    ts = []
    for i in range(3):
        if adj:
            plan.adjoint_alm2map(mp, alm, ncomp)
        else:
            plan.alm2map(alm, mp, ncomp)
        ts.append(plan.timings()["legendre_kernel"])
    sc, _ = plan.legendre_steps(spin)
    fl = (6.0 if spin == 0 else 26.0) * sc
    print(f"{spec}: kernel ms {min(ts):.2f}  -> {fl/min(ts)/1e9:.2f} TFLOP/s (algorithmic)", flush=True)
    plan.destroy()
    del alm, mp
    torch.cuda.empty_cache()
