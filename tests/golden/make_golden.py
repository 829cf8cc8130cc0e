"""Generates tests/golden/golden_v1.npz from the extended-precision oracle (oracle/sht_oracle.c).
The reference holds no golden vectors and cannot run here (no Julia/MPI/Ducc0), so these fixtures
pin OUR oracle's outputs; the one deterministic-input case of the reference's own test-suite
(map[i] = i in RING order, nside 64, lmax 191, spin-0 adjoint: test/test_MPI_adj_a2m.jl:18) is
reproduced verbatim.   Run:  python tests/golden/make_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

rng = np.random.default_rng(20261019)
nside, lmax = 16, 47
nalm = (lmax + 1) * (lmax + 2) // 2
out = dict(nside_a=nside, lmax_a=lmax)
for spin, key in ((0, "s0"), (2, "s2")):
    ncomp = 1 if spin == 0 else 2
    alm = (rng.standard_normal((ncomp, nalm)) + 1j * rng.standard_normal((ncomp, nalm))) * np.sqrt(0.5)
    mp, _ = O.full_alm2map(alm, spin, nside, lmax)
    out[f"alm_{key}"] = alm
    out[f"map_{key}"] = mp.astype(np.float64)
nside, lmax = 64, 191
iota = np.arange(1, 12 * nside * nside + 1, dtype=np.float64)[None, :]
out["adj_alm_iota"] = O.full_adjoint(iota, 0, nside, lmax).astype(np.complex128)[0]
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_v1.npz"), **out)
print("written", {k: np.shape(v) for k, v in out.items()})
