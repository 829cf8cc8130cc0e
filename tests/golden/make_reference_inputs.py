"""Inputs for pinning parity to the REAL reference (HealpixMPI.jl + Ducc0), which cannot run in the build image
(no Julia / MPI / Ducc0).  Writes, per case, the global alm (Healpix.jl Alm order: m-major, l = m..lmax) and the
global RING-ordered map as raw little-endian float64 files that tests/golden/make_reference_golden.jl reads on a
machine with Julia, plus a manifest.  The generators are bench.py's counter-based Philox streams, so any rank / any
language that follows the manifest sees identical data.

    python tests/golden/make_reference_inputs.py [outdir]      (default tests/golden/ref_inputs)

Cases: (64, 191) on 2 ranks and (1024, 3071) on 1 rank, spin 0 and spin 2."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from bench import synth_alm_shard, synth_map_shard  # noqa: E402

CASES = [dict(name="n64_l191", nside=64, lmax=191, ranks=2), dict(name="n1024_l3071", nside=1024, lmax=3071, ranks=1)]


def global_inputs(case, ncomp):
    nside, lmax = case["nside"], case["lmax"]
    seed = 7000 + nside + ncomp
    alm = synth_alm_shard(seed, ncomp, lmax, np.arange(lmax + 1))           # [ncomp, nalm], m-major = Alm order
    rings = np.arange(1, 4 * nside)
    nphi = np.where(rings < nside, 4 * rings, np.where(rings > 3 * nside, 4 * (4 * nside - rings), 4 * nside))
    mp = synth_map_shard(seed, ncomp, rings, nphi)                           # [ncomp, npix], RING order
    return alm, mp


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_inputs")
    os.makedirs(out, exist_ok=True)
    manifest = []
    for case in CASES:
        for ncomp in (1, 2):
            alm, mp = global_inputs(case, ncomp)
            tag = f"{case['name']}_s{0 if ncomp == 1 else 2}"
            alm.astype(np.complex128).tofile(os.path.join(out, tag + "_alm.c128"))
            mp.astype(np.float64).tofile(os.path.join(out, tag + "_map.f64"))
            manifest.append(dict(tag=tag, ncomp=ncomp, nalm=int(alm.shape[1]), npix=int(mp.shape[1]), **case))
    json.dump(manifest, open(os.path.join(out, "manifest.json"), "w"), indent=1)
    print("written", out, [m["tag"] for m in manifest])


if __name__ == "__main__":
    main()
