"""Packs what make_reference_golden.jl wrote into tests/golden/reference_<tag>.npz (inputs are regenerated from
the seeds by the test, only the reference's outputs are stored):  python tests/golden/pack_reference_golden.py [dir]"""
import json
import os
import sys

import numpy as np

here = os.path.dirname(os.path.abspath(__file__))
d = sys.argv[1] if len(sys.argv) > 1 else os.path.join(here, "ref_inputs")
for c in json.load(open(os.path.join(d, "manifest.json"))):
    fm, fa = os.path.join(d, c["tag"] + "_refmap.f64"), os.path.join(d, c["tag"] + "_refalm.c128")
    if not (os.path.exists(fm) and os.path.exists(fa)):
        continue
    mp = np.fromfile(fm, dtype=np.float64).reshape(c["ncomp"], c["npix"])
    alm = np.fromfile(fa, dtype=np.complex128).reshape(c["ncomp"], c["nalm"])
    np.savez_compressed(os.path.join(here, f"reference_{c['tag']}.npz"), refmap=mp, refalm=alm, **{k: c[k] for k in ("nside", "lmax", "ncomp", "ranks")})
    print("packed", c["tag"])
