#!/usr/bin/env python
"""Print BASELINE.md section 5 (the result table) from the bench lines under profiles/.

    python tools/baseline_table.py [round-prefix, default r02]
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PROF = os.path.join(HERE, "..", "profiles")
R = sys.argv[1] if len(sys.argv) > 1 else "r02"

ROWS = [  # (label, file stem pattern, [(spin, dir key)], GPU counts)
    ("(4096, 8191)", "bench_n{n}", ["spin0_a2m", "spin0_adj", "spin2_a2m", "spin2_adj"], [1, 2, 4, 8]),
    ("(2048, 6143)", "bench_config3_n{n}", ["spin0_a2m", "spin2_a2m"], [1, 8]),
    ("(1024, 3071)", "bench_config2_n{n}", ["spin0_a2m", "spin0_adj"], [1]),
    ("(8192, 16383)", "bench_config5_n{n}", ["spin0_adj"], [1, 8]),
]
KERNEL = {"spin0_a2m": "k_alm2leg_s0", "spin0_adj": "k_leg2alm_s0", "spin2_a2m": "k_alm2leg_s2", "spin2_adj": "k_leg2alm_s2"}


def load(stem):
    path = os.path.join(PROF, f"{R}_{stem}.json")
    if not os.path.exists(path):
        return None
    for line in open(path):
        if line.startswith("{"):
            return json.loads(line)
    return None


def frac_of(d, kern):
    r = d.get("roofline") or {}
    if r.get("kernel", "").startswith(kern):
        return r.get("frac")
    return ((r.get("other_kernels") or {}).get(kern) or {}).get("frac")


def f(x, nd=1):
    return "—" if x is None else f"{x:.{nd}f}"


print("| config | transform | GPUs | total ms (timed step) | Legendre kernels ms | of FP64 peak | ring FFT ms | algorithmic GB/s "
      "(of HBM copy) | exchange ms | egress GB/s per GPU (of 900) | sampled rings vs oracle (rel. L2) | adjointness |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|")
for label, stem, keys, gpus in ROWS:
    for n in gpus:
        d = load(stem.format(n=n))
        if d is None:
            continue
        st, b2b = d["stages_ms"], d.get("stages_back_to_back_ms") or d["stages_ms"]
        fft = (d.get("roofline_fft") or {}).get("per_transform") or {}
        nvl = (d.get("roofline_nvlink") or {}).get("per_transform") or {}
        par = d.get("parity") or {}
        for k in keys:
            spin = k[:5]
            ft, nv = fft.get(k) or {}, nvl.get(k) or {}
            rings = ((par.get("rings_vs_oracle") or {}).get(spin) or {}).get("rel_l2") if k.endswith("a2m") else None
            adj = (par.get("adjointness") or {}).get(spin)
            fr = frac_of(d, KERNEL[k])
            print(f"| {label} | {k.replace('_', ' ')} | {n} | {f(st[k + '_total_ms'], 2)} | {f(b2b[k + '_legendre_kernel_ms'], 2)} | "
                  f"{f(fr, 2)} | {f(b2b[k + '_ringfft_ms'], 2)} | "
                  f"{f(ft.get('gbs'), 0)} ({f(ft.get('frac'), 2)}) | {f(b2b[k + '_exchange_ms'], 2) if n > 1 else '—'} | "
                  f"{(f(nv.get('gbs'), 0) + ' (' + f(nv.get('frac'), 2) + ')') if nv else '—'} | "
                  f"{'—' if rings is None else format(rings, '.1e')} | {'—' if adj is None else format(adj, '.0e')} |")
print()
print("| config | GPUs | device-resident step ms | host-pointer step ms (`e2e`) | FP64 OpenMP CPU port, same step (threads) |")
print("|---|---|---|---|---|")
for label, stem, keys, gpus in ROWS:
    for n in gpus:
        d = load(stem.format(n=n))
        if d is None:
            continue
        e = (d.get("e2e") or {}).get("value")
        c = d.get("cpu_baseline") or {}
        print(f"| {label} {' + '.join(k.replace('_', ' ') for k in keys)} | {n} | {f(d['ms_per_step'], 2)} | {f(e, 2)} | "
              f"{(f(c.get('value'), 0) + ' ms (' + str(c.get('cores')) + ')') if c else '—'} |")
