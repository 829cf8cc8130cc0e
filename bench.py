#!/usr/bin/env python
"""bench.py — the reference's headline benchmark on synthetic data.

    python bench.py --gpus N --steps K --warmup W        (N>1: launched under torchrun, one rank/GPU)

One STEP = the headline workload of BASELINE.json (configs[3]): spin-0 alm2map! + spin-0
adjoint_alm2map! + spin-2 alm2map! + spin-2 adjoint_alm2map! at nside=4096, lmax=8191 on
DAlm{RR}/DMap{RR} shards spread over the N ranks (strong scaling: the problem is fixed).
It fits one B200 (about 12 GB), so N=1 runs the same workload.

  value       ms per step, shards resident in HBM when the timed region starts (CUDA events on the
              plan's stream, max over ranks)
  e2e         ms per step through the public API (alm2map_/adjoint_alm2map_ on DAlm/DMap) with HOST
              (pinned) buffers: H2D of the inputs and D2H of the results inside the timed region
  roofline    the Legendre kernel with the largest share of the step against the FP64 FMA pipe
  cpu_baseline  the FP64 OpenMP CPU port (oracle/cpu_port.*; kind "port": the real reference needs
              Julia+MPI+Ducc0, absent here) timed on a bounded sample and extrapolated to the step

--impl reference times that CPU port alone (rank 0 only) and prints the same JSON shape.
Smaller problems for smoke runs: --nside/--lmax.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "alm2map+adjoint ms and FP64 %peak at nside=4096 lmax=8191, 1/2/4/8 B200"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ------------------------------------------------------------------------------------------------
# synthetic data: counter-based, a function of (seed, component, global m / global ring) only, so
# every rank builds its own shard and 1/2/4/8-GPU runs see identical data
# ------------------------------------------------------------------------------------------------
def synth_alm_shard(seed, ncomp, lmax, mval):
    cols = []
    for c in range(ncomp):
        rows = []
        for m in mval:
            g = np.random.Generator(np.random.Philox(key=[seed, (c << 32) + int(m)]))
            n = lmax + 1 - int(m)
            rows.append((g.standard_normal(n) + 1j * g.standard_normal(n)) * np.sqrt(0.5))
        cols.append(np.concatenate(rows))
    return np.stack(cols)  # [ncomp, nalm_loc]  == Julia (nalm_loc, ncomp) column-major


def synth_map_shard(seed, ncomp, rings, nphi):
    cols = []
    for c in range(ncomp):
        rows = []
        for r, n in zip(rings, nphi):
            g = np.random.Generator(np.random.Philox(key=[seed + 7919, (c << 32) + int(r)]))
            rows.append(g.standard_normal(int(n)))
        cols.append(np.concatenate(rows))
    return np.stack(cols)


# ------------------------------------------------------------------------------------------------
# clocks sampler (B200_PROFILING.md recipe)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                 "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception as e:  # pragma: no cover
            log("clock sampler unavailable:", e)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"),
                               f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU baseline (port): bounded sample, extrapolated through the work model
# ------------------------------------------------------------------------------------------------
def cpu_port_step_ms(nside, lmax, budget_s=20.0):
    """Time the FP64 OpenMP port (Legendre: oracle/cpu_port.cpp; rings: scipy/pocketfft) on a sample of
    the workload and extrapolate to one full step (spin-0 + spin-2, both directions)."""
    from oracle import cpu_port as P
    from oracle import oracle as O

    cores = os.cpu_count() or 1
    nr = 4 * nside - 1
    rings = np.arange(1, nr + 1)
    nphi, theta, phi0, fp = O.healpix_rings(nside, rings)
    all_m = np.arange(lmax + 1)
    mstart_all = O.rr_mstart1(lmax, 1, all_m) - 1
    nalm = (lmax + 1) * (lmax + 2) // 2
    rng = np.random.default_rng(0)

    def leg_time(spin, stride):
        ncomp = 1 if spin == 0 else 2
        mval = all_m[::stride]
        alm = rng.standard_normal((ncomp, nalm)) + 1j * rng.standard_normal((ncomp, nalm))
        t0 = time.perf_counter()
        leg = P.alm2leg(alm, spin, lmax, mval, mstart_all[mval], theta)
        t1 = time.perf_counter()
        P.leg2alm(leg, spin, lmax, mval, mstart_all[mval], theta, nalm)
        t2 = time.perf_counter()
        s_c, _ = P.legendre_steps(spin, lmax, mval, theta)
        return (t1 - t0), (t2 - t1), s_c

    # calibrate on a sparse m-subset, then size the real sample for ~budget_s
    stride = max(1, (lmax + 1) // 16)
    a, b, s = leg_time(0, stride)
    rate = s / max(a, 1e-9)
    full0, _ = P.legendre_steps(0, lmax, all_m, theta)
    full2, _ = P.legendre_steps(2, lmax, all_m, theta)
    est_full = (full0 * 2 + full2 * 2 * 4.3) / rate
    stride = int(max(1, np.ceil(est_full / (0.6 * budget_s))))
    out = {}
    tot_ms = 0.0
    sample_steps = 0.0
    for spin, full in ((0, full0), (2, full2)):
        a, b, s = leg_time(spin, stride)
        tot_ms += (a + b) * (full / s) * 1e3
        sample_steps += 2 * s
        out[f"legendre_spin{spin}_steps_per_s"] = s / a
    # ring stage on a sample of rings, extrapolated by pixel count (spin 0: 1 comp, spin 2: 2 comps)
    nsel = max(8, min(nr, int(nr * 0.02)))
    sel = np.unique(np.linspace(0, nr - 1, nsel).astype(int))
    rs0 = np.concatenate([[0], np.cumsum(nphi[sel])[:-1]])
    leg = rng.standard_normal((1, len(sel), lmax + 1)) + 1j * rng.standard_normal((1, len(sel), lmax + 1))
    t0 = time.perf_counter()
    mp = P.leg2map(leg, nphi[sel], phi0[sel], rs0, int(nphi[sel].sum()))
    P.map2leg(mp, lmax + 1, nphi[sel], phi0[sel], rs0)
    t1 = time.perf_counter()
    ring_ms = (t1 - t0) * (nr / len(sel)) * 3 * 1e3  # 3 components per step, both directions timed
    tot_ms += ring_ms
    sample = (f"Legendre: every {stride}-th m of 0..{lmax} x all {nr} rings, spin 0 and spin 2, both "
              f"directions ({sample_steps:.3g} of {2 * (full0 + full2):.3g} steps), OpenMP {cores} threads; "
              f"ring FFT: {len(sel)} of {nr} rings via scipy.fft (single thread); extrapolated to the full step")
    return tot_ms, cores, sample, out


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--nside", type=int, default=4096)
    ap.add_argument("--lmax", type=int, default=8191)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cpu-budget", type=float, default=20.0)
    args = ap.parse_args()

    # stdout carries exactly one JSON line: route everything else (NCCL banners, library chatter) to
    # stderr until the final print
    sys.stdout.flush()
    _saved_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        sys.stdout.flush()
        os.dup2(_saved_stdout, 1)
        print(json.dumps(obj), flush=True)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    nside, lmax = args.nside, args.lmax
    workload = (f"spin-0 + spin-2 alm2map!/adjoint_alm2map! nside={nside} lmax={lmax}, DAlm{{RR}}/DMap{{RR}} "
                f"over {world} rank(s)")
    config = {"workload": workload, "nside": nside, "lmax": lmax, "ranks": world,
              "inputs": "counter-based N(0,1) alm / maps (Philox), larger than L2 (no flush needed)"}

    if args.impl == "reference":
        if rank != 0:
            return
        tot_ms = []
        cores, sample = None, None
        for i in range(args.warmup + args.steps):
            ms, cores, sample, _ = cpu_port_step_ms(nside, lmax, args.cpu_budget)
            if i >= args.warmup:
                tot_ms.append(ms)
        v = float(np.mean(tot_ms))
        emit({
            "impl": "reference", "metric": METRIC, "value": v, "unit": "ms", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": v, "higher_is_better": False,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
            "cpu_baseline": {"value": v, "unit": "ms", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "FP64 OpenMP CPU port of the same staged algorithm (not ducc; the reference itself "
                    "needs Julia+MPI+Ducc0 which are absent), bounded sample extrapolated to the full step"})
        return

    import torch
    import torch.distributed as dist

    import healpixmpi_jl_b200 as H

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        comm = H.TorchComm()
    else:
        comm = H.COMM_SELF

    t_setup = time.perf_counter()
    plan = H.Plan(nside, lmax, comm, 2, local_rank)
    mval, _ = plan.alm_info()
    minfo = plan.map_info()
    log(f"[rank {rank}] plan: nm_loc={plan.nm_loc} nalm_loc={plan.nalm_loc} nring_loc={plan.nring_loc} "
        f"npix_loc={plan.npix_loc}")

    # ---- shards: host (pinned) and device copies ----
    def pinned(a):
        t = torch.from_numpy(np.ascontiguousarray(a))
        return t.pin_memory()

    h_alm = {1: pinned(synth_alm_shard(4096 * 4 + 0, 1, lmax, mval).view(np.float64)),
             2: pinned(synth_alm_shard(4096 * 4 + 1, 2, lmax, mval).view(np.float64))}
    h_map = {1: pinned(synth_map_shard(4096 * 4 + 2, 1, minfo["rings"], minfo["nphi"])),
             2: pinned(synth_map_shard(4096 * 4 + 3, 2, minfo["rings"], minfo["nphi"]))}
    d_alm = {k: v.cuda() for k, v in h_alm.items()}
    d_map = {k: v.cuda() for k, v in h_map.items()}
    d_alm_out = {k: torch.empty_like(v) for k, v in d_alm.items()}
    d_map_out = {k: torch.empty_like(v) for k, v in d_map.items()}
    h_alm_out = {k: torch.empty_like(v).pin_memory() for k, v in h_alm.items()}
    h_map_out = {k: torch.empty_like(v).pin_memory() for k, v in h_map.items()}
    log(f"[rank {rank}] setup {time.perf_counter() - t_setup:.1f}s")

    stage_acc = {}

    def one_step(resident: bool):
        """4 transforms; returns (device ms summed from the plan's CUDA events, kernel launches)."""
        ms, launches = 0.0, 0
        for ncomp in (1, 2):
            a_in = d_alm[ncomp] if resident else h_alm[ncomp]
            m_out = d_map_out[ncomp] if resident else h_map_out[ncomp]
            plan.alm2map(a_in, m_out, ncomp)
            t = plan.timings()
            ms += t["total"]
            launches += plan.launches()
            for k, v in t.items():
                stage_acc[(ncomp, "a2m", k)] = stage_acc.get((ncomp, "a2m", k), 0.0) + v
            m_in = d_map[ncomp] if resident else h_map[ncomp]
            a_out = d_alm_out[ncomp] if resident else h_alm_out[ncomp]
            plan.adjoint_alm2map(m_in, a_out, ncomp)
            t = plan.timings()
            ms += t["total"]
            launches += plan.launches()
            for k, v in t.items():
                stage_acc[(ncomp, "adj", k)] = stage_acc.get((ncomp, "adj", k), 0.0) + v
        return ms, launches

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(resident: bool, steps: int, warmup: int):
        for _ in range(warmup):
            one_step(resident)
        stage_acc.clear()
        barrier()
        t0 = time.perf_counter()
        dev_ms, launches = 0.0, 0
        for _ in range(steps):
            ms, ln = one_step(resident)
            dev_ms += ms
            launches += ln
        barrier()
        wall_ms = (time.perf_counter() - t0) * 1e3
        if world > 1:
            t = torch.tensor([dev_ms, wall_ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dev_ms, wall_ms = float(t[0]), float(t[1])
        return dev_ms / steps, wall_ms / steps, launches

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_step, wall_step, launches = timed(True, args.steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else None
    stages = {f"spin{0 if c == 1 else 2}_{d}_{k}_ms": v / args.steps for (c, d, k), v in stage_acc.items()
              if k != "_"}

    e2e = None
    if not args.no_e2e:
        e_ms, e_wall, _ = timed(False, max(1, min(args.steps, 2)), 1)
        h2d = sum(h_alm[c].numel() * 8 + h_map[c].numel() * 8 for c in (1, 2))
        e2e = {"value": e_wall, "unit": "ms", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(h2d),
               "device_ms": e_ms, "api": "Plan.alm2map / Plan.adjoint_alm2map (hpsht_alm2map C-ABI) on pinned "
                                         "host shards, per-rank bytes"}

    # ---- roofline of the dominant kernel = the Legendre kernel with the largest share of the step ----
    steps2_c, steps2_d = plan.legendre_steps(2)
    steps0_c, _ = plan.legendre_steps(0)
    import ctypes as C
    tf, secs = C.c_double(), C.c_double()
    H._lib.check(H._lib.lib().hpsht_measure_fp64_peak(C.byref(tf), C.byref(secs)))
    # 26 flop (spin 2) / 6 flop (spin 0) per (l, m, ring-pair) step, culled count
    kernels = {
        "k_leg2alm_s2_w (spin-2 Legendre recurrence, adjoint)": (26.0 * steps2_c, stages.get("spin2_adj_legendre_kernel_ms")),
        "k_alm2leg_s2 (spin-2 Legendre recurrence, synthesis)": (26.0 * steps2_c, stages.get("spin2_a2m_legendre_kernel_ms")),
        "k_leg2alm_s0_w (spin-0 Legendre recurrence, adjoint)": (6.0 * steps0_c, stages.get("spin0_adj_legendre_kernel_ms")),
        "k_alm2leg_s0 (spin-0 Legendre recurrence, synthesis)": (6.0 * steps0_c, stages.get("spin0_a2m_legendre_kernel_ms")),
    }
    dom = max(kernels, key=lambda k: kernels[k][1] or 0.0)
    flop, kern_ms = kernels[dom]
    achieved = _tf(flop, kern_ms)
    # DRAM bytes per launch of that kernel from the committed ncu --set full capture (same size, P = 1)
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "r01_ncu_traffic.json")))
        ent = tj.get(dom.split(" ")[0])
        if ent and ent.get("nside") == nside and ent.get("lmax") == lmax and world == 1:
            traffic = ent["dram_bytes_per_launch"]
    except Exception:
        pass
    roofline = {"bound": "fp64_fma", "kernel": dom,
                "achieved": achieved, "peak": tf.value, "unit": "TFLOP/s",
                "frac": (achieved / tf.value) if achieved else None, "traffic": traffic,
                "share_of_step": (kern_ms / ms_step) if kern_ms else None,
                "peak_source": "in-repo DFMA micro-benchmark hpsht_measure_fp64_peak, same run (MEASURED_PEAKS.json "
                               "has no FP64 entry; nominal 37.2 TF at 1965 MHz)",
                "algorithmic_flop_per_launch": flop, "kernel_ms": kern_ms,
                "other_kernels": {k.split(" ")[0]: {"tflops": _tf(*v), "frac": (_tf(*v) / tf.value) if _tf(*v) else None,
                                                    "kernel_ms": v[1]}
                                  for k, v in kernels.items() if k != dom}}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            v, cores, sample, extra = cpu_port_step_ms(nside, lmax, args.cpu_budget)
            cpu_baseline = {"value": v, "unit": "ms", "cores": cores, "kind": "port", "sample": sample}
        except Exception as e:  # pragma: no cover
            log("cpu baseline failed:", e)

    if rank == 0:
        out = {"metric": METRIC, "value": ms_step, "unit": "ms", "n_gpus": world, "steps": args.steps,
               "warmup": args.warmup, "ms_per_step": ms_step, "wall_ms_per_step": wall_step,
               "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
               "data": "synthetic", "config": config, "clocks": clocks, "gpu_launches": int(launches),
               "e2e": e2e, "roofline": roofline, "cpu_baseline": cpu_baseline, "stages_ms": stages,
               "fp64_peak_tflops_measured": tf.value}
        emit(out)
    plan.destroy()
    if world > 1:
        dist.destroy_process_group()


def _tf(flop, ms):
    return (flop / (ms * 1e-3) / 1e12) if ms else None


if __name__ == "__main__":
    main()
