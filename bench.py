#!/usr/bin/env python
"""bench.py — the reference's headline benchmark on synthetic data.

    python bench.py --gpus N --steps K --warmup W        (N>1: launched under torchrun, one rank/GPU)

One STEP = the workload of one BASELINE.json config on DAlm{RR}/DMap{RR} shards spread over the N ranks
(strong scaling: the problem is fixed).  Default = the headline, configs[3]: spin-0 alm2map! + spin-0
adjoint_alm2map! + spin-2 alm2map! + spin-2 adjoint_alm2map! at nside=4096, lmax=8191 (fits one B200, so
N=1 runs the same workload).  --config 2 | 3 | 5 selects the other GPU configs of BASELINE.json.

  value         ms per step, shards resident in HBM when the timed region starts (CUDA events on the plan's
                stream, max over ranks)
  e2e           ms per step through the public API (Plan.alm2map / Plan.adjoint_alm2map) with HOST (pinned)
                buffers: H2D of the inputs and D2H of the results inside the timed region
  roofline      the Legendre kernel with the largest share of the step against the FP64 FMA pipe;
                roofline_fft: the ring-FFT stage against HBM; roofline_nvlink (N>1): the leg exchange
                against NVLink -- the last two from a short extra pass with the stages run back to back
                (in the timed region they overlap block by block)
  parity        after the timed region: the adjointness identity <f, Y a> = <Y^T f, a>_w over all ranks, and
                a sample of rank 0's rings of Y a against the long-double oracle (checker only)
  cpu_baseline  the FP64 OpenMP CPU port (oracle/cpu_port.*; kind "port": the real reference needs
                Julia+MPI+Ducc0, absent here) timed on a bounded sample of the same step

--impl reference times that CPU port alone (rank 0 only) and prints the same JSON shape.
Smaller problems for smoke runs: --nside/--lmax.
"""
from __future__ import annotations

import argparse
import ctypes as C
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

# BASELINE.json configs -> (nside, lmax, [(ncomp, direction)])
CONFIGS = {
    2: (1024, 3071, [(1, "a2m"), (1, "adj")]),
    3: (2048, 6143, [(1, "a2m"), (2, "a2m")]),
    4: (4096, 8191, [(1, "a2m"), (1, "adj"), (2, "a2m"), (2, "adj")]),
    5: (8192, 16383, [(1, "adj")]),
}
NVLINK_GBS = 900.0  # NVLink 5, per direction per GPU (B200_PROFILING.md)


def log(*a):
    print(*a, file=sys.stderr, flush=True)


_FULL_AFFINITY = None


def bind_to_gpu_numa_node(torch, local_rank):
    """Launcher-side policy for the host-pointer (e2e) path: run this rank -- and first-touch its pinned buffers --
    on the NUMA node its GPU hangs off (all ranks on node 0 was the round-1 default)."""
    try:
        pr = torch.cuda.get_device_properties(local_rank)
        bus = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        if node < 0:
            return {"numa_node": None, "pci": bus}
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        global _FULL_AFFINITY
        _FULL_AFFINITY = os.sched_getaffinity(0)
        cpus &= _FULL_AFFINITY
        if cpus:
            os.sched_setaffinity(0, cpus)
        return {"numa_node": node, "pci": bus, "cpus_bound": len(cpus)}
    except Exception as e:  # pragma: no cover
        return {"numa_node": None, "error": str(e)}


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
# CPU baseline (port): a bounded sample of the step, every stage in OpenMP C++
# ------------------------------------------------------------------------------------------------
class CpuPort:
    """The FP64 OpenMP port (oracle/cpu_port.cpp: Legendre recurrences + ring FFTs) on all host cores.  A step
    is timed on every `stride`-th m (Legendre) and every `stride`-th ring (ring stage) and scaled by the work
    model; stride 1 = the whole step, nothing extrapolated."""

    def __init__(self, nside, lmax, transforms):
        from oracle import cpu_port as P
        from oracle import oracle as O

        self.P = P
        self.nside, self.lmax, self.transforms = nside, lmax, transforms
        if _FULL_AFFINITY:  # the GPU arm bound this process to one NUMA node: the CPU arm gets the whole machine back
            os.sched_setaffinity(0, _FULL_AFFINITY)
        # torchrun exports OMP_NUM_THREADS=1: ask for the machine explicitly and report what OpenMP grants
        P.set_threads(os.cpu_count() or 1)
        self.threads = P.max_threads()
        nr = 4 * nside - 1
        self.nr = nr
        self.nphi, self.theta, self.phi0, _ = O.healpix_rings(nside, np.arange(1, nr + 1))
        self.all_m = np.arange(lmax + 1)
        self.mstart = O.rr_mstart1(lmax, 1, self.all_m) - 1
        self.nalm = (lmax + 1) * (lmax + 2) // 2
        self.full = {s: P.legendre_steps(s, lmax, self.all_m, self.theta)[0] for s in (0, 2)}
        self.rng = np.random.default_rng(0)
        self.rate = None  # spin-0 steps / s, calibrated once

    def _calibrate(self):
        mval = self.all_m[::max(1, (self.lmax + 1) // 16)]
        alm = self.rng.standard_normal((1, self.nalm)) + 1j * self.rng.standard_normal((1, self.nalm))
        self.P.alm2leg(alm, 0, self.lmax, mval, self.mstart[mval], self.theta)  # cold pass: threads, pages
        t0 = time.perf_counter()
        self.P.alm2leg(alm, 0, self.lmax, mval, self.mstart[mval], self.theta)
        dt = time.perf_counter() - t0
        self.rate = self.P.legendre_steps(0, self.lmax, mval, self.theta)[0] / max(dt, 1e-9)

    def stride_for(self, budget_s):
        """Sampling stride that makes one step_ms() call take about budget_s: a first guess from the calibrated
        rate, then a trial pass at four times that stride (it also warms threads, pages and FFT plans) whose wall
        time, input generation included, is scaled up."""
        if self.rate is None:
            self._calibrate()
        est = sum(self.full[0 if c == 1 else 2] * (1.0 if c == 1 else 4.3) for c, _ in self.transforms) / self.rate
        s1 = int(max(1, np.ceil(1.15 * est / max(budget_s, 1e-3))))
        s, trial = s1, 4 * s1
        for _ in range(3):   # the sparser the trial, the more its fixed costs inflate the estimate: refine
            t0 = time.perf_counter()
            self.step_ms(trial)
            full = (time.perf_counter() - t0) * trial
            s = int(max(1, np.ceil(1.2 * full / max(budget_s, 1e-3))))
            if 2 * s >= trial:
                break
            trial = 2 * s
        return s

    def step_ms(self, stride):
        P, lmax = self.P, self.lmax
        mval = self.all_m[::stride]
        rsel = np.arange(0, self.nr, stride)
        nphi, phi0 = self.nphi[rsel], self.phi0[rsel]
        rs0 = np.concatenate([[0], np.cumsum(nphi)[:-1]]).astype(np.int64)
        npix = int(nphi.sum())
        P.prepare_rings(nphi)  # FFT plans are built once and kept, as in any FFT library: not part of a step
        tot = leg_s = ring_s = 0.0
        for ncomp, d in self.transforms:
            spin = 0 if ncomp == 1 else 2
            scale_m = self.full[spin] / P.legendre_steps(spin, lmax, mval, self.theta)[0]
            scale_r = float(self.nphi.sum()) / npix
            if d == "a2m":
                alm = self.rng.standard_normal((ncomp, self.nalm)) + 1j * self.rng.standard_normal((ncomp, self.nalm))
                t0 = time.perf_counter()
                P.alm2leg(alm, spin, lmax, mval, self.mstart[mval], self.theta)
                t1 = time.perf_counter()
                legr = self.rng.standard_normal((ncomp, len(rsel), lmax + 1)) + 0j
                t2 = time.perf_counter()
                P.leg2map(legr, nphi, phi0, rs0, npix)
                t3 = time.perf_counter()
            else:
                mp = self.rng.standard_normal((ncomp, npix))
                t2 = time.perf_counter()
                P.map2leg(mp, lmax + 1, nphi, phi0, rs0)
                t3 = time.perf_counter()
                leg = self.rng.standard_normal((ncomp, self.nr, len(mval))) + 0j
                t0 = time.perf_counter()
                P.leg2alm(leg, spin, lmax, mval, self.mstart[mval], self.theta, self.nalm)
                t1 = time.perf_counter()
            leg_s += (t1 - t0) * scale_m
            ring_s += (t3 - t2) * scale_r
        tot = (leg_s + ring_s) * 1e3
        return tot, leg_s * 1e3, ring_s * 1e3

    def describe(self, stride, leg_ms, ring_ms):
        what = "the whole step, nothing extrapolated" if stride == 1 else (
            f"every {stride}-th m (Legendre) and every {stride}-th ring (ring FFT) of the step, scaled by the "
            f"work model (steps after the mlim cut / pixels)")
        return (f"{what}; OpenMP {self.threads} threads (omp_get_max_threads; host has {os.cpu_count()} CPUs); "
                f"Legendre {leg_ms:.0f} ms + ring stage {ring_ms:.0f} ms per step")


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=4, choices=sorted(CONFIGS))
    ap.add_argument("--nside", type=int, default=None)
    ap.add_argument("--lmax", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--cpu-budget", type=float, default=20.0,
                    help="seconds of CPU work for the cpu_baseline sample of the default arm")
    ap.add_argument("--ref-budget", type=float, default=150.0,
                    help="seconds of CPU work for the whole --impl reference run (all steps)")
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
    nside, lmax, transforms = CONFIGS[args.config]
    nside = args.nside or nside
    lmax = args.lmax or lmax
    tnames = ", ".join(f"spin-{0 if c == 1 else 2} {'alm2map!' if d == 'a2m' else 'adjoint_alm2map!'}" for c, d in transforms)
    workload = f"BASELINE config {args.config}: {tnames} at nside={nside} lmax={lmax}, DAlm{{RR}}/DMap{{RR}} over {world} rank(s)"
    config = {"workload": workload, "nside": nside, "lmax": lmax, "ranks": world, "baseline_config": args.config,
              "inputs": "counter-based N(0,1) alm / maps (Philox), larger than L2 (no flush needed)"}
    metric = METRIC if args.config == 4 else f"ms per step, {tnames} at nside={nside} lmax={lmax}"

    if args.impl == "reference":
        if rank != 0:
            return
        cpu = CpuPort(nside, lmax, transforms)
        stride = cpu.stride_for(args.ref_budget / max(1, args.steps + args.warmup))
        vals, leg_ms, ring_ms = [], 0.0, 0.0
        for i in range(args.warmup + args.steps):
            ms, leg_ms, ring_ms = cpu.step_ms(stride)
            if i >= args.warmup:
                vals.append(ms)
        v = float(np.mean(vals))
        emit({
            "impl": "reference", "metric": metric, "value": v, "unit": "ms", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": v, "higher_is_better": False,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
            "cpu_baseline": {"value": v, "unit": "ms", "cores": cpu.threads, "kind": "port", "extrapolated": stride > 1,
                             "sample_stride": stride, "sample": cpu.describe(stride, leg_ms, ring_ms)},
            "e2e": {"value": v, "unit": "ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "FP64 OpenMP CPU port of the same staged algorithm (NOT ducc: the reference itself needs "
                    "Julia+MPI+Ducc0, absent from this image; bench/reference_cpu.jl times the real one where "
                    "Julia exists).  Expect ducc's hand-vectorised kernels to be faster than this port."})
        return

    import torch
    import torch.distributed as dist

    import healpixmpi_jl_b200 as H

    torch.cuda.set_device(local_rank)
    host_binding = bind_to_gpu_numa_node(torch, local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        comm = H.TorchComm()
    else:
        comm = H.COMM_SELF

    ncomps = sorted(set(c for c, _ in transforms))
    t_setup = time.perf_counter()
    plan = H.Plan(nside, lmax, comm, max(ncomps), local_rank)
    mval, _ = plan.alm_info()
    minfo = plan.map_info()
    log(f"[rank {rank}] plan: nm_loc={plan.nm_loc} nalm_loc={plan.nalm_loc} nring_loc={plan.nring_loc} "
        f"npix_loc={plan.npix_loc}")

    # ---- shards: host (pinned) and device copies ----
    def pinned(a):
        t = torch.from_numpy(np.ascontiguousarray(a))
        return t.pin_memory()

    seed0 = 4096 * args.config
    h_alm = {c: pinned(synth_alm_shard(seed0 + c - 1, c, lmax, mval).view(np.float64)) for c in ncomps}
    h_map = {c: pinned(synth_map_shard(seed0 + c + 1, c, minfo["rings"], minfo["nphi"])) for c in ncomps}
    d_alm = {k: v.cuda() for k, v in h_alm.items()}
    d_map = {k: v.cuda() for k, v in h_map.items()}
    d_alm_out = {k: torch.empty_like(v) for k, v in d_alm.items()}
    d_map_out = {k: torch.empty_like(v) for k, v in d_map.items()}
    h_alm_out = {k: torch.empty_like(v).pin_memory() for k, v in h_alm.items()}
    h_map_out = {k: torch.empty_like(v).pin_memory() for k, v in h_map.items()}
    log(f"[rank {rank}] setup {time.perf_counter() - t_setup:.1f}s")

    stage_acc = {}

    def run_one(ncomp, d, resident):
        if d == "a2m":
            plan.alm2map(d_alm[ncomp] if resident else h_alm[ncomp],
                         d_map_out[ncomp] if resident else h_map_out[ncomp], ncomp)
        else:
            plan.adjoint_alm2map(d_map[ncomp] if resident else h_map[ncomp],
                                 d_alm_out[ncomp] if resident else h_alm_out[ncomp], ncomp)
        t = plan.timings()
        for k, v in t.items():
            stage_acc[(ncomp, d, k)] = stage_acc.get((ncomp, d, k), 0.0) + v
        return t["total"], plan.launches()

    def one_step(resident: bool):
        """all transforms of the config; returns (device ms summed from the plan's CUDA events, kernel launches)"""
        ms, launches = 0.0, 0
        for ncomp, d in transforms:
            a, b = run_one(ncomp, d, resident)
            ms += a
            launches += b
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

    def stage_table(steps):
        return {f"spin{0 if c == 1 else 2}_{d}_{k}_ms": v / steps for (c, d, k), v in stage_acc.items() if k != "_"}

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_step, wall_step, launches = timed(True, args.steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else None
    stages = stage_table(args.steps)

    e2e = None
    if not args.no_e2e:
        e_ms, e_wall, _ = timed(False, max(1, min(args.steps, 2)), 1)
        h2d = sum((h_alm[c].numel() if d == "a2m" else h_map[c].numel()) * 8 for c, d in transforms)
        d2h = sum((h_map[c].numel() if d == "a2m" else h_alm[c].numel()) * 8 for c, d in transforms)
        e2e = {"value": e_wall, "unit": "ms", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "device_ms": e_ms, "host_binding_rank0": host_binding, "api": "Plan.alm2map / Plan.adjoint_alm2map (hpsht_alm2map C-ABI) on pinned "
                                         "host shards, per-rank bytes"}

    # ---- stage pass: the same transforms with the stages back to back (no block overlap), for the per-stage
    #      rooflines.  At N = 1 the timed region already runs that way. ----
    stages_seq = stages
    if world > 1:
        plan.set_dev_blocks(False)
        timed(True, 2, 1)
        stages_seq = stage_table(2)
        plan.set_dev_blocks(True)

    # ---- roofline of the dominant kernel = the Legendre kernel with the largest share of the step ----
    tf, secs = C.c_double(), C.c_double()
    H._lib.check(H._lib.lib().hpsht_measure_fp64_peak(C.byref(tf), C.byref(secs)))
    steps_c = {c: plan.legendre_steps(0 if c == 1 else 2)[0] for c in ncomps}
    kname = {(1, "a2m"): "k_alm2leg_s0", (1, "adj"): "k_leg2alm_s0", (2, "a2m"): "k_alm2leg_s2", (2, "adj"): "k_leg2alm_s2"}
    # 26 flop (spin 2) / 6 flop (spin 0) per (l, m, ring-pair) step, culled count, this rank
    kernels = {kname[(c, d)]: ((26.0 if c == 2 else 6.0) * steps_c[c],
                               stages_seq.get(f"spin{0 if c == 1 else 2}_{d}_legendre_kernel_ms")) for c, d in transforms}
    dom = max(kernels, key=lambda k: kernels[k][1] or 0.0)
    flop, kern_ms = kernels[dom]
    achieved = _tf(flop, kern_ms)
    # DRAM bytes per launch of that kernel from the committed ncu --set full capture (same size, P = 1)
    traffic, traffic_src = None, None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")))
        ent = tj.get(dom)
        if ent and ent.get("nside") == nside and ent.get("lmax") == lmax and world == 1:
            traffic = ent["dram_bytes_per_launch"]
            traffic_src = ent.get("source")
    except Exception:
        pass
    roofline = {"bound": "fp64_fma", "kernel": dom + " (+ its polar / equatorial difference-form launches)",
                "achieved": achieved, "peak": tf.value, "unit": "TFLOP/s",
                "frac": (achieved / tf.value) if achieved else None, "traffic": traffic, "traffic_source": traffic_src,
                "share_of_step": (kern_ms / ms_step) if kern_ms else None,
                "peak_source": "in-repo DFMA micro-benchmark hpsht_measure_fp64_peak, same run (MEASURED_PEAKS.json "
                               "has no FP64 entry; nominal 37.2 TF at 1965 MHz)",
                "algorithmic_flop_per_launch": flop, "kernel_ms": kern_ms,
                "other_kernels": {k: {"tflops": _tf(*v), "frac": (_tf(*v) / tf.value) if _tf(*v) else None,
                                      "kernel_ms": v[1]} for k, v in kernels.items() if k != dom}}

    # ---- ring FFT against HBM, leg exchange against NVLink ----
    hbm = 6457.1
    try:
        hbm = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    fft_lines = {}
    for c, d in transforms:
        ms = stages_seq.get(f"spin{0 if c == 1 else 2}_{d}_ringfft_ms")
        by = c * (16.0 * (lmax + 1) * plan.nring_loc + 8.0 * plan.npix_loc)  # SURVEY 8d, this rank
        if ms:
            fft_lines[f"spin{0 if c == 1 else 2}_{d}"] = {"ms": ms, "algorithmic_bytes": by, "gbs": by / ms / 1e6,
                                                          "frac": by / ms / 1e6 / hbm}
    worst = min(fft_lines, key=lambda k: fft_lines[k]["frac"]) if fft_lines else None
    roofline_fft = None
    if worst:
        roofline_fft = {"bound": "hbm", "kernel": "k_leg2map / k_map2leg (ring FFT stage, " + worst + ")",
                        "achieved": fft_lines[worst]["gbs"], "peak": hbm, "unit": "GB/s", "frac": fft_lines[worst]["frac"],
                        "peak_source": "MEASURED_PEAKS.json hbm_gbs (copy bandwidth)", "per_transform": fft_lines}
    roofline_nvlink = None
    if world > 1:
        nrt = 4 * nside - 1
        x_lines = {}
        for c, d in transforms:
            ms = stages_seq.get(f"spin{0 if c == 1 else 2}_{d}_exchange_ms")
            by = c * 16.0 * plan.nm_loc * (nrt - plan.nring_loc)  # egress of this rank, SURVEY 8d
            if ms:
                x_lines[f"spin{0 if c == 1 else 2}_{d}"] = {"ms": ms, "egress_bytes": by, "gbs": by / ms / 1e6,
                                                            "frac": by / ms / 1e6 / NVLINK_GBS}
        if x_lines:
            w = min(x_lines, key=lambda k: x_lines[k]["frac"])
            roofline_nvlink = {"bound": "nvlink", "kernel": "grouped ncclSend/ncclRecv + k_pack / k_unpack (" + w + ")",
                               "achieved": x_lines[w]["gbs"], "peak": NVLINK_GBS, "unit": "GB/s", "frac": x_lines[w]["frac"],
                               "peak_source": "NVLink 5 nominal, per direction per GPU", "per_transform": x_lines,
                               "note": "measured with the stages back to back (HPSHT_DEV_BLOCKS off); in the timed "
                                       "region the exchange of a ring block runs behind the next block's Legendre kernels"}

    # ---- parity (after the timed region; the oracle is the checker only) ----
    parity = None
    if not args.no_parity:
        parity = parity_block(plan, H, torch, dist, world, rank, nside, lmax, ncomps, seed0, mval, minfo,
                              d_alm, d_map, d_alm_out, d_map_out)

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            cpu = CpuPort(nside, lmax, transforms)
            stride = cpu.stride_for(args.cpu_budget)
            v, leg_ms, ring_ms = cpu.step_ms(stride)
            cpu_baseline = {"value": v, "unit": "ms", "cores": cpu.threads, "kind": "port", "extrapolated": stride > 1,
                            "sample_stride": stride, "sample": cpu.describe(stride, leg_ms, ring_ms)}
        except Exception as e:  # pragma: no cover
            log("cpu baseline failed:", e)

    if rank == 0:
        out = {"metric": metric, "value": ms_step, "unit": "ms", "n_gpus": world, "steps": args.steps,
               "warmup": args.warmup, "ms_per_step": ms_step, "wall_ms_per_step": wall_step,
               "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
               "data": "synthetic", "config": config, "clocks": clocks, "gpu_launches": int(launches),
               "e2e": e2e, "roofline": roofline, "roofline_fft": roofline_fft, "roofline_nvlink": roofline_nvlink,
               "parity": parity, "cpu_baseline": cpu_baseline, "stages_ms": stages,
               "stages_back_to_back_ms": stages_seq if world > 1 else None, "fp64_peak_tflops_measured": tf.value}
        emit(out)
    plan.destroy()
    if world > 1:
        dist.destroy_process_group()


def parity_block(plan, H, torch, dist, world, rank, nside, lmax, ncomps, seed0, mval, minfo, d_alm, d_map,
                 d_alm_out, d_map_out):
    """(1) adjointness <f, Y a> = <Y^T f, a>_w (localdot weights, src/alm.jl:634-644) with both sides summed over
    all ranks; (2) a sample of rank 0's rings of Y a against the long-double oracle on the regenerated full alm."""
    out = {"tolerance": 1e-12, "adjointness": {}, "rings_vs_oracle": {}}
    w = torch.from_numpy(np.concatenate([np.full(lmax + 1 - int(m), 1.0 if m == 0 else 2.0) for m in mval])).cuda()
    ok = True
    for c in ncomps:
        spin = 0 if c == 1 else 2
        plan.alm2map(d_alm[c], d_map_out[c], c)
        plan.adjoint_alm2map(d_map[c], d_alm_out[c], c)
        a = d_alm[c].view(c, -1, 2)
        o = d_alm_out[c].view(c, -1, 2)
        v = torch.stack([(d_map[c] * d_map_out[c]).sum(), (w[None, :, None] * o * a).sum(),
                         (d_map[c] ** 2).sum(), (d_map_out[c] ** 2).sum()])
        if world > 1:
            dist.all_reduce(v)
        lhs, rhs, n1, n2 = (float(x) for x in v)
        res = abs(lhs - rhs) / np.sqrt(n1 * n2)
        out["adjointness"][f"spin{spin}"] = res
        ok = ok and res <= 1e-12
        if rank == 0:
            try:
                from oracle import parity as PR

                sel = PR.sample_rings(nside, minfo["rings"], n_bulk=4)
                sel = sel[np.linspace(0, len(sel) - 1, min(len(sel), 10)).astype(int)]
                loc = {int(r): (int(s), int(n)) for r, s, n in zip(minfo["rings"], minfo["rstart0"], minfo["nphi"])}
                mp = d_map_out[c]
                got = {int(r): mp[:, loc[int(r)][0]:loc[int(r)][0] + loc[int(r)][1]].cpu().numpy() for r in sel}
                alm_full = synth_alm_shard(seed0 + c - 1, c, lmax, np.arange(lmax + 1))
                res = PR.synthesis_parity(nside, lmax, spin, alm_full, got, sel)
                out["rings_vs_oracle"][f"spin{spin}"] = {k: res[k] for k in ("rel_l2", "max_ring", "max_polar_ring",
                                                                             "n_rings", "n_pixels")}
                ok = ok and res["rel_l2"] <= 1e-12
            except Exception as e:  # pragma: no cover
                out["rings_vs_oracle"][f"spin{spin}"] = {"error": str(e)}
                ok = False
    out["rel_l2"] = max([v["rel_l2"] for v in out["rings_vs_oracle"].values() if "rel_l2" in v and v["rel_l2"]] or [None]) \
        if out["rings_vs_oracle"] else None
    out["ok"] = bool(ok)
    out["what"] = ("adjointness: |<f,Ya> - <Y^T f,a>_w| / (|f| |Ya|), all ranks; rings_vs_oracle: relative L2 of rank 0's "
                   "sampled rings of Y a against oracle/sht_oracle.c (long double, __float128 on the rings next to the poles; reference ring cut and pair "
                   "colatitude applied, oracle/parity.py)")
    return out


def _tf(flop, ms):
    return (flop / (ms * 1e-3) / 1e12) if ms else None


if __name__ == "__main__":
    main()
