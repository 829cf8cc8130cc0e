"""Host-link roofline for the e2e (host-pointer) path: every rank copies pinned host memory to its GPU and back,
all ranks at once, as plain cudaMemcpyAsync.   torchrun --nproc-per-node N tools/h2d_bench.py [MB] [bind]
   bind = 1: bind each rank to its GPU's NUMA node first (what bench.py does)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from bench import bind_to_gpu_numa_node

rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); lr = int(os.environ.get("LOCAL_RANK", 0))
mb = int(sys.argv[1]) if len(sys.argv) > 1 else 800
bind = len(sys.argv) > 2 and sys.argv[2] == "1"
torch.cuda.set_device(lr)
info = bind_to_gpu_numa_node(torch, lr) if bind else None
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
h = torch.empty(mb * 1024 * 1024 // 8, dtype=torch.float64).pin_memory()
h.fill_(1.0)
d = torch.empty_like(h, device="cuda")
h2 = torch.empty_like(h).pin_memory()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(both):
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(4):
        with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
        if both:
            with torch.cuda.stream(s2): h2.copy_(d, non_blocking=True)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / 4
run(False); t_h2d = run(False); t_both = run(True)
res = torch.tensor([mb / 1024 / t_h2d, mb / 1024 / t_both], device="cuda")
if world > 1:
    mn = res.clone(); dist.all_reduce(mn, op=dist.ReduceOp.MIN)
    sm = res.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
else:
    mn = sm = res
if rank == 0:
    print(f"{world} rank(s), {mb} MB pinned per direction, bind={bind} {info}: H2D alone min/rank {float(mn[0]):.1f} GB/s (sum {float(sm[0]):.1f}); "
          f"H2D with concurrent D2H min/rank {float(mn[1]):.1f} GB/s per direction (sum {float(sm[1]):.1f})", flush=True)
if world > 1: dist.destroy_process_group()
