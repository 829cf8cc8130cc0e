# usage: bash tools/run_multi.sh N  -- multi-rank parity check, bench and host-link benchmark on N GPUs of one box
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
mkdir -p gpurun_out
$TR --nproc-per-node $N --master-port 29601 tools/multi_gpu_check.py > gpurun_out/mgpu_check_p$N.log 2>&1; echo "multi_gpu_check P=$N rc=$?"; grep -E "ALL OK|FAIL" gpurun_out/mgpu_check_p$N.log | tail -3
if [ "$N" -ge 4 ]; then
  HPSHT_HOST_BLOCKS=3 $TR --nproc-per-node 3 --master-port 29602 tools/multi_gpu_check.py > gpurun_out/mgpu_check_p3.log 2>&1; echo "multi_gpu_check P=3 (3 blocks) rc=$?"; grep -E "ALL OK|FAIL" gpurun_out/mgpu_check_p3.log | tail -3
fi
$TR --nproc-per-node $N --master-port 29603 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.log; echo "bench rc=$?"
$TR --nproc-per-node $N --master-port 29604 tools/h2d_bench.py 800 0 2>/dev/null | tail -1
$TR --nproc-per-node $N --master-port 29605 tools/h2d_bench.py 800 1 2>/dev/null | tail -1
python - <<PY
import json
d=json.load(open("gpurun_out/bench_n$N.json"))
print("value",round(d["value"],2),"e2e",round(d["e2e"]["value"],2),"parity",d["parity"]["ok"],d["parity"]["adjointness"],{k:v.get("rel_l2") for k,v in d["parity"]["rings_vs_oracle"].items()})
print("nvlink",d["roofline_nvlink"] and round(d["roofline_nvlink"]["frac"],3),"fft",round(d["roofline_fft"]["frac"],3),"roofline",round(d["roofline"]["frac"],3), d["e2e"].get("host_binding_rank0"))
print("blocked ",{k:round(v,2) for k,v in d["stages_ms"].items() if "total" in k})
print("back2back",{k:round(v,2) for k,v in d["stages_back_to_back_ms"].items() if "total" in k or "exchange" in k})
PY
