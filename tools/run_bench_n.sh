# usage: bash tools/run_bench_n.sh N [extra bench args]: bench.py on N GPUs, prints the main numbers
N=${1:-2}; shift
python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node $N --master-port 29613 bench.py --gpus $N --steps 5 --warmup 3 "$@" > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.log; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/bench_n$N.json"))
print("value",round(d["value"],2),"e2e",round(d["e2e"]["value"],2) if d["e2e"] else None,"parity",d["parity"] and d["parity"]["ok"])
print({k:round(v,2) for k,v in d["stages_ms"].items() if "total" in k})
PY
