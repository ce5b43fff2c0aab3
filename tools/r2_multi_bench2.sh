#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29577"
timeout 240 $TR bench.py --gpus $N --steps 3 --warmup 2 --no-cpu > gpurun_out/r2_bench_n${N}_final2.json 2> gpurun_out/r2_bench_n${N}_final2.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2_bench_n${N}_final2.json"))
    print({k: d[k] for k in ("value", "ms_per_step", "n_gpus", "score", "end_cell")}, d["golden_fullsize"]["equal"], d["stream"]["gcups"], d["e2e"]["value"])
except Exception as e:
    print("bench genome failed:", e)
PY
tail -2 gpurun_out/r2_bench_n${N}_final2.err
