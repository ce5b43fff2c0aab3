#!/bin/bash
# round 2, N-GPU call (N = 4 or 8): functional check of every multi-GPU path, bench lines (whole-genome pair, reads)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-4}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29577"
TB_N=1000000 timeout 400 $TR tools/multi_gpu_check.py > gpurun_out/r2_multi_check_n$N.log 2>&1
grep -E "sharded|ok on|Error|error" gpurun_out/r2_multi_check_n$N.log | tail -8
timeout 400 $TR bench.py --gpus $N --steps 3 --warmup 2 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2_bench_n$N.json"))
    print({k: d[k] for k in ("value", "ms_per_step", "n_gpus", "score", "end_cell")}, d["golden_fullsize"], d["stream"], d["e2e"]["value"], d["config"]["kernel_ms_per_step_max_over_ranks"])
except Exception as e:
    print("bench genome failed:", e)
PY
tail -2 gpurun_out/r2_bench_n$N.err
timeout 400 $TR bench.py --gpus $N --workload reads --steps 2 --warmup 1 --oracle-pairs 20000 > gpurun_out/r2_bench_reads_n$N.json 2> gpurun_out/r2_bench_reads_n$N.err
python - <<PY
import json
try:
    r = json.load(open("gpurun_out/r2_bench_reads_n$N.json"))
    print(r["value"], r["e2e"]["value"], r["per_scheme"], r["oracle_check"])
except Exception as e:
    print("bench reads failed:", e)
PY
tail -2 gpurun_out/r2_bench_reads_n$N.err
