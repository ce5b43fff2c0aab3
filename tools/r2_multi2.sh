#!/bin/bash
# round 2, 2-GPU call: functional check of every multi-GPU path (incl. the sharded traceback at 1 Mbp against the frozen
# CPU sha), bench lines at N = 2 (whole-genome pair, reads), parity suite on the new defaults
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29577"
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputest_m$N.log 2>&1
tail -3 gpurun_out/r2_gputest_m$N.log
TB_N=1000000 timeout 900 $TR tools/multi_gpu_check.py > gpurun_out/r2_multi_check_n$N.log 2>&1
grep -E "sharded|ok on|Error|error" gpurun_out/r2_multi_check_n$N.log | tail -8
timeout 900 $TR bench.py --gpus $N --steps 3 --warmup 2 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err
tail -c 1800 gpurun_out/r2_bench_n$N.json; tail -2 gpurun_out/r2_bench_n$N.err
timeout 900 $TR bench.py --gpus $N --workload reads --steps 2 --warmup 1 --oracle-pairs 20000 > gpurun_out/r2_bench_reads_n$N.json 2> gpurun_out/r2_bench_reads_n$N.err
tail -c 1500 gpurun_out/r2_bench_reads_n$N.json; tail -2 gpurun_out/r2_bench_reads_n$N.err
