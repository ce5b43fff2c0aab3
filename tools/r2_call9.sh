#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { echo "== $*" ; timeout 900 "$@" 2>&1 | grep -v "^$" ; }
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputest9.log 2>&1
tail -3 gpurun_out/r2_gputest9.log
{
WL=1.0 REPS=2 run python tools/gpu_perf.py 0 0 1 semiglobal 0 0 0
WL=1.0 REPS=1 run python tools/gpu_perf.py 0 0 1 local 0 0 0
WL=1.0 REPS=1 run python tools/gpu_perf.py 0 0 0 semiglobal 0 0 0
REPS=2 run python tools/gpu_perf.py 4641652 623616 1 semiglobal 0 0 0
REPS=2 run python tools/gpu_perf.py 4641652 575488 1 semiglobal 0 0 0
REPS=2 run python tools/gpu_perf.py 4641652 528576 1 semiglobal 0 0 0
REPS=2 run python tools/gpu_perf.py 4641652 1190912 1 semiglobal 0 0 0
REPS=2 run python tools/gpu_perf.py 4641652 2354176 1 semiglobal 0 0 0
REPS=3 run python tools/gpu_perf.py 1000000 1000000 1 semiglobal 0 0 0
REPS=3 run python tools/gpu_perf.py 100000 100000 1 semiglobal 0 0 0
REPS=3 run python tools/gpu_perf.py 8087 9011 0 global 0 0 0
run python tools/gpu_tb_time.py 1000000
} > gpurun_out/r2_sweep9.log 2>&1
grep -v "^==" gpurun_out/r2_sweep9.log
