#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputest18.log 2>&1
tail -3 gpurun_out/r2_gputest18.log
REPS=2 timeout 300 python tools/gpu_perf.py 4641652 575000 1 semiglobal 0 0 0 | grep GCUPS
REPS=2 timeout 300 python tools/gpu_perf.py 4641652 578752 1 semiglobal 0 0 0 | grep GCUPS
REPS=2 timeout 300 python tools/gpu_perf.py 4641652 575488 1 semiglobal 0 0 0 | grep GCUPS
WL=1.0 REPS=2 timeout 300 python tools/gpu_perf.py 0 0 1 semiglobal 0 0 0 | grep GCUPS
REPS=3 timeout 300 python tools/gpu_perf.py 8087 9011 0 global 0 0 0 | grep GCUPS
