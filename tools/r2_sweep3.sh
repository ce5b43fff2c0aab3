#!/bin/bash
# round 2, GPU call 3: single-band / band-slack planning, cell forms, narrow slices, start-up lag, traceback levels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B=anyseq_b200
run() { echo "== $*" ; timeout 900 "$@" 2>&1 | grep -v "^$" ; }
{
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
for lib in _build _build_f0 _build_f1r1; do
  export ANYSEQ_LIB=$PWD/$B/$lib/libanyseq_b200.so
  echo "#### lib $lib full width, band slack 1 / 2 / 4"
  for sl in 1 2 4; do
    SLACK=$sl WL=1.0 REPS=1 run python tools/gpu_perf.py 0 0 1 semiglobal 32 0 2,3
  done
  echo "#### lib $lib 575k slice (single band when strips <= warps)"
  REPS=2 run python tools/gpu_perf.py 4641652 575000 1 semiglobal 32 0 1,2
  REPS=2 run python tools/gpu_perf.py 4641652 575000 1 semiglobal 16 0 2,3
  REPS=2 run python tools/gpu_perf.py 4641652 575000 1 semiglobal 8 0 2,3
  echo "#### lib $lib start-up lag: 575k columns, 40k and 80k rows"
  REPS=3 run python tools/gpu_perf.py 40000 575000 1 semiglobal 32 0 1
  REPS=3 run python tools/gpu_perf.py 80000 575000 1 semiglobal 32 0 1
  REPS=3 run python tools/gpu_perf.py 40000 575000 1 semiglobal 16 0 2
  REPS=3 run python tools/gpu_perf.py 80000 575000 1 semiglobal 16 0 2
  echo "#### lib $lib small"
  REPS=3 run python tools/gpu_perf.py 10000 10000 1 semiglobal 4,8 0 0
  REPS=3 run python tools/gpu_perf.py 100000 100000 1 semiglobal 4,8,16 0 0
  REPS=2 run python tools/gpu_perf.py 1000000 1000000 0 local 16,32 0 0
done
for lib in _build_f1p _build_f0p; do
  export ANYSEQ_LIB=$PWD/$B/$lib/libanyseq_b200.so
  echo "#### profile lib $lib"
  SLACK=2 WL=1.0 REPS=1 run python tools/gpu_perf.py 0 0 1 semiglobal 32 0 2,3
  REPS=1 run python tools/gpu_perf.py 4641652 575000 1 semiglobal 32 0 1
  REPS=1 run python tools/gpu_perf.py 4641652 575000 1 semiglobal 16 0 2
done
unset ANYSEQ_LIB
echo "#### traceback levels (default lib)"
ANYSEQ_TRACE_LEVELS=1 run python tools/gpu_tb_time.py 1000000
} > gpurun_out/r2_sweep3.log 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputest3.log 2>&1
tail -5 gpurun_out/r2_gputest3.log
