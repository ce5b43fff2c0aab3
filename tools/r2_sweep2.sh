#!/bin/bash
# round 2, GPU call 2: cell form x prefetch lead x rows/step sweep (full width, 8-GPU slice width, small problems),
# in-kernel cycle breakdowns, per-level traceback times, parity of the new default
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B=anyseq_b200
run() { echo "== $*" ; timeout 600 "$@" 2>&1 | grep -v "^$" ; }
{
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
for lib in _build _build_f0 _build_f0L16 _build_f1L16 _build_f1r1; do
  export ANYSEQ_LIB=$PWD/$B/$lib/libanyseq_b200.so
  echo "#### lib $lib full width"
  WL=1.0 REPS=1 run python tools/gpu_perf.py 0 0 1 semiglobal 32 0 1,2,3
  echo "#### lib $lib 575k slice"
  REPS=2 run python tools/gpu_perf.py 4641652 575000 1 semiglobal 16 0 1,2
  REPS=2 run python tools/gpu_perf.py 4641652 575000 1 semiglobal 32 0 1
  echo "#### lib $lib small"
  REPS=3 run python tools/gpu_perf.py 10000 10000 1 semiglobal 4 0 0
  REPS=3 run python tools/gpu_perf.py 100000 100000 1 semiglobal 8 0 0
done
for lib in _build_f1p _build_f0p _build_f1r1p; do
  export ANYSEQ_LIB=$PWD/$B/$lib/libanyseq_b200.so
  echo "#### profile lib $lib"
  WL=1.0 REPS=1 run python tools/gpu_perf.py 0 0 1 semiglobal 32 0 1,2,3
  REPS=1 run python tools/gpu_perf.py 4641652 575000 1 semiglobal 16 0 1,2
  REPS=1 run python tools/gpu_perf.py 4641652 575000 1 semiglobal 32 0 1
done
unset ANYSEQ_LIB
echo "#### traceback levels (default lib)"
ANYSEQ_TRACE_LEVELS=1 run python tools/gpu_tb_time.py 1000000
} > gpurun_out/r2_sweep2.log 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputest2.log 2>&1
tail -5 gpurun_out/r2_gputest2.log
