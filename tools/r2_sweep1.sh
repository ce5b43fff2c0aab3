#!/bin/bash
# round 2, first GPU call: parity of the decoupled cell form + timing sweep of cell form x rows/step x K x CTAs/SM
# (full width and the 8-GPU slice width), plus the in-kernel cycle breakdown of the new form.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B=anyseq_b200
run() { echo "== $*" ; timeout 600 "$@" 2>&1 | grep -v "^$" ; }
{
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
for lib in _build _build_f1r1 _build_f0; do
  export ANYSEQ_LIB=$PWD/$B/$lib/libanyseq_b200.so
  echo "#### lib $lib full width"
  WL=1.0 REPS=2 run python tools/gpu_perf.py 0 0 1 semiglobal 32 0 1,2,3
  WL=1.0 REPS=1 run python tools/gpu_perf.py 0 0 1 semiglobal 16 0 2,3
  echo "#### lib $lib 575k slice"
  REPS=2 run python tools/gpu_perf.py 4641652 575000 1 semiglobal 16,32 0 1,2,3
  echo "#### lib $lib 10k x 10k and 100k"
  REPS=3 run python tools/gpu_perf.py 10000 10000 1 semiglobal 4,8 0 1,2
  REPS=3 run python tools/gpu_perf.py 100000 100000 1 semiglobal 4,8,16 0 1,2
  echo "#### lib $lib linear gaps, local"
  WL=1.0 REPS=1 run python tools/gpu_perf.py 0 0 0 semiglobal 32 0 2,3
  WL=1.0 REPS=1 run python tools/gpu_perf.py 0 0 1 local 32 0 2,3
done
for lib in _build_f1p _build_f1r1p; do
  export ANYSEQ_LIB=$PWD/$B/$lib/libanyseq_b200.so
  echo "#### profile lib $lib"
  WL=1.0 REPS=1 run python tools/gpu_perf.py 0 0 1 semiglobal 32 0 1,2
  REPS=1 run python tools/gpu_perf.py 4641652 575000 1 semiglobal 16,32 0 1,2
done
} > gpurun_out/r2_sweep1.log 2>&1
unset ANYSEQ_LIB
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputest1.log 2>&1
tail -5 gpurun_out/r2_gputest1.log
tail -100 gpurun_out/r2_sweep1.log
