#!/bin/bash
# round 2, GPU call 8: mixed cell form (even rows coupled, odd rows decoupled)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { echo "== $*" ; timeout 900 "$@" 2>&1 | grep -v "^$" ; }
timeout 1500 python -m pytest tests -m gpu -x -q -k "cell_forms or sweep or mismatch or medium" > gpurun_out/r2_gputest8.log 2>&1
tail -3 gpurun_out/r2_gputest8.log
{
for form in 2 1; do
  export ANYSEQ_CELL_FORM=$form
  echo "#### cell form $form"
  WL=1.0 REPS=2 run python tools/gpu_perf.py 0 0 1 semiglobal 32 0 3
  WL=1.0 REPS=1 run python tools/gpu_perf.py 0 0 1 semiglobal 32 0 2
  WL=1.0 REPS=1 run python tools/gpu_perf.py 0 0 1 local 32 0 3
  REPS=2 run python tools/gpu_perf.py 4641652 575488 1 semiglobal 16 0 2
  REPS=2 run python tools/gpu_perf.py 4641652 575488 1 semiglobal 32 0 1
  REPS=2 run python tools/gpu_perf.py 4641652 1150976 1 semiglobal 32 0 0
  REPS=2 run python tools/gpu_perf.py 1000000 1000000 1 semiglobal 16 0 0
done
} > gpurun_out/r2_sweep8.log 2>&1
grep -v "^==" gpurun_out/r2_sweep8.log
