#!/bin/bash
# round 2, GPU call 5: one CTA per SM with adjacent strips per scheduler; slices; bench lines
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { echo "== $*" ; timeout 900 "$@" 2>&1 | grep -v "^$" ; }
{
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
echo "#### full width, automatic; then cell form 1 with 2 and 3 warps per scheduler"
WL=1.0 REPS=2 run python tools/gpu_perf.py 0 0 1 semiglobal 0 0 0
ANYSEQ_CELL_FORM=1 WL=1.0 REPS=1 run python tools/gpu_perf.py 0 0 1 semiglobal 32 0 2,3
ANYSEQ_CELL_FORM=0 WL=1.0 REPS=1 run python tools/gpu_perf.py 0 0 1 semiglobal 32 0 2
WL=1.0 REPS=1 run python tools/gpu_perf.py 0 0 0 semiglobal 32 0 2,3
for form in 1 0; do
  export ANYSEQ_CELL_FORM=$form
  echo "#### cell form $form: 575488-column slice"
  REPS=2 run python tools/gpu_perf.py 4641652 575488 1 semiglobal 8 0 2,3
  REPS=2 run python tools/gpu_perf.py 4641652 575488 1 semiglobal 16 0 2,3
  REPS=2 run python tools/gpu_perf.py 4641652 575488 1 semiglobal 32 0 1
  echo "#### cell form $form: start-up lag"
  REPS=3 run python tools/gpu_perf.py 40000 575488 1 semiglobal 16 0 2
  REPS=3 run python tools/gpu_perf.py 80000 575488 1 semiglobal 16 0 2
  echo "#### cell form $form: 1150976 and 1M x 1M"
  REPS=2 run python tools/gpu_perf.py 4641652 1150976 1 semiglobal 16,32 0 0
  REPS=2 run python tools/gpu_perf.py 1000000 1000000 1 semiglobal 8,16 0 0
done
unset ANYSEQ_CELL_FORM
echo "#### traceback"
ANYSEQ_TRACE_LEVELS=1 run python tools/gpu_tb_time.py 1000000
} > gpurun_out/r2_sweep5.log 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputest5.log 2>&1
tail -5 gpurun_out/r2_gputest5.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench_n1_b.json 2> gpurun_out/r2_bench_n1_b.err
tail -c 2500 gpurun_out/r2_bench_n1_b.json; tail -3 gpurun_out/r2_bench_n1_b.err
