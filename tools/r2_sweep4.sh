#!/bin/bash
# round 2, GPU call 4: ragged-strip fix, per-launch cell form, slices, traceback K, bench lines (genome + reads)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { echo "== $*" ; timeout 900 "$@" 2>&1 | grep -v "^$" ; }
{
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
echo "#### full width, everything automatic"
WL=1.0 REPS=2 run python tools/gpu_perf.py 0 0 1 semiglobal 0 0 0
WL=1.0 REPS=1 run python tools/gpu_perf.py 0 0 1 local 0 0 0
WL=1.0 REPS=1 run python tools/gpu_perf.py 0 0 0 semiglobal 0 0 0
for form in 1 0; do
  export ANYSEQ_CELL_FORM=$form
  echo "#### cell form $form: 575k slice"
  REPS=2 run python tools/gpu_perf.py 4641652 575000 1 semiglobal 8,16,32 0 0
  REPS=2 run python tools/gpu_perf.py 4641652 575488 1 semiglobal 16,32 0 0
  echo "#### cell form $form: start-up lag, 575488 columns"
  REPS=3 run python tools/gpu_perf.py 40000 575488 1 semiglobal 16,32 0 0
  REPS=3 run python tools/gpu_perf.py 80000 575488 1 semiglobal 16,32 0 0
  echo "#### cell form $form: other widths"
  REPS=2 run python tools/gpu_perf.py 4641652 1150000 1 semiglobal 16,32 0 0
  REPS=2 run python tools/gpu_perf.py 4641652 2300000 1 semiglobal 16,32 0 0
  REPS=3 run python tools/gpu_perf.py 1000000 1000000 1 semiglobal 8,16,32 0 0
  REPS=3 run python tools/gpu_perf.py 100000 100000 1 semiglobal 4,8,16 0 0
  REPS=3 run python tools/gpu_perf.py 8087 9011 0 global 4,8 0 0
done
unset ANYSEQ_CELL_FORM
echo "#### traceback, K of the first levels 16 vs 32"
ANYSEQ_TRACE_LEVELS=1 ANYSEQ_K=32 run python tools/gpu_tb_time.py 1000000
ANYSEQ_K=16 run python tools/gpu_tb_time.py 1000000
} > gpurun_out/r2_sweep4.log 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputest4.log 2>&1
tail -5 gpurun_out/r2_gputest4.log
timeout 900 python bench.py --steps 2 --warmup 1 > gpurun_out/r2_bench_n1_a.json 2> gpurun_out/r2_bench_n1_a.err
tail -c 1500 gpurun_out/r2_bench_n1_a.json
timeout 900 python bench.py --workload reads --pairs 2000000 --steps 2 --warmup 1 --oracle-pairs 20000 > gpurun_out/r2_bench_reads_a.json 2> gpurun_out/r2_bench_reads_a.err
tail -c 1500 gpurun_out/r2_bench_reads_a.json; tail -3 gpurun_out/r2_bench_reads_a.err
