#!/bin/bash
# validation of the ramp-step change + K = 8 for small problems: parity suite, small-problem table, regression checks, bench line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputest21.log 2>&1
tail -3 gpurun_out/r2_gputest21.log
timeout 200 python tools/c1_probe.py > gpurun_out/r2_c1_table_after.log 2>&1
cat gpurun_out/r2_c1_table_after.log
rm -f gpurun_out/r2_c1_trace_after.jsonl
ANYSEQ_LIB=$PWD/anyseq_b200/_build_prof/libanyseq_b200.so ANYSEQ_TRACE_FILE=$PWD/gpurun_out/r2_c1_trace_after.jsonl timeout 200 python tools/c1_probe.py trace > gpurun_out/r2_c1_trace_after.log 2>&1
grep -E "trace K|phases|cycles/batch" gpurun_out/r2_c1_trace_after.log | tail -6
(
REPS=2 timeout 200 python tools/gpu_perf.py 100000 100000 1 semiglobal 4,8,16 0 0
REPS=2 timeout 200 python tools/gpu_perf.py 1000000 1000000 1 semiglobal 0 0 0
REPS=2 timeout 200 python tools/gpu_perf.py 4641652 575488 1 semiglobal 0 0 0
REPS=3 timeout 200 python tools/gpu_perf.py 40000 575488 1 semiglobal 0 0 0
REPS=3 timeout 200 python tools/gpu_perf.py 80000 575488 1 semiglobal 0 0 0
WL=1.0 REPS=2 timeout 200 python tools/gpu_perf.py 0 0 1 semiglobal 0 0 0
) 2>&1 | grep GCUPS | tee gpurun_out/r2_perf21.log
timeout 400 python bench.py > gpurun_out/r2_bench_n1_final3.json 2> gpurun_out/r2_bench_n1_final3.err
tail -c 1500 gpurun_out/r2_bench_n1_final3.json
