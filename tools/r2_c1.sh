#!/bin/bash
# small-problem diagnosis (C1): shape table with the shipped library and a (K, R) variant, per-strip timeline with the profile build
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/r2_c1_trace.jsonl
timeout 200 python tools/c1_probe.py > gpurun_out/r2_c1_table.log 2>&1
cat gpurun_out/r2_c1_table.log
ANYSEQ_LIB=$PWD/anyseq_b200/_build_r44/libanyseq_b200.so timeout 200 python tools/c1_probe.py > gpurun_out/r2_c1_table_r44.log 2>&1
grep device gpurun_out/r2_c1_table_r44.log
ANYSEQ_LIB=$PWD/anyseq_b200/_build_prof/libanyseq_b200.so ANYSEQ_TRACE_FILE=$PWD/gpurun_out/r2_c1_trace.jsonl timeout 200 python tools/c1_probe.py trace > gpurun_out/r2_c1_trace.log 2>&1
grep -E "trace K|phases|launch:|cycles/batch" gpurun_out/r2_c1_trace.log | tail -12
