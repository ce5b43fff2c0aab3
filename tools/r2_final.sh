#!/bin/bash
# round 2, final single-GPU call: whole parity suite, the bench line, cell form 2 vs 1 on this box, ncu of the final default
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputest_final.log 2>&1
tail -3 gpurun_out/r2_gputest_final.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench_n1_final.json 2> gpurun_out/r2_bench_n1_final.err
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_n1_final.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['alu_pipe_frac'], d['roofline']['issue_slot_frac'], d['golden_fullsize']['equal'], d['parity_check']['equal'], d['clocks'])"
for form in 2 1; do ANYSEQ_CELL_FORM=$form WL=1.0 REPS=2 timeout 300 python tools/gpu_perf.py 0 0 1 semiglobal 32 0 3 2>&1 | grep GCUPS; done
timeout 300 python tools/gpu_tb_time.py 1000000 2>&1 | grep -v "^$"
BENCH="python bench.py --steps 2 --warmup 1 --no-cpu"
$BENCH > gpurun_out/r2_prof_plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench_final.csv $BENCH > gpurun_out/r2_prof_ncu_bench.log 2>&1
echo "launch list rc=$?"
ONE="python tools/gpu_perf.py 0 0 1 semiglobal 0 0 0"
WL=1.0 REPS=1 $ONE > gpurun_out/r2_prof_plain_one.log 2>&1 &&
WL=1.0 REPS=1 ncu --set full --clock-control none --import-source on -k regex:strip_kernel -c 1 -o gpurun_out/r02_strip_kernel_fullsize_final $ONE > gpurun_out/r2_prof_ncu_one.log 2>&1
echo "full set rc=$?"
tail -2 gpurun_out/r2_prof_plain_one.log
