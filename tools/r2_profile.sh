#!/bin/bash
# round 2, profiling call (one GPU): ncu launch list of the bench command and one full-set capture of a full-size
# launch of the dominant kernel.  Each ncu run follows a plain run of the SAME command that exited 0.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
BENCH="python bench.py --steps 2 --warmup 1 --no-cpu"
$BENCH > gpurun_out/r2_prof_plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench.csv $BENCH > gpurun_out/r2_prof_ncu_bench.log 2>&1
echo "launch list rc=$?"
ONE="python tools/gpu_perf.py 0 0 1 semiglobal 0 0 0"
WL=1.0 REPS=1 $ONE > gpurun_out/r2_prof_plain_one.log 2>&1 &&
WL=1.0 REPS=1 ncu --set full --clock-control none --import-source on -k regex:strip_kernel -c 1 -o gpurun_out/r02_strip_kernel_fullsize $ONE > gpurun_out/r2_prof_ncu_one.log 2>&1
echo "full set rc=$?"
tail -3 gpurun_out/r2_prof_plain_one.log
ls -la gpurun_out/r02_*
