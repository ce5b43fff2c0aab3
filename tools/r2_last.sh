#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputest_last.log 2>&1
tail -3 gpurun_out/r2_gputest_last.log
timeout 900 python bench.py > gpurun_out/r2_bench_n1_last.json 2> gpurun_out/r2_bench_n1_last.err
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_n1_last.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['alu_pipe_frac'], d['roofline']['issue_slot_frac'], d['golden_fullsize']['equal'], d['parity_check']['equal'], d['gpu_launches'], d['clocks'])"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()"
WL=1.0 REPS=1 timeout 300 python tools/gpu_perf.py 0 0 1 local 0 0 0 | grep GCUPS
