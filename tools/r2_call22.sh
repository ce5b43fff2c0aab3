#!/bin/bash
# validation of the small-problem strip-width fit (engine.cu: pick_K): parity suite + the shape table, linear and Gotoh
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 240 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputest23.log 2>&1
tail -3 gpurun_out/r2_gputest23.log
timeout 120 python tools/c1_probe.py > gpurun_out/r2_c1_table_fit2.log 2>&1
cat gpurun_out/r2_c1_table_fit2.log
