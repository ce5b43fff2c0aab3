#!/bin/bash
# last validation of the round: parity suite + smoke on the shipped library
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 120 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputest24.log 2>&1
tail -3 gpurun_out/r2_gputest24.log
timeout 25 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
