#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputest7.log 2>&1
tail -4 gpurun_out/r2_gputest7.log
bash tools/r2_profile.sh
