#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29577"
echo "equal slices"; timeout 200 $TR tools/chain_probe.py > gpurun_out/r2_chain_probe_n$N.log 2>&1
grep "rank" gpurun_out/r2_chain_probe_n$N.log
echo "bench slices"; SLICES=bench timeout 200 $TR tools/chain_probe.py > gpurun_out/r2_chain_probe_bench_n$N.log 2>&1
grep "rank" gpurun_out/r2_chain_probe_bench_n$N.log
