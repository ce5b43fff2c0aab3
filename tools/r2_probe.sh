#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29577"
timeout 300 $TR tools/chain_probe.py > gpurun_out/r2_chain_probe_n$N.log 2>&1
grep "rank" gpurun_out/r2_chain_probe_n$N.log
ANYSEQ_K=32 timeout 300 $TR tools/chain_probe.py > gpurun_out/r2_chain_probe_k32_n$N.log 2>&1
grep "rank" gpurun_out/r2_chain_probe_k32_n$N.log
