#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "batch" > gpurun_out/r2_gputest14.log 2>&1
tail -3 gpurun_out/r2_gputest14.log
for quad in 1 0; do
  echo "batch_quad=$quad"
  ANYSEQ_BATCH_QUAD=$quad timeout 600 python bench.py --workload reads --pairs 2000000 --steps 3 --warmup 2 --oracle-pairs 20000 2> gpurun_out/r2_reads_q$quad.err | python -c "
import json,sys; r=json.loads(sys.stdin.read()); print({k:(round(v['gcups']),round(v['e2e_gcups']),v['checksum']) for k,v in r['per_scheme'].items()}, r['oracle_check'])"
done
