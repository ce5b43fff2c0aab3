#!/usr/bin/env python
"""Where does a chained multi-GPU slice lose time against the same slice alone?  torchrun --nproc-per-node N:
every rank relaxes an equal 575 488-column slice of a (4 641 652 x N*575 488) pair (a) alone, as its own matrix,
(b) chained to its neighbours; prints the device time of every rank for both."""
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import anyseq_b200 as A
from anyseq_b200 import workloads as W
from anyseq_b200.multigpu import StripWavefront

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
al = A.Aligner(local)
sch = A.affine_scoring_scheme()
m = int(os.environ.get("ROWS", "4641652"))
if os.environ.get("SLICES") == "bench":        # the slices bench.py uses for the 4.6 Mbp pair
    from anyseq_b200.multigpu import column_slices
    widths = [b - a for a, b in column_slices(4_600_000, world, rows=m)]
elif os.environ.get("SLICES"):
    widths = [int(x) for x in os.environ["SLICES"].split(",")]
else:
    widths = [int(os.environ.get("SLICE", "575488"))] * world
wcols = widths[rank]
n = sum(widths)
q = W.random_dna(m, 42)
s = W.mutated_copy(q, n, 43) if n <= m else np.concatenate([W.mutated_copy(q, m, 43), W.random_dna(n - m, 44)])
c0 = sum(widths[:rank]); c1 = c0 + wcols
d_q = torch.from_numpy(q).cuda()
d_s = torch.from_numpy(np.ascontiguousarray(s[c0:c1])).cuda()
res = {}
for rep in range(2):
    r = al.score_device("semiglobal", d_q.data_ptr(), m, d_s.data_ptr(), wcols, sch)
    res["alone"] = r.kernel_ms
wave = StripWavefront(al, rank, world, m, dist, depth=1)
wave.reset()
for rep in range(3):
    dist.barrier(); torch.cuda.synchronize()
    p = wave.run("semiglobal", sch, d_q.data_ptr(), m, d_s.data_ptr(), c0, c1, n)
    res["chained"] = p.kernel_ms
out = [None] * world
dist.all_gather_object(out, (rank, res))
if rank == 0:
    for r_, v in out:
        print(f"rank {r_}: {widths[r_]} columns, alone {v['alone']:.1f} ms, chained {v['chained']:.1f} ms", flush=True)
wave.close()
dist.destroy_process_group()
