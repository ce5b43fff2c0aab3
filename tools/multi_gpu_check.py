#!/usr/bin/env python
"""N-GPU functional check (run under torchrun): the sharded batch path and the streamed strip wavefront give
exactly the single-GPU results.  torchrun --nproc-per-node N tools/multi_gpu_check.py"""
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import anyseq_b200 as A
from anyseq_b200 import workloads as W
from anyseq_b200.multigpu import StripWavefront, column_slices, score_batch_sharded

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
al = A.Aligner(local)
sch = A.affine_scoring_scheme()

# (1) batch of pairs: contiguous ranges per rank, gathered scores == the whole batch on one GPU
qd, qo, sd, so = W.read_batch(40001)
got, _ = score_batch_sharded(al, dist, rank, world, "semiglobal", qd, qo, sd, so, sch)
ref, _ = al.score_batch("semiglobal", qd, qo, sd, so, sch)
assert got.shape == ref.shape and (got == ref).all(), "sharded batch differs"

# (2) one long pair, streamed runs (no barrier between them), every run must give the single-GPU score
q, s, _ = W.whole_genome_pair(0.06)
m, n = len(q), len(s)
want = al.score("semiglobal", q, s, sch).score
c0, c1 = column_slices(n, world)[rank]
d_q = torch.from_numpy(q).cuda(); d_s = torch.from_numpy(np.ascontiguousarray(s[c0:c1])).cuda()
wave = StripWavefront(al, rank, world, m, dist, depth=2)
wave.reset()
parts = [wave.run("semiglobal", sch, d_q.data_ptr(), m, d_s.data_ptr(), c0, c1, n) for _ in range(5)]
for p in parts:
    assert wave.combine("semiglobal", sch, p).score == want, "streamed wavefront differs"
wave.reset()
p = wave.run("semiglobal", sch, d_q.data_ptr(), m, d_s.data_ptr(), c0, c1, n)
assert wave.combine("semiglobal", sch, p).score == want
wave.close()

# (3) several alignments side by side in one launch per rank (different pairs of one shape), streamed launches
q2, s2, _ = W.whole_genome_pair(0.06)
q2 = q2[::-1].copy(); s2 = s2[::-1].copy()
want2 = al.score("semiglobal", q2, s2, sch).score
d_q2 = torch.from_numpy(q2).cuda(); d_s2 = torch.from_numpy(np.ascontiguousarray(s2[c0:c1])).cuda()
wave = StripWavefront(al, rank, world, m, dist, depth=2, pairs_per_launch=4)
wave.reset()
for launch in range(3):
    qs = [d_q.data_ptr(), d_q2.data_ptr(), d_q.data_ptr(), d_q2.data_ptr()][: 4 - launch]
    ss = [d_s.data_ptr(), d_s2.data_ptr(), d_s.data_ptr(), d_s2.data_ptr()][: 4 - launch]
    parts = wave.run_multi("semiglobal", sch, qs, m, ss, c0, c1, n)
    got = [wave.combine("semiglobal", sch, p).score for p in parts]
    assert got == [want, want2, want, want2][: 4 - launch], (launch, got, want, want2)
wave.close()
dist.barrier()

# (4) end cells of the combined ranks == single GPU, host-buffer strip entry (anyseq_score_strip)
wave = StripWavefront(al, rank, world, m, dist, depth=1)
wave.reset()
for mode in ("global", "semiglobal", "local"):
    one = al.score(mode, q, s, sch)
    dist.barrier()
    p = wave.run_host(mode, sch, q, np.ascontiguousarray(s[c0:c1]), c0, c1, n)
    r = wave.combine(mode, sch, p)
    assert (r.score, r.end_i, r.end_j) == (one.score, one.end_i, one.end_j), (mode, (r.score, r.end_i, r.end_j), (one.score, one.end_i, one.end_j))
wave.close()

# (5) multi-GPU linear-space traceback (anyseq_align_sharded, world a power of two): bit-identical strings and splits,
#     linear and Gotoh gaps; timing at the size given by TB_N (default 200 kbp; 1000000 = BASELINE configs[2], checked
#     against the frozen CPU sha of tests/golden/fullsize.json)
import hashlib, json, time
from anyseq_b200.multigpu import ShardedTraceback
if world & (world - 1) == 0:
    tb_n = int(os.environ.get("TB_N", "200000"))
    tq, ts = W.random_pair(tb_n, tb_n, 1, 2)
    st = ShardedTraceback(al, rank, world, dist)
    gold = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "fullsize.json")))
    al.set_option("align_with_score", 0)
    for name, mode, tsch in (("linear", "local", A.linear_scoring_scheme()), ("affine", "local", A.affine_scoring_scheme()),
                             ("affine", "global", A.affine_scoring_scheme())):
        st.align(mode, tq[:20000], ts[:20000], tsch)                       # warm-up
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        piece = st.align(mode, tq, ts, tsch)
        dist.barrier()
        t_align = time.perf_counter() - t0
        aq, as_, splits = st.gather(tb_n, tb_n, piece)
        dist.barrier()
        t_multi = time.perf_counter() - t0
        sha = hashlib.sha256(aq + b"\n" + as_).hexdigest()[:16]
        if rank == 0:
            t0 = time.perf_counter()
            one = al.align(mode, tq, ts, tsch)
            t_one = time.perf_counter() - t0
            same = (aq, as_) == (one.aligned_query, one.aligned_subject) and splits == al.last_splits()
            g = gold.get("c3affine_local") if (name == "affine" and mode == "local" and tb_n == 1000000) else None
            print(f"sharded traceback {name} {mode} {tb_n} x {tb_n}: {world} GPUs {t_multi*1e3:.1f} ms (align {t_align*1e3:.1f} + gather), 1 GPU {t_one*1e3:.1f} ms "
                  f"(x{t_one/t_multi:.2f}), identical={same}, sha={sha}" + (f", frozen CPU sha equal={sha == g['sha']}" if g else ""), flush=True)
            assert same
            if g:
                assert sha == g["sha"]
        dist.barrier()
    al.set_option("align_with_score", 1)

if rank == 0:
    print(f"multi-GPU check ok on {world} GPUs: batch of {len(ref)} pairs, wavefront {m} x {n} score {want}")
dist.destroy_process_group()
