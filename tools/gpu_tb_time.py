#!/usr/bin/env python
"""wall time of the linear-space traceback (all Hirschberg levels) and of small score calls. Usage: gpu_tb_time.py [N=1000000]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import anyseq_b200 as A
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
rng = np.random.default_rng(1)
ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
q = ACGT[rng.integers(0, 4, n)]; s = ACGT[rng.integers(0, 4, n)]
al = A.Aligner()
for mode, sch in (("local", A.linear_scoring_scheme()), ("global", A.linear_scoring_scheme()), ("global", A.affine_scoring_scheme())):
    best = 1e9
    for rep in range(2):
        t = time.perf_counter(); r = al.align(mode, q, s, sch); best = min(best, time.perf_counter() - t)
    print(f"traceback {n} x {n} {mode} affine={sch.affine}: {best*1e3:.1f} ms wall, {n*float(n)/best/1e9:.1f} GCUPS (m*n numerator), score {r.score}", flush=True)
for m in (9011, 20000, 35000):
    qq, ss = q[:m], s[:m]
    best = 1e9
    for rep in range(5):
        r = al.score("semiglobal", qq, ss, A.affine_scoring_scheme()); best = min(best, r.kernel_ms)
    print(f"score {m} x {m} semiglobal affine: {best:.3f} ms {m*float(m)/best/1e6:.1f} GCUPS", flush=True)
