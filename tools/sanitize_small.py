#!/usr/bin/env python
"""small workload touching every kernel family, meant to run under compute-sanitizer --tool memcheck"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import anyseq_b200 as A
from anyseq_b200 import workloads as W
rng = np.random.default_rng(3)
ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
al = A.Aligner()
al.tune(watchdog_ms=60000)
allb = np.arange(256, dtype=np.uint8)
for (m, n) in [(1, 1), (100, 129), (700, 2100), (2000, 1300)]:
    q = ACGT[rng.integers(0, 4, m)]; s = ACGT[rng.integers(0, 4, n)]
    qb = allb[rng.integers(0, 256, m)]; sb = allb[rng.integers(0, 256, n)]
    for K in (4, 8, 16, 32):
        al.tune(cols_per_lane=K, band_rows=96, watchdog_ms=60000)
        for mode in ("global", "semiglobal", "local"):
            for sch in (A.linear_scoring_scheme(), A.affine_scoring_scheme()):
                al.score(mode, q, s, sch)
                if K <= 16:
                    al.score(mode, qb, sb, sch)
    al.tune(0, 0, 0, 60000)
    for mode in ("global", "semiglobal", "local"):
        al.align(mode, q, s)
        al.align(mode, q, s, A.affine_scoring_scheme())
qd, qo, sd, so = W.read_batch(300)
for mode in ("global", "semiglobal", "local"):
    al.score_batch(mode, qd, qo, sd, so, A.affine_scoring_scheme())
    al.score_batch(mode, qd, qo, sd, so, A.linear_scoring_scheme())
al.measure_int_peak(0)
print("sanitize workload done")
