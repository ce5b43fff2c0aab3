#!/usr/bin/env python
"""quick traceback parity check against the oracle (development helper)"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import anyseq_b200 as A
from oracle import oracle as O
rng = np.random.default_rng(3)
ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
al = A.Aligner(); al.tune(watchdog_ms=5000)
bad = 0
for (m, n) in [(861, 914), (10, 8), (300, 70), (100, 129), (200, 300), (1300, 2300), (5000, 9000), (2500, 16385), (20000, 30000)]:
    if (m, n) == (861, 914):
        q, s = O.reference_random_pair(256, 1024)
    else:
        q = ACGT[rng.integers(0, 4, m)]; s = ACGT[rng.integers(0, 4, n)]
        k = min(m, n); s[:k] = np.where(rng.random(k) < 0.9, q[:k], s[:k])
    for mode in ("global", "semiglobal", "local"):
        t = time.time(); r = al.align(mode, q, s); dt = time.time() - t
        ret, aq, as_, sp = O.traceback_lintime(mode, q, s, threads=8)
        ok_s = al.last_splits() == sp.tolist()
        ok_a = (r.aligned_query, r.aligned_subject) == (aq, as_)
        if not (ok_s and ok_a):
            bad += 1
            print("MISMATCH", m, n, mode, "splits", ok_s, "strings", ok_a, al.last_splits()[:12], sp.tolist()[:12])
        else:
            print("ok", m, n, mode, f"{dt*1e3:.1f} ms", "score", r.score)
print("BAD", bad)
sys.exit(1 if bad else 0)
