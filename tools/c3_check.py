#!/usr/bin/env python
"""BASELINE.json configs[2] at full size: local, linear gaps (2,-1,-1), linear-space traceback of a random
1 Mbp x 1 Mbp pair; GPU result vs the CPU oracle, bit for bit (SURVEY.md 8d "C3").  Writes a JSON summary."""
import hashlib, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import anyseq_b200 as A
from anyseq_b200 import workloads as W
from oracle import oracle as O

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
q, s = W.random_pair(n, n, 1, 2)
al = A.Aligner()
out = {"rows": n, "cols": n, "seeds": [1, 2]}
for mode in (["local"] if len(sys.argv) < 3 else sys.argv[2].split(",")):
    al.set_option("align_with_score", 0)
    al.align(mode, q[:2000], s[:2000])                      # warm-up
    t0 = time.perf_counter(); r = al.align(mode, q, s); t_gpu = time.perf_counter() - t0
    al.set_option("align_with_score", 1)
    sc = al.score(mode, q, s)
    t0 = time.perf_counter(); ret, aq, as_, sp = O.traceback_lintime(mode, q, s, threads=os.cpu_count()); t_cpu = time.perf_counter() - t0
    same = (r.aligned_query == aq and r.aligned_subject == as_ and al.last_splits() == sp.tolist())
    out[mode] = {"bit_exact": bool(same), "sha": hashlib.sha256(aq + b"\n" + as_).hexdigest()[:16],
                 "gpu_s": t_gpu, "gpu_kernel_ms": r.kernel_ms, "gpu_gcups": n * n / t_gpu / 1e9,
                 "oracle_s": t_cpu, "oracle_gcups": n * n / t_cpu / 1e9, "oracle_threads": os.cpu_count(),
                 "optimal_score": int(sc.score), "cigar_len": len(A.cigar(r.aligned_query, r.aligned_subject))}
    print(mode, out[mode], flush=True)
print(json.dumps(out))
