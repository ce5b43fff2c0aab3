#!/usr/bin/env python
"""First-contact GPU check: integer-pipe microbenchmarks, a parity sweep of the
strip kernel against the CPU oracle (all modes x linear/affine x K x bands), and
a few timings.  Writes gpurun_out/gpu_check.log style output to stdout."""
import itertools
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import anyseq_b200 as A  # noqa: E402
from oracle import oracle as O  # noqa: E402

rng = np.random.default_rng(1234)
ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def rand_seq(n, alphabet=ACGT):
    return alphabet[rng.integers(0, len(alphabet), size=n)]


def related(q, n, sub=0.05, indel=0.02):
    """a mutated copy of q cut/extended to n symbols"""
    out = []
    for c in q:
        r = rng.random()
        if r < indel / 2:
            continue
        if r < indel:
            out.append(ACGT[rng.integers(0, 4)])
        out.append(ACGT[rng.integers(0, 4)] if rng.random() < sub else c)
    out = np.array(out[:n], dtype=np.uint8)
    if len(out) < n:
        out = np.concatenate([out, rand_seq(n - len(out))])
    return out


def main():
    al = A.Aligner()
    print("device:", al.device_info(), flush=True)
    al.tune(watchdog_ms=4000)

    if "--no-micro" not in sys.argv:
        for kind, name in enumerate(["dpx_only", "linear_mix5", "affine_mix7", "imad_only", "dpx+imad"]):
            ops, mhz = al.measure_int_peak(kind)
            print(f"int_peak kind={kind} {name}: {ops/1e12:.3f} Tlaneop/s  clk~{mhz:.0f} MHz  "
                  f"-> per SM per clk {ops/(mhz*1e6)/148:.1f}", flush=True)

    nfail = 0
    ncase = 0
    t0 = time.time()
    schemes = [A.linear_scoring_scheme(2, -1, -1), A.affine_scoring_scheme(2, -1, -2, -1),
               A.linear_scoring_scheme(3, -2, -4), A.affine_scoring_scheme(5, -4, -10, -1)]
    shapes = [(1, 1), (1, 7), (7, 1), (2, 129), (31, 33), (33, 31), (64, 128), (100, 127), (100, 128), (100, 129),
              (257, 255), (300, 1024), (300, 1025), (1000, 513), (700, 2048), (129, 4096), (1500, 3000),
              (4097, 1000), (2500, 2500)]
    for (m, n) in shapes:
        q = rand_seq(m)
        s = related(q, n) if (m > 50 and n > 50) else rand_seq(n)
        for K, band in [(4, 0), (8, 64), (16, 32), (32, 0), (4, 96), (32, 160)]:
            al.tune(cols_per_lane=K, band_rows=band, watchdog_ms=4000)
            for mode in ("global", "semiglobal", "local"):
                for sch in schemes:
                    ncase += 1
                    try:
                        r = al.score(mode, q, s, sch)
                    except A.AnyseqError as e:
                        print("ERROR", m, n, K, band, mode, sch, e, flush=True)
                        nfail += 1
                        if nfail > 20:
                            print("too many failures"); return 1
                        continue
                    if sch.affine:
                        ref = O.textbook_affine(mode, q, s, sch.same, sch.diff, sch.gap_init, sch.gap_extend)
                    else:
                        ref = O.textbook_linear(mode, q, s, sch.same, sch.diff, sch.gap_extend)
                    if r.score != ref:
                        nfail += 1
                        print(f"MISMATCH m={m} n={n} K={K} band={band} {mode} {sch}: gpu={r.score} ref={ref}", flush=True)
                        if nfail > 20:
                            print("too many failures"); return 1
    print(f"parity sweep: {ncase} cases, {nfail} failures, {time.time()-t0:.1f}s", flush=True)

    # end positions (semiglobal/global) vs the restated reference
    for (m, n) in [(300, 1025), (1500, 3000)]:
        q = rand_seq(m); s = related(q, n)
        al.tune(cols_per_lane=8, band_rows=128, watchdog_ms=4000)
        for mode in ("global", "semiglobal"):
            r = al.score(mode, q, s)
            ref = O.score_linear(mode, q, s)
            ok = (r.score, r.end_i, r.end_j) == ref
            print("pos", mode, m, n, (r.score, r.end_i, r.end_j), ref, "OK" if ok else "MISMATCH", flush=True)
            nfail += 0 if ok else 1

    # reference-RNG inputs
    q, s = O.reference_random_pair(10000, 1024)
    al.tune(0, 0, 0, 4000)
    got = [al.score(mo, q, s).score for mo in ("global", "semiglobal", "local")]
    print("align -r 10000 scores:", got, "expected [6317, 6334, 6335]", flush=True)
    nfail += 0 if got == [6317, 6334, 6335] else 1

    # timings
    if "--no-time" not in sys.argv:
        for (m, n) in [(8087, 9011), (100_000, 100_000), (400_000, 400_000)]:
            q = rand_seq(m); s = related(q[: min(m, 200000)], n) if n <= 200000 else rand_seq(n)
            for sch in (A.linear_scoring_scheme(), A.affine_scoring_scheme()):
                for K in (8, 16, 32):
                    al.tune(cols_per_lane=K, band_rows=0, watchdog_ms=8000)
                    best = None
                    for rep in range(2):
                        r = al.score("semiglobal", q, s, sch)
                        best = r.kernel_ms if best is None else min(best, r.kernel_ms)
                    print(f"time m={m} n={n} affine={sch.affine} K={K}: {best:.3f} ms  "
                          f"{m*n/best/1e6:.1f} GCUPS score={r.score}", flush=True)
    print("TOTAL FAILURES", nfail, flush=True)
    return 1 if nfail else 0


if __name__ == "__main__":
    sys.exit(main())
