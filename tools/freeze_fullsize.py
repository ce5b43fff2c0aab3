#!/usr/bin/env python
"""One-off CPU runs of BASELINE.json's full-size configurations, frozen as fixtures in tests/golden/fullsize.json
so that the GPU tests can assert the headline workload's result against a CPU-computed value.

  c2        configs[1]/[4]: the 4 641 652 x 4 600 000 synthetic ecoli-like pair, semiglobal Gotoh (2,-1,-2,-1), score + end
            cell, by oracle/fullsize_check.c (the vectorised second CPU implementation, validated against the scalar
            restatement by tests/test_oracle.py::test_fullsize_check_equals_oracle; the scalar restatement itself needs
            about 4 h on this container's 8 cores).
  c3affine  configs[2] as written: LOCAL, AFFINE, linear-space traceback of the random 1 Mbp pair (seeds 1/2) by the scalar
            restatement oracle_traceback_lintime_affine; the alignment strings are frozen as a sha256.
  c3linear  same with linear gaps (the reference's own scheme), local + global.

Usage: python tools/freeze_fullsize.py c2|c3affine|c3linear [threads] [scale]
(CPU only; takes 0.5-1.5 h per entry on 6-8 cores.)  Results are merged into tests/golden/fullsize.json.
"""
import hashlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from anyseq_b200 import workloads as W  # noqa: E402
from oracle import oracle as O  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "fullsize.json")


def merge(key, val):
    data = {}
    if os.path.exists(OUT):
        data = json.load(open(OUT))
    data[key] = val
    with open(OUT, "w") as f:
        json.dump(data, f, indent=1, sort_keys=True)
    print(key, json.dumps(val), flush=True)


def sha(aq, as_):
    return hashlib.sha256(aq + b"\n" + as_).hexdigest()[:16]


def main():
    what = sys.argv[1]
    threads = int(sys.argv[2]) if len(sys.argv) > 2 else os.cpu_count()
    scale = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
    if what == "c2":
        q, s, desc = W.whole_genome_pair(scale)
        t0 = time.time()
        sc, pi, pj = O.fullsize_score("semiglobal", q, s, 2, -1, -2, -1, threads=threads)
        dt = time.time() - t0
        key = "c2_semiglobal_affine" if scale == 1.0 else f"c2_semiglobal_affine_scale_{scale}"
        merge(key, {"workload": desc, "m": len(q), "n": len(s), "scheme": [2, -1, -2, -1], "mode": "semiglobal",
                    "score": sc, "end_i": pi, "end_j": pj, "fnv_q": f"{O.fnv1a64(q):016x}", "fnv_s": f"{O.fnv1a64(s):016x}",
                    "by": "oracle/fullsize_check.c", "cpu_seconds": round(dt, 1), "threads": threads,
                    "gcups": round(len(q) * len(s) / dt / 1e9, 2)})
    elif what in ("c3affine", "c3linear"):
        n = int(1_000_000 * scale)
        q, s = W.random_pair(n, n, 1, 2)
        modes = ["local"] if what == "c3affine" else ["local", "global"]
        for mode in modes:
            t0 = time.time()
            if what == "c3affine":
                ret, aq, as_, sp, ty = O.traceback_lintime_affine(mode, q, s, 2, -1, -2, -1, threads=threads)
                extra = {"types_sha": hashlib.sha256(ty.tobytes()).hexdigest()[:16], "scheme": [2, -1, -2, -1]}
                by = "oracle_traceback_lintime_affine"
            else:
                ret, aq, as_, sp = O.traceback_lintime(mode, q, s, 2, -1, -1, threads=threads)
                extra = {"scheme": [2, -1, -1]}
                by = "oracle_traceback_lintime"
            dt = time.time() - t0
            key = f"{what}_{mode}" if scale == 1.0 else f"{what}_{mode}_scale_{scale}"
            merge(key, dict({"m": n, "n": n, "seeds": [1, 2], "mode": mode, "ret": int(ret), "sha": sha(aq, as_),
                             "splits_sha": hashlib.sha256(np.asarray(sp, dtype=np.int32).tobytes()).hexdigest()[:16],
                             "fnv_q": f"{O.fnv1a64(q):016x}", "fnv_s": f"{O.fnv1a64(s):016x}", "by": by,
                             "cpu_seconds": round(dt, 1), "threads": threads}, **extra))
    else:
        raise SystemExit(__doc__)


if __name__ == "__main__":
    main()
