#!/usr/bin/env python
"""Generates tests/golden/golden.json.

The reference ships no golden vectors and cannot be executed here (its compute
code is Impala; the AnyDSL toolchain is absent), so these fixtures are produced
by the committed CPU oracle (oracle/anyseq_oracle.c), AFTER that oracle has been
pinned against
  * the independent textbook DPs (scores), and
  * the restatement-derived values of SURVEY.md Appendix C (scores, split rows,
    traceback hashes for the reference-RNG input `align -r`).
They freeze the oracle's behaviour so that both the oracle (tests -m "not gpu")
and the CUDA path (tests -m gpu) are checked against fixed bytes.

    python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def sha(aq, as_):
    return hashlib.sha256(aq + b"\n" + as_).hexdigest()[:16]


def mutate(rng, q, n):
    out = []
    for c in q:
        r = rng.random()
        if r < 0.01:
            continue
        if r < 0.02:
            out.append(int(ACGT[rng.integers(0, 4)]))
        out.append(int(ACGT[rng.integers(0, 4)]) if rng.random() < 0.05 else int(c))
    out = np.array(out[:n], dtype=np.uint8)
    if len(out) < n:
        out = np.concatenate([out, ACGT[rng.integers(0, 4, n - len(out))]])
    return out


def main():
    rng = np.random.default_rng(20261018)
    cases = []
    inputs = []
    q, s = O.reference_random_pair(256, 1024)
    inputs.append(("align -r (reference RNG, 861 x 914)", q, s))
    for (m, n) in [(1, 1), (5, 3), (17, 64), (40, 65), (64, 128), (100, 129), (200, 300), (333, 1024),
                   (700, 1500), (1300, 2300), (2100, 260), (97, 4100)]:
        qq = ACGT[rng.integers(0, 4, m)]
        ss = mutate(rng, qq, n) if min(m, n) > 50 else ACGT[rng.integers(0, 4, n)]
        inputs.append((f"random {m} x {n}", qq, ss))
    # non-ACGT bytes: symbols are compared raw (quirk Q8)
    inputs.append(("bytes", np.frombuffer(b"acgtNNNN\rACGTacgtnnACGT" * 9, dtype=np.uint8),
                   np.frombuffer(b"ACGTNNacgt\rACGGTacgtnnACGT" * 8, dtype=np.uint8)))
    for name, qq, ss in inputs:
        c = {"name": name, "q": bytes(qq).decode("latin-1"), "s": bytes(ss).decode("latin-1"), "linear": {}, "affine": {},
             "traceback": {}, "traceback_full": {}}
        for mode in ("global", "semiglobal", "local"):
            for key, (sa, di, ga) in {"2,-1,-1": (2, -1, -1), "3,-2,-4": (3, -2, -4)}.items():
                sc = O.score_linear(mode, qq, ss, sa, di, ga)
                assert sc[0] == O.textbook_linear(mode, qq, ss, sa, di, ga)
                c["linear"].setdefault(mode, {})[key] = list(sc)
            for key, (sa, di, gi, ge) in {"2,-1,-2,-1": (2, -1, -2, -1), "5,-4,-10,-1": (5, -4, -10, -1)}.items():
                sc = O.score_affine(mode, qq, ss, sa, di, gi, ge)
                assert sc[0] == O.textbook_affine(mode, qq, ss, sa, di, gi, ge)
                c["affine"].setdefault(mode, {})[key] = sc[0]
            ret, aq, as_, sp = O.traceback_lintime(mode, qq, ss)
            c["traceback"][mode] = {"ret": ret, "sha": sha(aq, as_), "splits": sp.tolist(),
                                    "column_score": O.column_score(aq, as_),
                                    "aq": aq.decode("latin-1") if len(aq) <= 2000 else None,
                                    "as": as_.decode("latin-1") if len(as_) <= 2000 else None}
            # Gotoh linear-space traceback (build-defined, parity unpinned vs the reference): frozen all the same
            ret, aq, as_, sp, ty = O.traceback_lintime_affine(mode, qq, ss, 2, -1, -2, -1)
            c.setdefault("traceback_affine", {})[mode] = {"ret": ret, "sha": sha(aq, as_), "splits": sp.tolist(),
                                                          "types": ty.tolist(),
                                                          "column_score": O.column_score_affine(aq, as_, 2, -1, -2, -1)}
            # traceback_full (src/align.impala:190-216); "affine" = build-defined Gotoh variant (2,-1,-2,-1)
            sc, aq, as_, st = O.traceback_full(mode, qq, ss)
            assert sc == O.score_linear(mode, qq, ss)[0]
            sca, aqa, asa, sta = O.traceback_full(mode, qq, ss, 2, -1, -2, -1)
            assert sca == O.score_affine(mode, qq, ss)[0]
            c["traceback_full"][mode] = {"score": sc, "sha": sha(aq, as_), "start": list(st),
                                         "column_score": O.column_score(aq, as_),
                                         "affine": {"score": sca, "sha": sha(aqa, asa), "start": list(sta),
                                                    "column_score": O.column_score_affine(aqa, asa)}}
        cases.append(c)
    out = {"generator": "tests/golden/make_golden.py", "appendix_c": {
        "align -r": {"m": 861, "n": 914, "scores": [654, 659, 659], "fnv_q": "4de67cd69011a8a7", "fnv_s": "08f2452b0e6358ff",
                     "splits_global": [0, 108, 248, 391, 518, 610, 722, 846, 861],
                     "sha": {"global": "5e478e07b593867a", "semiglobal": "dd7799193cb7d12c", "local": "dd7799193cb7d12c"},
                     "nonblank": {"global": 1053, "semiglobal": 1037, "local": 1037},
                     "column_score": {"global": 654, "semiglobal": 661, "local": 661},
                     "legacy_return": [-861, 0, -2147483647]},
        "align -r 10000": {"m": 8087, "n": 9011, "scores": [6317, 6334, 6335], "fnv_q": "8b55329edaf9f77e", "fnv_s": "72a453f2b40c4d37"},
        "align -r 10000 10000": {"m": 10000, "n": 10000, "scores": [7499, 7502, 7502], "fnv_q": "38641031bf865586", "fnv_s": "99ac2addc143b3f7"},
    }, "cases": cases}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=0)
    print("wrote", path, os.path.getsize(path), "bytes,", len(cases), "cases")


if __name__ == "__main__":
    main()
