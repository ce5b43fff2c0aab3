#!/usr/bin/env python
"""Where the time of a small alignment goes (BASELINE configs[0]: 8087 x 9011, linear gaps, global).

   c1_probe.py          timing table over matrix shapes and strip widths: T = c0 + a * rows + b * columns
   ANYSEQ_LIB=<a -DANYSEQ_PROFILE build> ANYSEQ_TRACE_FILE=out.jsonl c1_probe.py trace
                        per-strip timeline (start, first border batch, end) of the C1 shape
"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import anyseq_b200 as A  # noqa: E402


def seqs(m, n, seed=1):
    rng = np.random.default_rng(seed)
    lut = np.frombuffer(b"ACGT", dtype=np.uint8)
    return lut[rng.integers(0, 4, m)], lut[rng.integers(0, 4, n)]


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "table"
    al = A.Aligner()
    lin = A.linear_scoring_scheme()
    aff = A.affine_scoring_scheme()
    if what == "trace":
        hq, hs = seqs(8087, 9011)
        q = torch.from_numpy(hq).cuda(); s = torch.from_numpy(hs).cuda()
        for K in (4, 8):
            al.tune(cols_per_lane=K, watchdog_ms=20000)
            for _ in range(3):
                r = al.score_device("global", q.data_ptr(), len(hq), s.data_ptr(), len(hs), lin)
            print(f"trace K={K}: {r.kernel_ms * 1e3:.0f} us score={r.score}", flush=True)
        return
    shapes = [(64, 128), (8087, 128), (64, 9011), (8087, 9011), (16174, 9011), (8087, 18022), (32348, 9011), (2048, 2048), (1024, 9011)]
    for (sch, mode, name) in [(lin, "global", "linear global"), (aff, "semiglobal", "Gotoh semiglobal")]:
        for (m, n) in shapes:
            hq, hs = seqs(m, n)
            q = torch.from_numpy(hq).cuda(); s = torch.from_numpy(hs).cuda()
            torch.cuda.synchronize()
            row = []
            for K in (0, 4, 8, 16):         # 0 = the engine's own choice
                al.tune(cols_per_lane=K, watchdog_ms=20000)
                best = 1e30
                for _ in range(5):
                    r = al.score_device(mode, q.data_ptr(), m, s.data_ptr(), n, sch)
                    best = min(best, r.kernel_ms)
                row.append(f"K={K if K else 'auto'}: {best * 1e3:7.1f} us")
            print(f"device {m:6d} x {n:6d} {name}   " + "  ".join(row) + f"  score={r.score}", flush=True)
    # the call `align -r 10000` makes: host buffers in, score out (wall clock, median and best of 30)
    al.tune(cols_per_lane=0, watchdog_ms=20000)
    for (m, n, sch, mode, name) in [(8087, 9011, lin, "global", "linear global"), (8087, 9011, aff, "semiglobal", "Gotoh semiglobal"),
                                    (1024, 1024, lin, "global", "linear global"), (64, 128, lin, "global", "linear global")]:
        hq, hs = seqs(m, n)
        for _ in range(3):
            al.score(mode, hq, hs, sch)
        ts, ks = [], []
        for _ in range(30):
            t0 = time.perf_counter()
            r = al.score(mode, hq, hs, sch)
            ts.append(time.perf_counter() - t0)
            ks.append(r.kernel_ms)
        ts.sort(); ks.sort()
        print(f"host   {m:6d} x {n:6d} {name}: wall median {ts[15] * 1e6:.0f} us, best {ts[0] * 1e6:.0f} us; device span median {ks[15] * 1e3:.0f} us, "
              f"best {ks[0] * 1e3:.0f} us  ({m * n / ts[15] / 1e9:.1f} GCUPS wall)  score={r.score}", flush=True)


if __name__ == "__main__":
    main()
