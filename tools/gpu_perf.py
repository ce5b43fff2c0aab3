#!/usr/bin/env python
"""Timing sweep of the strip kernel (device-resident inputs). Usage:
   gpu_perf.py M N [affine=1] [mode=semiglobal] [Ks=16,32] [bands=0] [bps=0]"""
import os
import sys
import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import anyseq_b200 as A  # noqa: E402


def main():
    m = int(sys.argv[1]); n = int(sys.argv[2])
    affine = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    mode = sys.argv[4] if len(sys.argv) > 4 else "semiglobal"
    Ks = [int(x) for x in (sys.argv[5] if len(sys.argv) > 5 else "16,32").split(",")]
    bands = [int(x) for x in (sys.argv[6] if len(sys.argv) > 6 else "0").split(",")]
    bpss = [int(x) for x in (sys.argv[7] if len(sys.argv) > 7 else "0").split(",")]
    reps = int(os.environ.get("REPS", "2"))
    if os.environ.get("WL"):
        from anyseq_b200 import workloads as W
        hq, hs, _ = W.whole_genome_pair(float(os.environ["WL"]))
        q = torch.from_numpy(hq).cuda(); s = torch.from_numpy(hs).cuda(); m, n = len(hq), len(hs)
    else:
        g = torch.Generator(device="cuda"); g.manual_seed(1)
        lut = torch.tensor([65, 67, 71, 84], dtype=torch.uint8, device="cuda")
        q = lut[torch.randint(0, 4, (m,), device="cuda", generator=g)]
        s = lut[torch.randint(0, 4, (n,), device="cuda", generator=g)]
    torch.cuda.synchronize()
    al = A.Aligner()
    if os.environ.get("SLACK"):
        al.set_option("band_slack", int(os.environ["SLACK"]))
    sch = A.affine_scoring_scheme() if affine else A.linear_scoring_scheme()
    for K in Ks:
        for band in bands:
            for bps in bpss:
                al.tune(cols_per_lane=K, band_rows=band, blocks_per_sm=bps, watchdog_ms=20000)
                best = 1e30
                for rep in range(reps):
                    r = al.score_device(mode, q.data_ptr(), m, s.data_ptr(), n, sch)
                    best = min(best, r.kernel_ms)
                print(f"m={m} n={n} affine={affine} {mode} K={K} band={band} bps={bps}: {best:.2f} ms "
                      f"{m*n/best/1e6:.1f} GCUPS score={r.score}", flush=True)


if __name__ == "__main__":
    main()
