#!/usr/bin/env python
"""throughput of P same-shape slices relaxed side by side in one launch (anyseq_score_strip_device_multi).
Usage: gpu_multi_perf.py ROWS COLS [Ps=1,2,3]"""
import ctypes as C, os, sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import anyseq_b200 as A
from anyseq_b200 import capi
from anyseq_b200.capi import StripPartial, make_scoring
m, w = int(sys.argv[1]), int(sys.argv[2])
Ps = [int(x) for x in (sys.argv[3] if len(sys.argv) > 3 else "1,2,3").split(",")]
Ks = [int(x) for x in (sys.argv[4] if len(sys.argv) > 4 else "0").split(",")]
g = torch.Generator(device="cuda"); g.manual_seed(1)
lut = torch.tensor([65, 67, 71, 84], dtype=torch.uint8, device="cuda")
q = lut[torch.randint(0, 4, (m,), device="cuda", generator=g)]
s = lut[torch.randint(0, 4, (w,), device="cuda", generator=g)]
torch.cuda.synchronize()
al = A.Aligner(); L = capi.load_library(); vp = C.c_void_p
sc = make_scoring("semiglobal", 2, -1, -2, -1)
for P, K in [(P, K) for P in Ps for K in Ks]:
    al.tune(cols_per_lane=K, watchdog_ms=20000)
    best = 1e30
    for rep in range(2):
        parts = (StripPartial * P)()
        rc = L.anyseq_score_strip_device_multi(al.handle, C.byref(sc), P, (vp * P)(*[vp(q.data_ptr())] * P), m,
                                               (vp * P)(*[vp(s.data_ptr())] * P), 0, w, w, None, None, parts)
        assert rc == 0, L.anyseq_last_error()
        best = min(best, parts[0].kernel_ms)
    print(f"m={m} w={w} K={K} pairs per launch {P}: {best:.1f} ms {P*m*float(w)/best/1e6:.1f} GCUPS row_best={parts[P-1].row_best}", flush=True)
