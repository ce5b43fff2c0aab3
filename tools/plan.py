#!/usr/bin/env python
"""Print the launch the engine plans for score-only problems (no GPU needed):  plan.py M N [affine=0] [mode=global] [sms=148]"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import anyseq_b200 as A  # noqa: E402

m, n = int(sys.argv[1]), int(sys.argv[2])
affine = bool(int(sys.argv[3])) if len(sys.argv) > 3 else False
mode = sys.argv[4] if len(sys.argv) > 4 else "global"
sms = int(sys.argv[5]) if len(sys.argv) > 5 else 148
for chained in (False, True):
    print("rank of a multi-GPU wavefront:" if chained else "whole problem on one GPU:   ", A.plan_launch(mode, m, n, affine, sms, chained))
