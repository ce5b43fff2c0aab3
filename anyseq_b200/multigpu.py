"""Multi-GPU column-strip wavefront for one long pair.

One process per GPU (torch.distributed supplies rendezvous + the tiny
exchanges); rank r owns a contiguous slice of subject columns.  The only
data-path exchange is the strip-boundary column (H, and E for Gotoh), streamed
32 rows at a time by the producing rank's strip kernel straight into the next
rank's inbox over NVLink (peer stores + a system-scope release counter) --
no host, no NCCL call inside the wavefront.
"""
from __future__ import annotations

import ctypes as C

from . import capi
from .api import Aligner, AlignmentResult, ScoringScheme
from .capi import Result, StripPartial, make_scoring


def column_slices(n: int, world: int, align: int = 1024):
    """equal slices of [0, n), boundaries rounded to `align` columns"""
    cuts = [0]
    for r in range(1, world):
        c = (n * r // world) // align * align
        cuts.append(max(c, cuts[-1]))
    cuts.append(n)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


class StripWavefront:
    def __init__(self, aligner: Aligner, rank: int, world: int, rows: int, dist=None):
        self.al, self.rank, self.world, self.rows, self.dist = aligner, rank, world, rows, dist
        self._lib = capi.load_library()
        self.inbox = C.c_void_p()
        self.next_inbox = C.c_void_p()
        handle = (C.c_ubyte * 64)()
        if world > 1 and rank > 0:
            aligner._check(self._lib.anyseq_strip_inbox_create(aligner.handle, rows, C.byref(self.inbox), handle))
        if world > 1:
            handles = [None] * world
            dist.all_gather_object(handles, bytes(handle))
            if rank < world - 1:
                buf = (C.c_ubyte * 64).from_buffer_copy(handles[rank + 1])
                aligner._check(self._lib.anyseq_strip_inbox_open(aligner.handle, buf, rows, C.byref(self.next_inbox)))

    def reset(self):
        """re-synchronise the run counters of both ends of every inbox and clear the owned
        one; every rank must call it (ends with a barrier).  Not needed between runs: border
        records carry a per-run tag."""
        if self.inbox:
            self.al._check(self._lib.anyseq_strip_inbox_reset(self.al.handle, self.inbox))
        if self.next_inbox:
            self.al._check(self._lib.anyseq_strip_inbox_reset(self.al.handle, self.next_inbox))
        if self.world > 1:
            self.dist.barrier()

    def run(self, mode, scoring: ScoringScheme, d_query: int, m: int, d_subject_slice: int,
            col_begin: int, col_end: int, n_total: int) -> StripPartial:
        sc = make_scoring(mode, scoring.same, scoring.diff, scoring.gap_init, scoring.gap_extend)
        part = StripPartial()
        self.al._check(self._lib.anyseq_score_strip_device(
            self.al.handle, C.byref(sc), C.c_void_p(d_query), m, C.c_void_p(d_subject_slice),
            col_begin, col_end, n_total, self.inbox if self.inbox else None,
            self.next_inbox if self.next_inbox else None, C.byref(part)))
        return part

    def combine(self, mode, scoring: ScoringScheme, part: StripPartial) -> AlignmentResult:
        """gather the per-rank partials and combine them exactly like a single-GPU run"""
        parts = [bytes(part)]
        if self.world > 1:
            parts = [None] * self.world
            self.dist.all_gather_object(parts, bytes(part))
        arr = (StripPartial * self.world)(*[StripPartial.from_buffer_copy(p) for p in parts])
        sc = make_scoring(mode, scoring.same, scoring.diff, scoring.gap_init, scoring.gap_extend)
        res = Result()
        self.al._check(self._lib.anyseq_strip_combine(C.byref(sc), arr, self.world, C.byref(res)))
        return AlignmentResult(res.score, res.end_i, res.end_j, res.kernel_ms, res.kernel_launches)

    def close(self):
        if self.inbox:
            self._lib.anyseq_strip_inbox_destroy(self.al.handle, self.inbox)
            self.inbox = C.c_void_p()
        if self.next_inbox:
            self._lib.anyseq_strip_inbox_destroy(self.al.handle, self.next_inbox)
            self.next_inbox = C.c_void_p()
