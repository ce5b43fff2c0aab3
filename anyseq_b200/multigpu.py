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

import numpy as np

from . import capi
from .api import Aligner, AlignmentResult, ScoringScheme
from .capi import Result, StripPartial, make_scoring


def _strip_cols(width: int, sm_count: int = 148) -> int:
    """strip width the engine picks for a slice of `width` columns (engine.cu: pick_K, small alphabets)"""
    for k in (32, 16, 8):
        if width // (32 * k) >= 7 * sm_count:
            return 32 * k
    return 128


def column_slices(n: int, world: int, align: int = 1024, rows: int = 0, strip_cols: int = 0, lag_rows: int = 100,
                  sm_count: int = 148):
    """slices of [0, n), boundaries rounded to `align` columns.  rows == 0 (default, what bench.py uses): equal slices.
    rows > 0: slices that shrink geometrically from rank to rank (a slice is 1 / (1 + f) of its left neighbour, f = strips
    of a slice * lag_rows / rows), capped at two warps per scheduler.  The idea -- later ranks start later, so give them
    less -- does NOT pay while every strip of a slice has its own warp: the time of such a slice is rows x (time per
    row of a strip) whatever its width (measured on 8 B200s, tools/chain_probe.py: 606 208 and 544 768 columns both take
    831 ms alone), so the makespan is the pace of a strip plus the start delay of the last rank either way (988 ms with
    equal slices, 1011 ms with shrinking ones).  Kept for launches that run several bands per slice."""
    weights = [1.0] * world
    cap = float(n)
    if rows > 0 and world > 1:
        sc = strip_cols or _strip_cols(n // world)
        f = (n / world / sc) * lag_rows / float(rows)
        weights = [(1.0 / (1.0 + f)) ** r for r in range(world)]
        # a slice must not need more strips than two warps per scheduler can take (the most loaded scheduler sets the
        # pace of a single-band chain, engine.cu: default_blocks_per_sm), if the other ranks can absorb the rest
        if sc * 8 * sm_count * world >= n:
            cap = float(sc * 8 * sm_count)
    widths = [n * w / sum(weights) for w in weights]
    for _ in range(world):                       # clip to the cap, hand the excess to the unclipped slices
        over = sum(max(0.0, w - cap) for w in widths)
        free = [i for i, w in enumerate(widths) if w < cap]
        if over <= 0 or not free:
            break
        fsum = sum(widths[i] for i in free)
        widths = [min(w, cap) for w in widths]
        for i in free:
            widths[i] += over * widths[i] / fsum
    cuts, acc = [0], 0.0
    for r in range(world - 1):
        acc += widths[r]
        c = int(acc) // align * align
        cuts.append(min(max(c, cuts[-1]), n))
    cuts.append(n)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


# ---------------------------------------------------------------------------
# batches of independent pairs: contiguous ranges per rank, no data-path collective
# ---------------------------------------------------------------------------
def pair_ranges(npairs: int, world: int):
    """contiguous, near-equal ranges of pair ids, one per rank"""
    return [(npairs * r // world, npairs * (r + 1) // world) for r in range(world)]


def shard_batch(q_off, s_off, rank: int, world: int):
    """-> (p0, p1, byte range of the queries, byte range of the subjects, rebased offsets) of this rank"""
    import numpy as np
    qo = np.asarray(q_off, dtype=np.int64)
    so = np.asarray(s_off, dtype=np.int64)
    p0, p1 = pair_ranges(len(qo) - 1, world)[rank]
    return p0, p1, (int(qo[p0]), int(qo[p1])), (int(so[p0]), int(so[p1])), qo[p0:p1 + 1] - qo[p0], so[p0:p1 + 1] - so[p0]


def score_batch_sharded(aligner: Aligner, dist, rank: int, world: int, mode, queries, q_off, subjects, s_off,
                        scoring: ScoringScheme, gather: bool = True):
    """every rank scores its contiguous range of pairs on its own GPU; the only communication is the
    gather of the int32 scores at the end (gather=False: each rank keeps just its range)"""
    import numpy as np
    p0, p1, (qa, qb), (sa, sb), qo, so = shard_batch(q_off, s_off, rank, world)
    local, res = aligner.score_batch(mode, capi.as_u8(queries)[qa:qb], qo, capi.as_u8(subjects)[sa:sb], so, scoring)
    if not gather or world == 1:
        return local, res
    parts = [None] * world
    dist.all_gather_object(parts, local.tobytes())
    return np.concatenate([np.frombuffer(b, dtype=np.int32) for b in parts]), res


class RunTokens:
    """Host-side flow control between neighbouring ranks for back-to-back runs.

    Run k of rank r streams its right border into inbox k % depth of rank r + 1, so it must not
    start before rank r + 1 has FINISHED run k - depth (the previous user of those rows).  Rank
    r + 1 says so with a token (the run index, one int64) sent to rank r after every run; rank r
    receives the token of run k - depth before it starts run k.  With depth >= 2 the wait is
    already satisfied in steady state and consecutive alignments overlap: rank 0 starts the next
    pair while the wavefront of the previous one is still travelling through the later ranks.
    `group` must be a CPU (gloo) group: tokens never touch a GPU stream."""

    def __init__(self, dist, group, rank: int, world: int, depth: int):
        import torch
        self._torch = torch
        self.dist, self.group, self.rank, self.world, self.depth = dist, group, rank, world, depth
        self._pending = []

    def acquire(self, k: int):
        if self.world > 1 and self.rank < self.world - 1 and k >= self.depth:
            t = self._torch.zeros(1, dtype=self._torch.int64)
            self.dist.recv(t, src=self.rank + 1, group=self.group)
            if int(t.item()) != k - self.depth:
                raise RuntimeError(f"run token out of order: got {int(t.item())}, expected {k - self.depth}")

    def release(self, k: int):
        if self.world > 1 and self.rank > 0:
            t = self._torch.full((1,), k, dtype=self._torch.int64)
            self._pending.append((self.dist.isend(t, dst=self.rank - 1, group=self.group), t))
            self._pending = [(w, t) for (w, t) in self._pending if not w.is_completed()]

    def drain(self, runs: int):
        """receive the tokens nobody waited for (end of a stream of `runs` runs; every rank calls it)"""
        if self.world > 1 and self.rank < self.world - 1:
            for k in range(max(0, runs - self.depth), runs):
                t = self._torch.zeros(1, dtype=self._torch.int64)
                self.dist.recv(t, src=self.rank + 1, group=self.group)
        for (w, _) in self._pending:
            w.wait()
        self._pending = []


class StripWavefront:
    """depth = number of inboxes per rank boundary.  depth 1: every rank must have finished run k
    before any rank starts run k + 1 (the caller puts a barrier between runs).  depth >= 2:
    back-to-back runs are flow-controlled by RunTokens and overlap (no barrier between runs)."""

    def __init__(self, aligner: Aligner, rank: int, world: int, rows: int, dist=None, depth: int = 1,
                 pairs_per_launch: int = 1):
        """pairs_per_launch > 1: run_multi() relaxes that many alignments of one shape in a single launch (a narrow
        slice alone cannot fill a B200); every launch then uses its own set of pairs_per_launch inboxes."""
        self.al, self.rank, self.world, self.rows, self.dist = aligner, rank, world, rows, dist
        self.depth = max(1, int(depth)) if world > 1 else 1
        self.ppl = max(1, int(pairs_per_launch))
        self._lib = capi.load_library()
        self.inboxes = [C.c_void_p() for _ in range(self.depth * self.ppl)]
        self.next_inboxes = [C.c_void_p() for _ in range(self.depth * self.ppl)]
        self.k = 0
        self.tokens = None
        for d in range(self.depth * self.ppl):
            handle = (C.c_ubyte * 64)()
            if world > 1 and rank > 0:
                aligner._check(self._lib.anyseq_strip_inbox_create(aligner.handle, rows, C.byref(self.inboxes[d]), handle))
            if world > 1:
                handles = [None] * world
                dist.all_gather_object(handles, bytes(handle))
                if rank < world - 1:
                    buf = (C.c_ubyte * 64).from_buffer_copy(handles[rank + 1])
                    aligner._check(self._lib.anyseq_strip_inbox_open(aligner.handle, buf, rows, C.byref(self.next_inboxes[d])))
        if world > 1 and self.depth > 1:
            self.tokens = RunTokens(dist, dist.new_group(backend="gloo"), rank, world, self.depth)

    # single-inbox views (depth 1 users)
    @property
    def inbox(self):
        return self.inboxes[0]

    @property
    def next_inbox(self):
        return self.next_inboxes[0]

    def reset(self):
        """re-synchronise the run counters of both ends of every inbox and clear the owned
        ones; every rank must call it (ends with a barrier).  Not needed between runs: border
        records carry a per-run tag."""
        if self.tokens is not None:
            self.tokens.drain(self.k)
        for h in self.inboxes + self.next_inboxes:
            if h:
                self.al._check(self._lib.anyseq_strip_inbox_reset(self.al.handle, h))
        self.k = 0
        if self.world > 1:
            self.dist.barrier()

    def run(self, mode, scoring: ScoringScheme, d_query: int, m: int, d_subject_slice: int,
            col_begin: int, col_end: int, n_total: int) -> StripPartial:
        sc = make_scoring(mode, scoring.same, scoring.diff, scoring.gap_init, scoring.gap_extend)
        part = StripPartial()
        k = self.k
        slot = (k % self.depth) * self.ppl
        inbox, nxt = self.inboxes[slot], self.next_inboxes[slot]
        if self.tokens is not None:
            self.tokens.acquire(k)
        self.al._check(self._lib.anyseq_score_strip_device(
            self.al.handle, C.byref(sc), C.c_void_p(d_query), m, C.c_void_p(d_subject_slice),
            col_begin, col_end, n_total, inbox if inbox else None, nxt if nxt else None, C.byref(part)))
        if self.tokens is not None:
            self.tokens.release(k)
        self.k = k + 1
        return part

    def run_host(self, mode, scoring: ScoringScheme, query: np.ndarray, subject_slice: np.ndarray,
                 col_begin: int, col_end: int, n_total: int) -> StripPartial:
        """run() with HOST buffers (this rank's query and subject slice): anyseq_score_strip copies them to the device"""
        sc = make_scoring(mode, scoring.same, scoring.diff, scoring.gap_init, scoring.gap_extend)
        part = StripPartial()
        k = self.k
        slot = (k % self.depth) * self.ppl
        inbox, nxt = self.inboxes[slot], self.next_inboxes[slot]
        if self.tokens is not None:
            self.tokens.acquire(k)
        self.al._check(self._lib.anyseq_score_strip(
            self.al.handle, C.byref(sc), C.c_void_p(query.ctypes.data), len(query), C.c_void_p(subject_slice.ctypes.data),
            col_begin, col_end, n_total, inbox if inbox else None, nxt if nxt else None, C.byref(part)))
        if self.tokens is not None:
            self.tokens.release(k)
        self.k = k + 1
        return part

    def run_multi(self, mode, scoring: ScoringScheme, d_queries, m: int, d_subject_slices,
                  col_begin: int, col_end: int, n_total: int):
        """len(d_queries) <= pairs_per_launch alignments of one shape in ONE launch -> list of StripPartial"""
        npairs = len(d_queries)
        if npairs < 1 or npairs > self.ppl or npairs != len(d_subject_slices):
            raise ValueError("run_multi: between 1 and pairs_per_launch pairs")
        sc = make_scoring(mode, scoring.same, scoring.diff, scoring.gap_init, scoring.gap_extend)
        k = self.k
        slot = (k % self.depth) * self.ppl
        vp = C.c_void_p
        qs = (vp * npairs)(*[vp(int(x)) for x in d_queries])
        ss = (vp * npairs)(*[vp(int(x)) for x in d_subject_slices])
        has_in = bool(self.inboxes[slot])
        has_nx = bool(self.next_inboxes[slot])
        ins = (vp * npairs)(*[self.inboxes[slot + p] for p in range(npairs)]) if has_in else None
        nxs = (vp * npairs)(*[self.next_inboxes[slot + p] for p in range(npairs)]) if has_nx else None
        parts = (StripPartial * npairs)()
        if self.tokens is not None:
            self.tokens.acquire(k)
        self.al._check(self._lib.anyseq_score_strip_device_multi(
            self.al.handle, C.byref(sc), npairs, qs, m, ss, col_begin, col_end, n_total, ins, nxs, parts))
        if self.tokens is not None:
            self.tokens.release(k)
        self.k = k + 1
        return [StripPartial.from_buffer_copy(bytes(parts[p])) for p in range(npairs)]

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
        if self.tokens is not None:
            self.tokens.drain(self.k)
            self.k = 0
        for lst in (self.inboxes, self.next_inboxes):
            for i, h in enumerate(lst):
                if h:
                    self._lib.anyseq_strip_inbox_destroy(self.al.handle, h)
                    lst[i] = C.c_void_p()


# --------------------------------------------------------------------------- multi-GPU linear-space traceback
def traceback_half_owner(h: int, np_full: int, world: int) -> int:
    """rank that relaxes half h (= 2 * part + side) of a Hirschberg level with np_full parts -- the rule of
    TracebackShard (csrc/engine.cuh): whole parts per rank once 2 * np_full >= world, before that one half per rank"""
    halves = 2 * np_full
    return h // (halves // world) if halves >= world else h * (world // halves)


def merge_regions(lenq: int, lens: int, pieces):
    """pieces: per rank (lo, hi, aligned_query[lo:hi], aligned_subject[lo:hi]) in rank order -> the two full strings.
    The ranges of the ranks tile [0, lenq + lens) (a rank without blocks contributes an empty range)."""
    total = lenq + lens
    aq, as_ = bytearray(b" " * total), bytearray(b" " * total)
    pos = 0
    for lo, hi, q, s in pieces:
        if hi <= lo:
            continue
        if lo != pos or len(q) != hi - lo or len(s) != hi - lo:
            raise ValueError(f"traceback regions do not tile the output: expected a piece starting at {pos}, got [{lo}, {hi})")
        aq[lo:hi] = q
        as_[lo:hi] = s
        pos = hi
    if pos != total:
        raise ValueError(f"traceback regions end at {pos}, not at {total}")
    return bytes(aq), bytes(as_)


def merge_splits(per_rank):
    """split rows of all ranks (anyseq_last_splits, -1 = decided elsewhere) -> the single-GPU splits vector"""
    out = np.max(np.asarray(per_rank, dtype=np.int64), axis=0)
    return [int(x) for x in out]


class _DevView:
    """zero-copy view of raw device memory for torch.as_tensor (CUDA array interface)"""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


class ShardedTraceback:
    """anyseq_align_sharded over torch.distributed: one process per GPU, world a power of two.  The library calls back
    for every exchange of the first log2(world) Hirschberg levels (a broadcast of last-column border records, 16 bytes
    per query row and half); it runs on NCCL over the raw device buffer.  gather() assembles the strings on every rank."""

    def __init__(self, aligner: Aligner, rank: int, world: int, dist=None):
        import torch
        self.al, self.rank, self.world, self.dist, self._torch = aligner, rank, world, dist, torch
        self._lib = capi.load_library()
        self.bcast_bytes = 0

        def _bcast(user, d_buffer, nbytes, src):
            try:
                t = torch.as_tensor(_DevView(d_buffer, nbytes), device="cuda")
                self.dist.broadcast(t, src=int(src))
                torch.cuda.synchronize()
                self.bcast_bytes += int(nbytes)
                return 0
            except Exception as e:      # never let an exception cross the C boundary
                print("ShardedTraceback: broadcast failed:", e, flush=True)
                return 1

        self._cb = capi.BCAST_FN(_bcast)

    def align(self, mode, query, subject, scoring: ScoringScheme):
        """-> (lo, hi, aligned_query[lo:hi], aligned_subject[lo:hi], splits of this rank, kernel_ms)"""
        q, s = capi.as_u8(query), capi.as_u8(subject)
        m, n = len(q), len(s)
        sc = make_scoring(mode, scoring.same, scoring.diff, scoring.gap_init, scoring.gap_extend)
        aq = np.empty(m + n, dtype=np.uint8)
        as_ = np.empty(m + n, dtype=np.uint8)
        lo, hi = C.c_int64(), C.c_int64()
        res = Result()
        self.al._check(self._lib.anyseq_align_sharded(
            self.al.handle, C.byref(sc), capi._ptr(q), m, capi._ptr(s), n, self.rank, self.world, self._cb, None,
            capi._ptr(aq), capi._ptr(as_), C.byref(lo), C.byref(hi), C.byref(res)))
        return lo.value, hi.value, aq[lo.value:hi.value].tobytes(), as_[lo.value:hi.value].tobytes(), self.al.last_splits(), res.kernel_ms

    def gather(self, m: int, n: int, piece):
        """all-gather of the pieces -> (aligned_query, aligned_subject, splits) identical on every rank"""
        lo, hi, q, s, splits, _ = piece
        if self.world > 1 and self.dist.get_backend() == "nccl":
            # the strings over NVLink: every rank fills its columns of a zeroed device buffer, one all-reduce (the
            # ranges are disjoint); the ranges and split rows (small) as objects
            torch = self._torch
            total = m + n
            buf = torch.zeros(2 * total, dtype=torch.uint8, device="cuda")
            if hi > lo:
                buf[lo:hi] = torch.frombuffer(bytearray(q), dtype=torch.uint8).cuda()
                buf[total + lo: total + hi] = torch.frombuffer(bytearray(s), dtype=torch.uint8).cuda()
            self.dist.all_reduce(buf, op=self.dist.ReduceOp.SUM)
            meta = [None] * self.world
            self.dist.all_gather_object(meta, (lo, hi, splits))
            pos = 0
            for (l, h, _) in meta:                      # the ranges must tile [0, m + n)
                if h > l:
                    if l != pos:
                        raise ValueError(f"traceback regions do not tile the output: expected a piece starting at {pos}, got [{l}, {h})")
                    pos = h
            if pos != total:
                raise ValueError(f"traceback regions end at {pos}, not at {total}")
            host = buf.cpu().numpy()
            return host[:total].tobytes(), host[total:].tobytes(), merge_splits([x[2] for x in meta])
        pieces = [(lo, hi, q, s, splits)]
        if self.world > 1:
            pieces = [None] * self.world
            self.dist.all_gather_object(pieces, (lo, hi, q, s, splits))
        aq, as_ = merge_regions(m, n, [(p[0], p[1], p[2], p[3]) for p in pieces])
        return aq, as_, merge_splits([p[4] for p in pieces])
