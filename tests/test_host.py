"""Host-side logic on CPU: derived CIGAR view, workload generators, column slicing,
multi-GPU result combination (world_size-2 gloo), CLI argument surface, FASTA reader."""
import ctypes as C
import os
import subprocess
import sys
import tempfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cigar_view():
    import anyseq_b200 as A
    assert A.cigar(b"AC_GT  ", b"ACGG_  ") == "2=1I1=1D"
    assert A.cigar(b"   ", b"   ") == ""
    assert A.cigar(b" ACGT", b" ACCT") == "2=1X1="


def test_workloads_are_deterministic():
    from anyseq_b200 import workloads as W
    q1, s1, d = W.whole_genome_pair(0.001)
    q2, s2, _ = W.whole_genome_pair(0.001)
    assert (q1 == q2).all() and (s1 == s2).all() and len(q1) == 4641 and len(s1) == 4600
    assert set(np.unique(q1)) <= set(b"ACGT")
    r = W.read_batch(100)
    assert r[0].shape == (15000,) and r[2].shape == (50000,) and r[1][-1] == 15000 and r[3][-1] == 50000


def test_column_slices():
    from anyseq_b200.multigpu import column_slices
    for n in (4_600_000, 1000, 5):
        for w in (1, 2, 4, 8):
            sl = column_slices(n, w)
            assert sl[0][0] == 0 and sl[-1][1] == n and all(a[1] == b[0] for a, b in zip(sl, sl[1:]))
    assert all(c0 % 1024 == 0 for c0, _ in column_slices(4_600_000, 8))
    # wavefront slices: aligned, tiling, non-increasing widths, first about 18 % wider than the last on 8 ranks
    sl = column_slices(4_600_000, 8, rows=4_641_652)
    assert sl[0][0] == 0 and sl[-1][1] == 4_600_000 and all(a[1] == b[0] for a, b in zip(sl, sl[1:]))
    assert all(c0 % 1024 == 0 for c0, _ in sl)
    w = [b - a for a, b in sl]
    assert all(x >= y - 1024 for x, y in zip(w, w[1:])) and 1.05 < w[0] / w[-1] < 1.35
    assert column_slices(1000, 1, rows=500) == [(0, 1000)]


_WORKER = r'''
import os, sys, ctypes as C
sys.path.insert(0, sys.argv[1])
import numpy as np
import torch.distributed as dist
from anyseq_b200 import capi
from anyseq_b200.capi import StripPartial, Result, make_scoring
from oracle import oracle as O

dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
rng = np.random.default_rng(1)
ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
q = ACGT[rng.integers(0, 4, 300)]; s = ACGT[rng.integers(0, 4, 500)]
m, n = len(q), len(s)
# full DP matrix on every rank (tiny), then each rank reports only what it owns
H = np.zeros((m + 1, n + 1), dtype=np.int64)
for i in range(1, m + 1):
    for j in range(1, n + 1):
        H[i, j] = max(H[i-1, j-1] + (2 if q[i-1] == s[j-1] else -1), H[i, j-1] - 1, H[i-1, j] - 1)
c0, c1 = (0, 256) if rank == 0 else (256, n)
row = H[m, c0 + 1:c1 + 1]
part = StripPartial()
part.row_best = int(row.max()); part.row_best_j = int(c0 + int(np.argmax(row)))
col = H[1:, c1]
part.col_best = int(col.max()); part.col_best_i = int(np.argmax(col))
part.local_best = 0; part.corner = int(H[m, c1]); part.kernel_ms = 1.0; part.kernel_launches = 3
part.lenq = m; part.lens_total = n
parts = [None] * world
dist.all_gather_object(parts, bytes(part))
arr = (StripPartial * world)(*[StripPartial.from_buffer_copy(p) for p in parts])
L = capi.load_library()
sc = make_scoring("semiglobal", 2, -1, 0, -1)
res = Result()
assert L.anyseq_strip_combine(C.byref(sc), arr, world, C.byref(res)) == 0
want = O.score_linear("semiglobal", q, s)
assert (res.score, res.end_i, res.end_j) == want, ((res.score, res.end_i, res.end_j), want)   # end cell as on one GPU
assert res.kernel_launches == 3 * world
dist.destroy_process_group()
print("rank", rank, "ok", res.score)
'''


_WORKER_TOKENS = r'''
import os, sys, time, random
sys.path.insert(0, sys.argv[1])
import numpy as np
import torch
import torch.distributed as dist
from anyseq_b200.multigpu import RunTokens, shard_batch, pair_ranges
from oracle import oracle as O

dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
# --- run tokens: rank r may start run k only after rank r+1 has finished run k - depth
depth, runs = 2, 9
tok = RunTokens(dist, dist.new_group(backend="gloo"), rank, world, depth)
log = []
random.seed(rank)
for k in range(runs):
    tok.acquire(k)
    log.append(("start", k, time.monotonic()))
    time.sleep(random.random() * 0.02 * (3 if rank == world - 1 else 1))   # the last rank is the slow one
    log.append(("end", k, time.monotonic()))
    tok.release(k)
tok.drain(runs)
logs = [None] * world
dist.all_gather_object(logs, log)
for r in range(world - 1):
    start = {k: t for (e, k, t) in logs[r] if e == "start"}
    end_next = {k: t for (e, k, t) in logs[r + 1] if e == "end"}
    for k in range(depth, runs):
        assert end_next[k - depth] <= start[k] + 1e-4, (r, k)
# --- batch sharding: contiguous ranges, scores gathered in pair order
rng = np.random.default_rng(3)
ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
qs = [ACGT[rng.integers(0, 4, int(rng.integers(1, 40)))] for _ in range(23)]
ss = [ACGT[rng.integers(0, 4, int(rng.integers(1, 90)))] for _ in range(23)]
qo = np.concatenate([[0], np.cumsum([len(x) for x in qs])]); so = np.concatenate([[0], np.cumsum([len(x) for x in ss])])
qd, sd = np.concatenate(qs), np.concatenate(ss)
p0, p1, (qa, qb), (sa, sb), lqo, lso = shard_batch(qo, so, rank, world)
assert (p0, p1) == pair_ranges(23, world)[rank] and lqo[0] == 0 and lso[0] == 0
lq, ls = qd[qa:qb], sd[sa:sb]
local = np.array([O.score_linear("global", lq[lqo[i]:lqo[i + 1]], ls[lso[i]:lso[i + 1]])[0] for i in range(p1 - p0)], dtype=np.int32)
parts = [None] * world
dist.all_gather_object(parts, local.tobytes())
allsc = np.concatenate([np.frombuffer(b, dtype=np.int32) for b in parts])
want = [O.score_linear("global", qs[i], ss[i])[0] for i in range(23)]
assert allsc.tolist() == want
dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_run_tokens_and_batch_sharding_gloo():
    """N > 1 host logic on CPU (gloo, world_size 2): neighbour run tokens that let consecutive alignments
    overlap across ranks without overwriting an inbox in use, and contiguous sharding of a batch of pairs"""
    with tempfile.TemporaryDirectory() as td:
        w = os.path.join(td, "worker.py")
        open(w, "w").write(_WORKER_TOKENS)
        env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29519")
        r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                            "--master-addr", "127.0.0.1", "--master-port", "29519", w, ROOT],
                           capture_output=True, text=True, env=env, timeout=240)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        assert r.stdout.count("ok") == 2


def test_multi_rank_combine_gloo():
    """N > 1 host path: per-rank partial results gathered with torch.distributed (gloo,
    world_size 2) and combined exactly like a single-GPU semiglobal run"""
    with tempfile.TemporaryDirectory() as td:
        w = os.path.join(td, "worker.py")
        open(w, "w").write(_WORKER)
        env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29517")
        r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                            "--master-addr", "127.0.0.1", "--master-port", "29517", w, ROOT],
                           capture_output=True, text=True, env=env, timeout=240)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        assert r.stdout.count("ok") == 2


def _cli():
    from anyseq_b200 import build
    build.build()
    assert os.path.exists(build.CLI)
    return build.CLI


def test_cli_usage_and_errors():
    """argument surface of src/main.cpp:140-174,201-204 (no GPU needed for these paths)"""
    cli = _cli()
    r = subprocess.run([cli], capture_output=True, text=True)
    assert r.returncode == 0 and "SYNOPSIS" in r.stdout
    r = subprocess.run([cli, "--bogus"], capture_output=True, text=True)
    assert r.returncode == 0 and "Unknown command line arguments" in r.stdout and "'--bogus'" in r.stdout
    r = subprocess.run([cli, "-r", "0"], capture_output=True, text=True)
    assert r.returncode == 1 and "greater than zero" in r.stderr
    r = subprocess.run([cli, "-i", "/nonexistent/a.fa", "/nonexistent/b.fa"], capture_output=True, text=True)
    assert r.returncode == 1 and "can't open file" in r.stderr
    assert r.stdout.startswith("input sequences: /nonexistent/a.fa, /nonexistent/b.fa")


def test_bench_reference_arm_cpu():
    """bench.py --impl reference runs the oracle port on host cores and prints the contract line"""
    import json
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0", "--cpu-sample", "3000", "--scale", "0.01"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "GCUPS" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["e2e"]["h2d_bytes_per_step"] == 0


def test_reader_and_printer_match_the_reference_host_code():
    """host surface: our FASTA/FASTQ reader against the answers of the reference's OWN reader
    (src/sequence_io.cpp:62-241, compiled where it lies by tests/golden/make_reader_golden.py) on tricky files: multi-line
    records, CRLF (the '\\r' stays in the data), blank / comment lines, truncated FASTQ, extension vs first-character
    sniffing, empty files, skip().  When the reference sources are present the comparison is also made live."""
    import json
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import make_reader_golden as G
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "reader_golden.json")))
    assert gold["files"] == G.FILES
    with tempfile.TemporaryDirectory() as td:
        exe = os.path.join(td, "seqio_dump_ours")
        subprocess.run(["/usr/bin/g++", "-O1", "-std=c++17", "-I" + os.path.join(ROOT, "anyseq_b200", "csrc"),
                        os.path.join(ROOT, "tests", "host", "seqio_dump_ours.cpp"),
                        os.path.join(ROOT, "anyseq_b200", "csrc", "sequence_io.cpp"), "-o", exe], check=True)
        ours = G.run_all(exe, td)
        assert ours == gold["expected"]
        assert any(v["rc"] == 1 for v in ours.values()) and any("\t10\t" in v["stdout"] for v in ours.values())
        if os.path.exists(os.path.join(G.REF, "sequence_io.cpp")):
            assert G.run_all(G.build_ref_driver(td), td) == ours
        # print_alignment (src/alignment_io.cpp:13-38), same scheme
        assert gold["print_alignment_cases"] == G.ALN_CASES
        exe = os.path.join(td, "alnio_dump_ours")
        subprocess.run(["/usr/bin/g++", "-O1", "-std=c++17", "-I" + os.path.join(ROOT, "anyseq_b200", "csrc"),
                        os.path.join(ROOT, "tests", "host", "alnio_dump_ours.cpp"),
                        os.path.join(ROOT, "anyseq_b200", "csrc", "alignment_io.cpp"), "-o", exe], check=True)
        mine = G.run_printer(exe)
        assert mine == gold["print_alignment"] and mine["stdout"].count("<<<") == len(G.ALN_CASES)
        if os.path.exists(os.path.join(G.REF, "alignment_io.cpp")):
            assert G.run_printer(G.build_ref_printer(td)) == mine


_TB_WORKER = r"""
import os, sys
sys.path.insert(0, sys.argv[1])
import numpy as np
import torch.distributed as dist
from anyseq_b200.multigpu import merge_regions, merge_splits, traceback_half_owner

dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
m, n = 700, 900
full = (b"Q" * 350 + b"_" * 450 + b"q" * 800, b"S" * 1600)
# rank r owns the output columns of its blocks: here simply two unequal contiguous ranges
lo, hi = (0, 650) if rank == 0 else (650, m + n)
splits = [0, 100, -1, -1, 700] if rank == 0 else [0, -1, 300, 420, 700]      # -1 = decided by the other rank
pieces = [None] * world
dist.all_gather_object(pieces, (lo, hi, full[0][lo:hi], full[1][lo:hi], splits))
aq, as_ = merge_regions(m, n, [(p[0], p[1], p[2], p[3]) for p in pieces])
assert (aq, as_) == full
assert merge_splits([p[4] for p in pieces]) == [0, 100, 300, 420, 700]
# a gap between the regions is an error, not silently blanked
try:
    merge_regions(m, n, [(0, 600, full[0][:600], full[1][:600]), (650, m + n, full[0][650:], full[1][650:])])
    raise SystemExit("gap not detected")
except ValueError:
    pass
# ownership rule of the halves (csrc/traceback.cu: traceback_half_owner): level 0 of 2 ranks = one half each, later
# levels whole parts; every half has exactly one owner and the owners are monotone in h
for w in (1, 2, 4, 8):
    for np_full in (1, 2, 4, 8, 16, 64):
        owners = [traceback_half_owner(h, np_full, w) for h in range(2 * np_full)]
        assert owners == sorted(owners) and 0 <= owners[0] and owners[-1] < w
        if 2 * np_full >= w:
            assert sorted(set(owners)) == list(range(w))
            if np_full >= w:
                assert all(owners[2 * p] == owners[2 * p + 1] for p in range(np_full))
assert [traceback_half_owner(h, 1, 2) for h in range(2)] == [0, 1]
assert [traceback_half_owner(h, 1, 8) for h in range(2)] == [0, 4]
dist.destroy_process_group()
print("rank", rank, "ok")
"""


def test_sharded_traceback_gather_gloo():
    """host side of the multi-GPU traceback on CPU (gloo, world_size 2): all-gather of the per-rank pieces and split
    rows, their merge, and the ownership rule of the Hirschberg halves"""
    import subprocess, sys, tempfile
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "worker.py")
        with open(path, "w") as f:
            f.write(_TB_WORKER)
        r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                            "--master-addr", "127.0.0.1", "--master-port", "29641", path, ROOT],
                           capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stdout + r.stderr
        assert r.stdout.count("ok") == 2


def test_launch_planner_without_a_device():
    """anyseq_plan_launch: the engine's launch planning (strip width, tile height, bands, grid) is host logic and runs
    without a GPU.  The expectations are the measured optima recorded in profiles/r02_summary.md."""
    import anyseq_b200 as A
    # BASELINE configs[1]: the whole-genome pair -- widest strips, three warps per scheduler, decoupled Gotoh cells
    p = A.plan_launch("semiglobal", 4641652, 4600000, affine=True)
    assert (p["cols_per_lane"], p["rows_per_step"], p["cell_form"], p["warps_per_scheduler"]) == (32, 2, 1, 3)
    assert p["strips"] == (4600000 + 1023) // 1024 and p["grid"] == 148 and p["warps_per_cta"] == 12
    assert p["bands"] > 1 and p["band_rows"] % 32 == 0 and p["bands"] * p["band_rows"] >= 4641652
    assert p["first_items"] == 148 * 12
    # wide LOCAL Gotoh launches take the mixed cells
    assert A.plan_launch("local", 4641652, 4600000, affine=True)["cell_form"] == 2
    # one rank's slice of the 8-GPU wavefront: 1124 strips of 512 columns, two warps per scheduler, ONE band
    p = A.plan_launch("semiglobal", 4641652, 575488, affine=True, chained=True)
    assert (p["cols_per_lane"], p["strips"], p["warps_per_scheduler"], p["bands"]) == (16, 1124, 2, 1)
    assert p["first_items"] == 1124 and p["grid"] == 148 and p["warps_per_cta"] == 8
    # small problems: strip width from the measured critical-path fit (profiles/r02_c1_table_fit.log)
    expect = {(8087, 9011, False): 16, (8087, 18022, False): 16, (64, 9011, False): 16, (1024, 9011, False): 16,
              (8087, 128, False): 4, (16174, 9011, False): 4, (32348, 9011, False): 4,
              (8087, 9011, True): 4, (8087, 18022, True): 16, (1024, 9011, True): 16, (32348, 9011, True): 4, (2048, 2048, True): 4}
    for (m, n, affine), K in expect.items():
        p = A.plan_launch("semiglobal" if affine else "global", m, n, affine=affine)
        assert p["cols_per_lane"] == K, (m, n, affine, p)
        assert p["bands"] == 1 and p["strips"] == -(-n // (32 * K)) and p["first_items"] == p["strips"]
        assert p["grid"] == min(148, -(-p["strips"] // 4)) and p["warps_per_cta"] == 4
    # ... but never for a rank of a multi-GPU wavefront (its rows are shared with the other ranks' slices)
    assert A.plan_launch("global", 8087, 9011, chained=True)["cols_per_lane"] == 4
    # more strips than warps: two or three warps per scheduler by how full the rounds of items are
    # (profiles/r02_perf_1m_warps_per_scheduler.log: 1 Mbp x 1 Mbp 3166 GCUPS with two, 2585 with three)
    p = A.plan_launch("semiglobal", 1000000, 1000000, affine=True)
    assert (p["cols_per_lane"], p["strips"], p["warps_per_scheduler"], p["bands"], p["warps_per_cta"]) == (16, 1954, 2, 3, 8)
    # the 2-GPU slice of the whole-genome pair keeps three (as measured in the N = 2 bench line)
    assert A.plan_launch("semiglobal", 4641652, 2300000, affine=True, chained=True)["warps_per_scheduler"] == 3
    # a smaller GPU gets a smaller grid and narrower strips for the same problem
    assert A.plan_launch("semiglobal", 4641652, 4600000, affine=True, sm_count=74)["grid"] == 74
    with pytest.raises(A.AnyseqError):
        A.plan_launch("global", 0, 10)
