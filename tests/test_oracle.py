"""CPU tests of the parity oracle (no GPU): the restated reference CPU path against
the known answers of SURVEY.md Appendix C, against independent textbook DPs, and
against the frozen golden fixtures."""
import hashlib

import numpy as np
import pytest

ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
MODES = ("global", "semiglobal", "local")


def _rand(rng, n):
    return ACGT[rng.integers(0, 4, n)]


def test_reference_rng_inputs(oracle, golden):
    """inputs of `align -r ...` regenerated with the reference's recipe (src/main.cpp:90-120,207-209)"""
    for (lo, hi), key in (((256, 1024), "align -r"), ((10000, 1024), "align -r 10000"), ((10000, 10000), "align -r 10000 10000")):
        g = golden["appendix_c"][key]
        q, s = oracle.reference_random_pair(lo, hi)
        assert (len(q), len(s)) == (g["m"], g["n"])
        assert "%016x" % oracle.fnv1a64(q) == g["fnv_q"]
        assert "%016x" % oracle.fnv1a64(s) == g["fnv_s"]
        assert bytes(q[:16]) == b"CGTACCAGCCGAGGTC"


def test_reference_rng_inputs_against_the_reference_generator(golden):
    """the same input hashes, produced by the reference's OWN generator: src/main.cpp is included where it lies (main()
    renamed) by tests/host/refinput_dump_ref.cpp and random_string() is called with the default-seeded mt19937_64 --
    this pins the inputs of every Appendix C fixture to real reference code, not to a restatement"""
    import os, subprocess, tempfile
    ref = "/root/reference/src"
    if not os.path.exists(os.path.join(ref, "main.cpp")):
        pytest.skip("reference sources not on this machine")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with tempfile.TemporaryDirectory() as td:
        exe = os.path.join(td, "refinput")
        subprocess.run(["/usr/bin/g++", "-O1", "-std=c++14", "-I" + ref, os.path.join(root, "tests", "host", "refinput_dump_ref.cpp"),
                        os.path.join(ref, "sequence_io.cpp"), os.path.join(ref, "alignment_io.cpp"), "-o", exe], check=True)
        out = subprocess.run([exe, "256", "1024", "10000", "1024", "10000", "10000"], capture_output=True, text=True, check=True).stdout
    rows = [ln.split() for ln in out.strip().splitlines()]
    for row, key in zip(rows, ("align -r", "align -r 10000", "align -r 10000 10000")):
        g = golden["appendix_c"][key]
        assert [int(row[0]), row[1], int(row[2]), row[3]] == [g["m"], g["fnv_q"], g["n"], g["fnv_s"]], key


@pytest.mark.parametrize("key,lo,hi", [("align -r", 256, 1024), ("align -r 10000", 10000, 1024),
                                       ("align -r 10000 10000", 10000, 10000)])
def test_appendix_c_scores(oracle, golden, key, lo, hi):
    q, s = oracle.reference_random_pair(lo, hi)
    want = golden["appendix_c"][key]["scores"]
    assert [oracle.score_linear(m, q, s)[0] for m in MODES] == want
    assert [oracle.textbook_linear(m, q, s) for m in MODES] == want


def test_appendix_c_traceback(oracle, golden):
    g = golden["appendix_c"]["align -r"]
    q, s = oracle.reference_random_pair(256, 1024)
    for k, mode in enumerate(MODES):
        ret, aq, as_, sp = oracle.traceback_lintime(mode, q, s)
        assert ret == g["legacy_return"][k]                       # quirk Q1
        assert hashlib.sha256(aq + b"\n" + as_).hexdigest()[:16] == g["sha"][mode]
        assert sum(1 for a, b in zip(aq, as_) if not (a == 32 and b == 32)) == g["nonblank"][mode]
        assert oracle.column_score(aq, as_) == g["column_score"][mode]
        if mode == "global":
            assert sp.tolist() == g["splits_global"]


def test_semiglobal_end_cell(oracle):
    q, s = oracle.reference_random_pair(256, 1024)
    assert oracle.score_linear("semiglobal", q, s) == (659, 860, 911)
    assert oracle.score_linear("local", q, s) == (659, 860, 911)


def test_restatement_equals_textbook_random(oracle):
    rng = np.random.default_rng(7)
    for _ in range(60):
        m, n = int(rng.integers(1, 400)), int(rng.integers(1, 400))
        q, s = _rand(rng, m), _rand(rng, n)
        same, diff, gap = int(rng.integers(1, 6)), -int(rng.integers(0, 5)), -int(rng.integers(0, 5))
        for mode in MODES:
            assert oracle.score_linear(mode, q, s, same, diff, gap)[0] == oracle.textbook_linear(mode, q, s, same, diff, gap)
            gi, ge = -int(rng.integers(1, 12)), -int(rng.integers(0, 4))
            assert oracle.score_affine(mode, q, s, same, diff, gi, ge)[0] == oracle.textbook_affine(mode, q, s, same, diff, gi, ge)


def test_scores_do_not_depend_on_blocking(oracle):
    """any dependency-respecting order gives the same H"""
    rng = np.random.default_rng(11)
    q, s = _rand(rng, 700), _rand(rng, 900)
    for mode in MODES:
        ref = oracle.score_linear(mode, q, s)[0]
        for bw, bh in ((1024, 1024), (128, 1280), (64, 64), (100, 37), (7, 5000)):
            assert oracle.score_linear(mode, q, s, block_w=bw, block_h=bh)[0] == ref
            assert oracle.score_affine(mode, q, s, block_w=bw, block_h=bh)[0] == oracle.score_affine(mode, q, s)[0]
        for thr in (1, 2, 4, 8):
            assert oracle.score_linear(mode, q, s, threads=thr) == oracle.score_linear(mode, q, s, threads=1)


def test_affine_with_zero_gap_init_is_linear(oracle):
    """SURVEY.md A.7 regression: gi = 0 must reproduce the linear recurrence"""
    rng = np.random.default_rng(3)
    q, s = _rand(rng, 500), _rand(rng, 450)
    for mode in MODES:
        for gap in (-1, -3):
            assert oracle.score_affine(mode, q, s, 2, -1, 0, gap)[0] == oracle.score_linear(mode, q, s, 2, -1, gap)[0]


def test_reduce_max_lowest_index(oracle):
    """src/utils.impala:30-49: strict '>' everywhere => lowest index of the maximum"""
    rng = np.random.default_rng(5)
    for n in (1, 2, 63, 64, 65, 129, 1000, 4097):
        v = rng.integers(-5, 5, n + 1).astype(np.int32)        # element 0 is slot -1
        sc, ix = oracle.reduce_max(v, -1, n + 1)
        assert sc == v.max() and ix == int(np.argmax(v)) - 1
        sc, ix = oracle.reduce_max(v[1:], 0, n)
        assert sc == v[1:].max() and ix == int(np.argmax(v[1:]))


def test_next_pow_2(oracle):
    assert [oracle.next_pow_2(i) for i in (0, 1, 2, 3, 4, 5, 128, 129, 1000)] == [0, 1, 2, 4, 4, 8, 128, 256, 1024]


def _degap(a: bytes) -> bytes:
    return bytes(c for c in a if c not in (ord(" "), ord("_")))


def test_global_traceback_is_optimal_alignment(oracle):
    """structural invariants: de-gapped rows reproduce the inputs,
    the column score equals the optimal global score, rows have equal occupancy"""
    rng = np.random.default_rng(9)
    for (m, n) in ((50, 65), (300, 200), (129, 128), (1300, 2300), (900, 4100), (3000, 130)):
        q = _rand(rng, m)
        s = q.copy()[: min(m, n)]
        s = np.concatenate([s, _rand(rng, n - len(s))]) if len(s) < n else s
        s[rng.integers(0, n, n // 10)] = ord("A")
        ret, aq, as_, sp = oracle.traceback_lintime("global", q, s)
        assert ret == -m
        assert _degap(aq) == bytes(q) and _degap(as_) == bytes(s)
        assert oracle.column_score(aq, as_) == oracle.score_linear("global", q, s)[0]
        assert all((a == 32) == (b == 32) for a, b in zip(aq, as_))
        assert sp[0] == 0 and sp[-1] == m and all(sp[i] <= sp[i + 1] for i in range(len(sp) - 1))


def test_short_subject_quirk_q4(oracle):
    """n <= 64: block height 0 => all of s against gaps, none of q (global); nothing otherwise"""
    q = np.frombuffer(b"ACGTACGTAC", dtype=np.uint8)
    s = np.frombuffer(b"ACGTTGCA", dtype=np.uint8)
    ret, aq, as_, sp = oracle.traceback_lintime("global", q, s)
    assert aq[:8] == b"_" * 8 and as_[:8] == bytes(s) and set(aq[8:]) == {32}
    ret, aq, as_, sp = oracle.traceback_lintime("local", q, s)
    assert set(aq) == {32} and set(as_) == {32}


def test_golden_fixtures(oracle, golden):
    for c in golden["cases"]:
        q, s = c["q"].encode("latin-1"), c["s"].encode("latin-1")
        for mode in MODES:
            for key, exp in c["linear"][mode].items():
                sa, di, ga = map(int, key.split(","))
                assert list(oracle.score_linear(mode, q, s, sa, di, ga)) == exp, (c["name"], mode, key)
            for key, exp in c["affine"][mode].items():
                sa, di, gi, ge = map(int, key.split(","))
                assert oracle.score_affine(mode, q, s, sa, di, gi, ge)[0] == exp
            t = c["traceback"][mode]
            ret, aq, as_, sp = oracle.traceback_lintime(mode, q, s)
            assert ret == t["ret"] and sp.tolist() == t["splits"]
            assert hashlib.sha256(aq + b"\n" + as_).hexdigest()[:16] == t["sha"]
            ta = c["traceback_affine"][mode]
            ret, aq, as_, sp, ty = oracle.traceback_lintime_affine(mode, q, s, 2, -1, -2, -1)
            assert (ret, sp.tolist(), ty.tolist()) == (ta["ret"], ta["splits"], ta["types"])
            assert hashlib.sha256(aq + b"\n" + as_).hexdigest()[:16] == ta["sha"]
            if mode == "global" and len(s) > 64:        # subjects of <= 64 symbols: the reference's quirk Q4 output
                assert ta["column_score"] == c["affine"]["global"]["2,-1,-2,-1"]


def test_traceback_full_properties_and_fixtures(oracle, golden):
    """traceback_full (src/align.impala:190-216): one walk from get_score_pos() through the whole predecessor
    matrix.  For EVERY scheme the emitted columns score exactly the optimum (the linear-space path only
    guarantees that for global), de-gapped rows are substrings starting at get_alignment_start(), gap_init = 0
    of the Gotoh variant is the reference path, and the frozen fixtures hold."""
    rng = np.random.default_rng(77)
    ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
    for (m, n) in [(1, 1), (3, 40), (200, 150), (700, 1300), (1500, 40)]:
        q = ACGT[rng.integers(0, 4, m)]
        s = ACGT[rng.integers(0, 4, n)]
        k = min(m, n) // 2
        s[n - k:] = q[:k]                                     # overlap: suffix of s = prefix of q
        for mode in MODES:
            for (gi, ge) in [(0, -1), (-2, -1), (-5, -2)]:
                sc, aq, as_, st = oracle.traceback_full(mode, q, s, 2, -1, gi, ge)
                want = oracle.score_linear(mode, q, s, 2, -1, ge) if gi == 0 else oracle.score_affine(mode, q, s, 2, -1, gi, ge)
                assert sc == want[0]
                col = oracle.column_score(aq, as_, 2, -1, ge) if gi == 0 else oracle.column_score_affine(aq, as_, 2, -1, gi, ge)
                a = np.frombuffer(aq, np.uint8); b = np.frombuffer(as_, np.uint8)
                nonblank = int(((a != 32) | (b != 32)).sum())
                assert col == sc or (nonblank == 0 and sc <= 0), (m, n, mode, gi, col, sc)
                dq = bytes(a[(a != 32) & (a != 95)]); ds = bytes(b[(b != 32) & (b != 95)])
                assert bytes(q)[st[0]:st[0] + len(dq)] == dq and bytes(s)[st[1]:st[1] + len(ds)] == ds
                if nonblank:
                    assert (st[0] + len(dq) - 1, st[1] + len(ds) - 1) == tuple(want[1:])     # ends at get_score_pos()
                if mode == "global":
                    assert dq == bytes(q) and ds == bytes(s)
    for c in golden["cases"]:
        q, s = c["q"].encode("latin-1"), c["s"].encode("latin-1")
        for mode in MODES:
            t = c["traceback_full"][mode]
            sc, aq, as_, st = oracle.traceback_full(mode, q, s)
            assert (sc, list(st), hashlib.sha256(aq + b"\n" + as_).hexdigest()[:16]) == (t["score"], t["start"], t["sha"])
            sc, aq, as_, st = oracle.traceback_full(mode, q, s, 2, -1, -2, -1)
            assert (sc, list(st), hashlib.sha256(aq + b"\n" + as_).hexdigest()[:16]) == \
                   (t["affine"]["score"], t["affine"]["start"], t["affine"]["sha"])


def test_affine_traceback_is_optimal_global(oracle):
    """build-defined Gotoh traceback (parity unpinned vs the reference): for the global scheme the emitted
    alignment must be a valid optimal one -- de-gapped rows reproduce the inputs and the affine column score
    equals the independent textbook Gotoh optimum; joins inside horizontal gaps (type E) must occur"""
    rng = np.random.default_rng(17)
    etypes = 0
    for trial in range(25):
        m = int(rng.integers(200, 1800))
        q = _rand(rng, m)
        parts, p = [], 0
        while p < m:
            L = int(rng.integers(50, 200)); parts.append(q[p:p + L]); p += L
            parts.append(_rand(rng, int(rng.integers(5, 60))))
        s = np.concatenate(parts)
        for (sa, di, gi, ge) in ((2, -1, -6, -1), (5, -4, -10, -1)):
            ret, aq, as_, sp, ty = oracle.traceback_lintime_affine("global", q, s, sa, di, gi, ge)
            assert _degap(aq) == bytes(q) and _degap(as_) == bytes(s)
            assert oracle.column_score_affine(aq, as_, sa, di, gi, ge) == oracle.textbook_affine("global", q, s, sa, di, gi, ge)
            assert ret == gi + m * ge
            etypes += int(ty.sum())
    assert etypes > 0


def test_fullsize_check_equals_oracle(oracle):
    """oracle/fullsize_check.c (the vectorised second CPU implementation that freezes the full-size C2 result,
    tools/freeze_fullsize.py) against the scalar restatement: scores and end cells, all schemes, ragged shapes around
    its 2048 x 512 tiles, sequences with long gaps and non-ACGT bytes"""
    rng = np.random.default_rng(99)
    shapes = [(1, 1), (3, 2047), (511, 2048), (513, 2049), (700, 4097), (2500, 6000), (5000, 300), (1025, 10241)]
    for (m, n) in shapes:
        a = _rand(rng, m)
        b = _rand(rng, n)
        if m > 400:                      # related sequences with a long gap: E and F chains cross tile borders
            b = np.concatenate([a[: m // 2], _rand(rng, 300), a[m // 2:], b])[:n]
        for mode in MODES:
            for (gi, ge) in ((-2, -1), (0, -1), (-7, -3)):
                got = oracle.fullsize_score(mode, a, b, 2, -1, gi, ge, threads=3)
                want = oracle.score_affine(mode, a, b, 2, -1, gi, ge) if gi else oracle.score_linear(mode, a, b, 2, -1, ge)
                assert got[0] == want[0], (m, n, mode, gi, ge)
                if mode != "local":
                    assert got == want, (m, n, mode, gi, ge)
    raw = rng.integers(0, 256, 3000).astype(np.uint8)
    raw2 = np.concatenate([raw[100:2000], rng.integers(0, 256, 900).astype(np.uint8)])
    for mode in MODES:
        assert oracle.fullsize_score(mode, raw, raw2, 3, -2, -4, -1)[0] == oracle.score_affine(mode, raw, raw2, 3, -2, -4, -1)[0]
    # the frozen scaled-down C2 entry is reproduced by the scalar restatement
    import json, os
    from anyseq_b200 import workloads as W
    gold = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fullsize.json")))
    g = gold["c2_semiglobal_affine_scale_0.02"]
    q, s, _ = W.whole_genome_pair(0.02)
    assert "%016x" % oracle.fnv1a64(q) == g["fnv_q"] and "%016x" % oracle.fnv1a64(s) == g["fnv_s"]
    assert oracle.score_affine("semiglobal", q, s, threads=8) == (g["score"], g["end_i"], g["end_j"])
