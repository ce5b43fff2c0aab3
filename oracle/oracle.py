"""ctypes binding of the CPU parity oracle -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module (see oracle/anyseq_oracle.h).  The product package
``anyseq_b200`` never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")

GLOBAL, SEMIGLOBAL, LOCAL = 0, 1, 2
MODES = {"global": GLOBAL, "semiglobal": SEMIGLOBAL, "local": LOCAL}
SCORE_MIN = -2147483647


class _Result(C.Structure):
    _fields_ = [("score", C.c_int32), ("pos_i", C.c_int32), ("pos_j", C.c_int32)]


def build(force: bool = False) -> str:
    """Compile oracle/liboracle.so with the committed Makefile (gcc/g++ only)."""
    srcs = [os.path.join(_HERE, f) for f in ("anyseq_oracle.c", "anyseq_oracle.h", "refinput.cpp", "fullsize_check.c", "Makefile")]
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs)
    if force or stale:
        subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
    return _LIB_PATH


REF_HOST = os.path.join(_HERE, "_ref", "align_reference_host")


def build_reference_host(reference_root: str = "/root/reference"):
    """The reference's own host program (src/main.cpp) linked against libanyseq_b200.so; returns its path or
    None when the reference sources are not on this machine (GPU box: the prebuilt binary travels)."""
    if os.path.exists(os.path.join(reference_root, "src", "main.cpp")):
        subprocess.run(["make", "-C", _HERE, "ref_host", f"REF={reference_root}"], check=True, capture_output=True)
    return REF_HOST if os.path.exists(REF_HOST) else None


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        build()
    L = C.CDLL(_LIB_PATH)
    u8p = C.POINTER(C.c_uint8)
    i32p = C.POINTER(C.c_int32)
    L.oracle_score_linear.restype = _Result
    L.oracle_score_linear.argtypes = [C.c_int, u8p, C.c_int, u8p, C.c_int,
                                      C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    L.oracle_score_affine.restype = _Result
    L.oracle_score_affine.argtypes = [C.c_int, u8p, C.c_int, u8p, C.c_int,
                                      C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    L.fullsize_score.restype = _Result
    L.fullsize_score.argtypes = [C.c_int, u8p, C.c_int, u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    L.textbook_score_linear.restype = C.c_int32
    L.textbook_score_linear.argtypes = [C.c_int, u8p, C.c_int, u8p, C.c_int, C.c_int, C.c_int, C.c_int]
    L.textbook_score_affine.restype = C.c_int32
    L.textbook_score_affine.argtypes = [C.c_int, u8p, C.c_int, u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    L.oracle_traceback_lintime.restype = C.c_int32
    L.oracle_traceback_lintime.argtypes = [C.c_int, u8p, C.c_int, u8p, C.c_int, C.c_int, C.c_int, C.c_int,
                                           u8p, u8p, i32p, C.c_int]
    L.oracle_traceback_lintime_affine.restype = C.c_int32
    L.oracle_traceback_lintime_affine.argtypes = [C.c_int, u8p, C.c_int, u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                                  u8p, u8p, i32p, i32p, C.c_int]
    L.oracle_traceback_full.restype = C.c_int32
    L.oracle_traceback_full.argtypes = [C.c_int, u8p, C.c_int, u8p, C.c_int, C.c_int, C.c_int, C.c_int,
                                        u8p, u8p, i32p, C.c_int]
    L.oracle_traceback_full_affine.restype = C.c_int32
    L.oracle_traceback_full_affine.argtypes = [C.c_int, u8p, C.c_int, u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                               u8p, u8p, i32p, C.c_int]
    L.oracle_alignment_column_score_affine.restype = C.c_int64
    L.oracle_alignment_column_score_affine.argtypes = [u8p, u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    L.oracle_reduce_max.restype = None
    L.oracle_reduce_max.argtypes = [i32p, C.c_int, C.c_int, i32p, i32p]
    L.oracle_next_pow_2.restype = C.c_int32
    L.oracle_next_pow_2.argtypes = [C.c_int32]
    L.oracle_alignment_column_score.restype = C.c_int64
    L.oracle_alignment_column_score.argtypes = [u8p, u8p, C.c_int, C.c_int, C.c_int, C.c_int]
    L.oracle_reference_random_pair.restype = None
    L.oracle_reference_random_pair.argtypes = [C.c_int64, C.c_int64, u8p, C.POINTER(C.c_int), u8p, C.POINTER(C.c_int)]
    L.oracle_score_batch.restype = None
    L.oracle_score_batch.argtypes = [C.c_int, u8p, C.POINTER(C.c_int64), u8p, C.POINTER(C.c_int64), C.c_int64,
                                     C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, i32p]
    L.oracle_fnv1a64.restype = C.c_uint64
    L.oracle_fnv1a64.argtypes = [u8p, C.c_int64]
    _lib = L
    return L


def _u8(a) -> np.ndarray:
    if isinstance(a, (bytes, bytearray)):
        a = np.frombuffer(bytes(a), dtype=np.uint8)
    elif isinstance(a, str):
        a = np.frombuffer(a.encode("latin-1"), dtype=np.uint8)
    a = np.ascontiguousarray(a, dtype=np.uint8)
    if a.size == 0:           # keep a valid pointer for empty inputs
        a = np.zeros(1, dtype=np.uint8)[:0]
    return a


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_uint8))


def _mode(mode) -> int:
    return MODES[mode] if isinstance(mode, str) else int(mode)


def score_linear(mode, q, s, same=2, diff=-1, gap=-1, threads=4, block_w=0, block_h=0):
    """(score, pos_i, pos_j) of the restated reference CPU path."""
    q, s = _u8(q), _u8(s)
    r = lib().oracle_score_linear(_mode(mode), _ptr(q), len(q), _ptr(s), len(s),
                                  same, diff, gap, threads, block_w, block_h)
    return r.score, r.pos_i, r.pos_j


def score_affine(mode, q, s, same=2, diff=-1, gap_init=-2, gap_extend=-1, threads=4, block_w=0, block_h=0):
    q, s = _u8(q), _u8(s)
    r = lib().oracle_score_affine(_mode(mode), _ptr(q), len(q), _ptr(s), len(s),
                                  same, diff, gap_init, gap_extend, threads, block_w, block_h)
    return r.score, r.pos_i, r.pos_j


def fullsize_score(mode, q, s, same=2, diff=-1, gap_init=-2, gap_extend=-1, threads=4):
    """(score, pos_i, pos_j) of the vectorised second CPU implementation (oracle/fullsize_check.c): the same
    recurrences and result extraction as score_affine / score_linear (gap_init = 0), fast enough for 4.6 Mbp^2.
    Local: score only (pos = -1)."""
    q, s = _u8(q), _u8(s)
    r = lib().fullsize_score(_mode(mode), _ptr(q), len(q), _ptr(s), len(s), same, diff, gap_init, gap_extend, threads)
    return r.score, r.pos_i, r.pos_j


def score_batch(mode, q, q_off, s, s_off, same=2, diff=-1, gap_init=-2, gap_extend=-1, threads=4) -> np.ndarray:
    """scores of a batch of independent pairs (packed sequences + offsets), pair by pair through the restated path"""
    q, s = _u8(q), _u8(s)
    qo = np.ascontiguousarray(q_off, dtype=np.int64)
    so = np.ascontiguousarray(s_off, dtype=np.int64)
    n = len(qo) - 1
    out = np.zeros(max(n, 1), dtype=np.int32)
    i64p = C.POINTER(C.c_int64)
    lib().oracle_score_batch(_mode(mode), _ptr(q), qo.ctypes.data_as(i64p), _ptr(s), so.ctypes.data_as(i64p), n,
                             same, diff, gap_init, gap_extend, threads, out.ctypes.data_as(C.POINTER(C.c_int32)))
    return out[:n]


def textbook_linear(mode, q, s, same=2, diff=-1, gap=-1) -> int:
    q, s = _u8(q), _u8(s)
    return lib().textbook_score_linear(_mode(mode), _ptr(q), len(q), _ptr(s), len(s), same, diff, gap)


def textbook_affine(mode, q, s, same=2, diff=-1, gap_init=-2, gap_extend=-1) -> int:
    q, s = _u8(q), _u8(s)
    return lib().textbook_score_affine(_mode(mode), _ptr(q), len(q), _ptr(s), len(s),
                                       same, diff, gap_init, gap_extend)


def traceback_lintime(mode, q, s, same=2, diff=-1, gap=-1, threads=4):
    """Returns (ret, aligned_q bytes, aligned_s bytes, splits ndarray incl. the -1 slot)."""
    q, s = _u8(q), _u8(s)
    m, n = len(q), len(s)
    oq = np.zeros(max(m + n, 1), dtype=np.uint8)
    os_ = np.zeros(max(m + n, 1), dtype=np.uint8)
    nb = (n + 127) // 128
    splits = np.zeros(nb + 1, dtype=np.int32)
    ret = lib().oracle_traceback_lintime(_mode(mode), _ptr(q), m, _ptr(s), n, same, diff, gap,
                                         _ptr(oq), _ptr(os_),
                                         splits.ctypes.data_as(C.POINTER(C.c_int32)), threads)
    return ret, oq[:m + n].tobytes(), os_[:m + n].tobytes(), splits


def traceback_lintime_affine(mode, q, s, same=2, diff=-1, gap_init=-2, gap_extend=-1, threads=4):
    """Returns (ret, aligned_q, aligned_s, splits, vertex types).  Build-defined Gotoh traceback."""
    q, s = _u8(q), _u8(s)
    m, n = len(q), len(s)
    oq = np.zeros(max(m + n, 1), dtype=np.uint8)
    os_ = np.zeros(max(m + n, 1), dtype=np.uint8)
    nb = (n + 127) // 128
    splits = np.zeros(nb + 1, dtype=np.int32)
    types = np.zeros(nb + 1, dtype=np.int32)
    i32p = C.POINTER(C.c_int32)
    ret = lib().oracle_traceback_lintime_affine(_mode(mode), _ptr(q), m, _ptr(s), n, same, diff, gap_init, gap_extend,
                                                _ptr(oq), _ptr(os_), splits.ctypes.data_as(i32p),
                                                types.ctypes.data_as(i32p), threads)
    return ret, oq[:m + n].tobytes(), os_[:m + n].tobytes(), splits, types


def traceback_full(mode, q, s, same=2, diff=-1, gap_init=0, gap_extend=-1, threads=4):
    """traceback_full of the reference (gap_init == 0) or the build-defined Gotoh variant.
    Returns (score, aligned_q, aligned_s, (start_i, start_j))."""
    q, s = _u8(q), _u8(s)
    m, n = len(q), len(s)
    oq = np.zeros(max(m + n, 1), dtype=np.uint8)
    os_ = np.zeros(max(m + n, 1), dtype=np.uint8)
    start = np.zeros(2, dtype=np.int32)
    sp = start.ctypes.data_as(C.POINTER(C.c_int32))
    if gap_init == 0:
        ret = lib().oracle_traceback_full(_mode(mode), _ptr(q), m, _ptr(s), n, same, diff, gap_extend,
                                          _ptr(oq), _ptr(os_), sp, threads)
    else:
        ret = lib().oracle_traceback_full_affine(_mode(mode), _ptr(q), m, _ptr(s), n, same, diff, gap_init, gap_extend,
                                                 _ptr(oq), _ptr(os_), sp, threads)
    return ret, oq[:m + n].tobytes(), os_[:m + n].tobytes(), (int(start[0]), int(start[1]))


def column_score_affine(aq: bytes, as_: bytes, same=2, diff=-1, gap_init=-2, gap_extend=-1) -> int:
    a, b = _u8(aq), _u8(as_)
    return lib().oracle_alignment_column_score_affine(_ptr(a), _ptr(b), len(a), same, diff, gap_init, gap_extend)


def reduce_max(vec, offset, length):
    """vec: int32 array whose element 0 is logical index -1 when offset == -1."""
    v = np.ascontiguousarray(vec, dtype=np.int32)
    sc, ix = C.c_int32(), C.c_int32()
    base = v.ctypes.data + (4 if offset < 0 else 0)
    lib().oracle_reduce_max(C.cast(base, C.POINTER(C.c_int32)), offset, length, C.byref(sc), C.byref(ix))
    return sc.value, ix.value


def next_pow_2(i: int) -> int:
    return lib().oracle_next_pow_2(i)


def column_score(aq: bytes, as_: bytes, same=2, diff=-1, gap=-1) -> int:
    a, b = _u8(aq), _u8(as_)
    return lib().oracle_alignment_column_score(_ptr(a), _ptr(b), len(a), same, diff, gap)


def reference_random_pair(minlen=256, maxlen=1024):
    """The (query, subject) `align -r [min [max]]` of the reference generates."""
    hi = max(minlen, maxlen)
    q = np.zeros(hi, dtype=np.uint8)
    s = np.zeros(hi, dtype=np.uint8)
    m, n = C.c_int(), C.c_int()
    lib().oracle_reference_random_pair(minlen, maxlen, _ptr(q), C.byref(m), _ptr(s), C.byref(n))
    return q[:m.value].copy(), s[:n.value].copy()


def fnv1a64(a) -> int:
    a = _u8(a)
    return lib().oracle_fnv1a64(_ptr(a), len(a))
