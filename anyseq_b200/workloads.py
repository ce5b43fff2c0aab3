"""Synthetic workloads of BASELINE.json (the bundled genomes are missing from the
reference mount: /root/reference/.MISSING_LARGE_BLOBS).  Deterministic numpy
generators; real FASTA files dropped into sequences/ are used when present.
"""
from __future__ import annotations

import os

import numpy as np

ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def random_dna(n: int, seed: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    return ACGT[rng.integers(0, 4, size=n, dtype=np.uint8)]


def mutated_copy(src: np.ndarray, n_out: int, seed: int, sub_rate=0.02, indel_rate=0.001, indel_mean=3.0):
    """copy of `src` with substitutions and geometric-length indels, cut / padded to n_out"""
    rng = np.random.default_rng(seed)
    s = src.copy()
    sub = rng.random(len(s)) < sub_rate
    s[sub] = ACGT[rng.integers(0, 4, size=int(sub.sum()), dtype=np.uint8)]
    ev = np.flatnonzero(rng.random(len(s)) < indel_rate)
    lens = rng.geometric(1.0 / indel_mean, size=len(ev))
    is_del = rng.random(len(ev)) < 0.5
    keep = np.ones(len(s), dtype=bool)
    for p, L in zip(ev[is_del], lens[is_del]):
        keep[p:p + L] = False
    ins_pos = ev[~is_del]
    ins_len = lens[~is_del]
    ins_idx = np.repeat(ins_pos, ins_len)
    ins_val = ACGT[rng.integers(0, 4, size=len(ins_idx), dtype=np.uint8)]
    keep_ins = np.insert(keep, ins_idx, True)
    s = np.insert(s, ins_idx, ins_val)[keep_ins]
    if len(s) >= n_out:
        return np.ascontiguousarray(s[:n_out])
    return np.concatenate([s, ACGT[rng.integers(0, 4, size=n_out - len(s), dtype=np.uint8)]])


def _read_fasta_first(path: str) -> np.ndarray:
    chunks = []
    with open(path, "rb") as f:
        header = f.readline()
        assert header[:1] == b">", "not a FASTA file"
        for line in f:
            if line[:1] == b">":
                break
            chunks.append(line.rstrip(b"\n"))
    return np.frombuffer(b"".join(chunks), dtype=np.uint8)


def whole_genome_pair(scale: float = 1.0):
    """C2 / C5: ecoli x sboydii.  Query = sboydii, subject = ecoli in the reference's
    own benchmark (benchmark.sh:7); synthetic stand-ins of SURVEY.md 8(d) unless the
    real files exist under sequences/.  Returns (query, subject, description)."""
    eco = os.path.join(_ROOT, "sequences", "ecoli.fna")
    sbo = os.path.join(_ROOT, "sequences", "sboydii.fna")
    if scale == 1.0 and os.path.exists(eco) and os.path.exists(sbo):
        return _read_fasta_first(sbo), _read_fasta_first(eco), "sequences/sboydii.fna x sequences/ecoli.fna"
    m = int(4_641_652 * scale)
    n = int(4_600_000 * scale)
    q = random_dna(m, 42)
    s = mutated_copy(q, n, 43)
    return q, s, f"synthetic ecoli-like {m} x mutated copy {n} (seeds 42/43)"


def random_pair(m: int, n: int, seed_q: int = 1, seed_s: int = 2):
    """C3: uniform random ACGT pair"""
    return random_dna(m, seed_q), random_dna(n, seed_s)


def read_batch(npairs: int, read_len: int = 150, window: int = 500, seed: int = 7):
    """C4: reads of `read_len` vs windows of `window` containing a mutated copy
    (5 % substitutions, 1 % indels) at a random offset.  Packed + offsets."""
    rng = np.random.default_rng(seed)
    reads = ACGT[rng.integers(0, 4, size=(npairs, read_len), dtype=np.uint8)]
    windows = ACGT[rng.integers(0, 4, size=(npairs, window), dtype=np.uint8)]
    off = rng.integers(0, window - read_len + 1, size=npairs)
    cols = off[:, None] + np.arange(read_len)[None, :]
    mut = reads.copy()
    sub = rng.random(mut.shape) < 0.05
    mut[sub] = ACGT[rng.integers(0, 4, size=int(sub.sum()), dtype=np.uint8)]
    # indels as single-base shifts inside the implanted copy (keeps the window length)
    sh = rng.random(mut.shape) < 0.01
    mut = np.where(sh, np.roll(mut, 1, axis=1), mut)
    np.put_along_axis(windows, cols, mut, axis=1)
    q_off = np.arange(npairs + 1, dtype=np.int64) * read_len
    s_off = np.arange(npairs + 1, dtype=np.int64) * window
    return reads.reshape(-1), q_off, windows.reshape(-1), s_off
