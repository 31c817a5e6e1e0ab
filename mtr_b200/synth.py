"""Seeded synthetic read generators (numpy only; no GPU, no oracle).

``rand_seq_reads`` follows the model of the reference's test generator
(/root/reference/test_single_TR/util/rand_seq.cpp:48-222): a non-periodic random unit, ``copies``
copies of it with substitutions / insertions / deletions placed on distinct repeat positions, and random
flanks -- but seeded, so that tests and benches are reproducible (the reference seeds from random_device).

``long_reads`` is the C5 workload of SURVEY.md 8(d): reads of U[10 000, 20 000] bases with one tandem repeat
(unit length log-uniform in [2, 500]) covering 50-80 % of the read, per-read error rate U[5 %, 15 %] split
evenly between substitutions, insertions and deletions.

``standin_reads`` replaces the missing PacBio_Nanopore_read/*.fasta files.
"""
from __future__ import annotations

import numpy as np

_ALPHA = np.frombuffer(b"ACGT", dtype=np.uint8)


def _is_periodic(unit: np.ndarray) -> bool:
    n = len(unit)
    for p in range(1, n):
        if n % p == 0 and np.array_equal(np.tile(unit[:p], n // p), unit):
            return True
    return False


def random_unit(rng: np.random.Generator, unit_len: int) -> np.ndarray:
    while True:
        u = rng.integers(0, 4, unit_len, dtype=np.int8)
        if unit_len < 2 or not _is_periodic(u):
            return u


def mutate_repeat(rng: np.random.Generator, unit: np.ndarray, copies: int,
                  sub: float, ins: float, dele: float) -> np.ndarray:
    """`copies` copies of `unit` with round(len*rate) errors of each kind on distinct positions
    (rand_seq.cpp:55-57, 82-123, 176-213).  Rates are fractions (0.05 = 5 %)."""
    rep = np.tile(unit, copies)
    n = len(rep)
    n_sub, n_ins, n_del = int(round(n * sub)), int(round(n * ins)), int(round(n * dele))
    n_err = min(n, n_sub + n_ins + n_del)
    pos = rng.choice(n, n_err, replace=False)
    kind = np.zeros(n, dtype=np.int8)
    kind[pos[:n_sub]] = 1
    kind[pos[n_sub:n_sub + n_ins]] = 2
    kind[pos[n_sub + n_ins:]] = 3
    sub_shift = rng.integers(1, 4, n, dtype=np.int8)
    ins_base = rng.integers(0, 4, n, dtype=np.int8)
    counts = np.ones(n, dtype=np.int64)
    counts[kind == 2] = 2
    counts[kind == 3] = 0
    out = np.repeat(rep, counts)
    start = np.cumsum(counts) - counts
    sub = kind == 1
    out[start[sub]] = (rep[sub] + sub_shift[sub]) % 4
    ins = kind == 2
    out[start[ins] + 1] = ins_base[ins]
    return out.astype(np.int8)


def rand_seq_reads(unit_len: int, copies: int, sub: float, ins: float, dele: float,
                   pre: int, post: int, n_reads: int, seed: int):
    """Returns (reads, units): lists of int8 arrays with values 0..3 (A, C, G, T)."""
    rng = np.random.default_rng(seed)
    reads, units = [], []
    for _ in range(n_reads):
        unit = random_unit(rng, unit_len)
        body = mutate_repeat(rng, unit, copies, sub, ins, dele)
        read = np.concatenate([rng.integers(0, 4, pre, dtype=np.int8), body,
                               rng.integers(0, 4, post, dtype=np.int8)])
        reads.append(read); units.append(unit)
    return reads, units


def long_reads(n_reads: int, seed: int, len_lo: int = 10000, len_hi: int = 20000,
               unit_lo: int = 2, unit_hi: int = 500, err_lo: float = 0.05, err_hi: float = 0.15):
    """C5 workload (BASELINE.json configs[4])."""
    rng = np.random.default_rng(seed)
    reads, units = [], []
    for _ in range(n_reads):
        L = int(rng.integers(len_lo, len_hi + 1))
        ulen = int(round(np.exp(rng.uniform(np.log(unit_lo), np.log(unit_hi)))))
        ulen = max(unit_lo, min(unit_hi, ulen))
        frac = rng.uniform(0.5, 0.8)
        copies = max(6, int(L * frac) // ulen)
        err = rng.uniform(err_lo, err_hi) / 3.0
        unit = random_unit(rng, ulen)
        body = mutate_repeat(rng, unit, copies, err, err, err)
        if len(body) > L - 20:
            body = body[:L - 20]
        pre = int(rng.integers(0, L - len(body) + 1))
        post = L - len(body) - pre
        read = np.concatenate([rng.integers(0, 4, pre, dtype=np.int8), body,
                               rng.integers(0, 4, post, dtype=np.int8)])
        reads.append(read); units.append(unit)
    return reads, units


def standin_reads(kind: str, n_reads: int, seed: int):
    """Stand-ins for PacBio_Nanopore_read/100_100_nanopore.fasta ('nanopore': unit 100 x 100 copies,
    4 % sub / 4 % ins / 5 % del, 3 kb flanks) and 200_200_pacbio_1.fasta ('pacbio': unit 200 x 200 copies,
    1.5 % sub / 6 % ins / 4 % del, 2 kb flanks); see PacBio_Nanopore_read/Readme:3-11."""
    if kind == "nanopore":
        return rand_seq_reads(100, 100, 0.04, 0.04, 0.05, 3000, 3000, n_reads, seed)
    if kind == "pacbio":
        return rand_seq_reads(200, 200, 0.015, 0.06, 0.04, 2000, 2000, n_reads, seed)
    raise ValueError(kind)


def to_text(read: np.ndarray) -> str:
    return _ALPHA[np.asarray(read, dtype=np.int64)].tobytes().decode()


def write_fasta(path: str, reads, ids=None, line_width: int = 0) -> None:
    """One read per line like rand_seq.cpp (line_width=0), or wrapped."""
    with open(path, "w") as f:
        for i, r in enumerate(reads):
            f.write(">%s\n" % (ids[i] if ids is not None else i))
            s = to_text(r)
            if line_width:
                for j in range(0, len(s), line_width):
                    f.write(s[j:j + line_width] + "\n")
            else:
                f.write(s + "\n")
