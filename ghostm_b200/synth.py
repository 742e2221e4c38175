"""Seeded synthetic inputs for the BASELINE.json configs.

Everything is numpy on the host and deterministic in (seed, sizes).  Residues are
i.i.d. from the Robinson frequencies the reference ships (statistics.cpp:72-93)
unless a generator says otherwise.  Sizes are parameters so that the same generator
serves the seconds-sized parity tests and the full-size bench workloads.
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np

from . import formats

# statistics.cpp:72-93, by residue code 0..19
ROBINSON = np.zeros(20)
for _code, _p in {10: 90.19, 0: 78.05, 7: 73.77, 15: 71.20, 19: 64.41, 6: 62.95, 16: 58.41,
                  11: 57.44, 3: 53.64, 14: 52.03, 9: 51.42, 1: 51.29, 2: 44.87, 5: 42.64,
                  13: 38.56, 18: 32.16, 12: 22.43, 8: 21.99, 4: 19.25, 17: 13.30}.items():
    ROBINSON[_code] = _p
ROBINSON /= ROBINSON.sum()


def random_residues(rng: np.random.Generator, n: int) -> np.ndarray:
    return rng.choice(20, size=n, p=ROBINSON).astype(np.uint8)


def protein_db(seed: int, n_residues: int, min_len: int = 100, max_len: int = 600
               ) -> Tuple[List[np.ndarray], List[str]]:
    """DB sequences of length U[min_len, max_len], names s<i>."""
    rng = np.random.default_rng(seed)
    seqs, names, total = [], [], 0
    while total < n_residues:
        ln = int(rng.integers(min_len, max_len + 1))
        ln = min(ln, max(n_residues - total, 1))
        seqs.append(random_residues(rng, ln))
        names.append(f"s{len(names)}")
        total += ln
    return seqs, names


def mutate(rng: np.random.Generator, s: np.ndarray, sub_rate: float) -> np.ndarray:
    s = s.copy()
    m = rng.random(s.shape[0]) < sub_rate
    s[m] = random_residues(rng, int(m.sum()))
    return s


def queries_from_db(seed: int, db_seqs: List[np.ndarray], n: int, length: int,
                    sub_rate: float = 0.15, frac_db: float = 0.5, min_length: int = 0,
                    name_prefix: str = "q", group: int = 1
                    ) -> Tuple[List[np.ndarray], List[str]]:
    """Half (frac_db) mutated db substrings, the rest random (C3/C4 of SURVEY §8d).

    min_length > 0 draws each query length from U[min_length, length] (C4).  group > 1
    gives `group` consecutive queries the same name, which is what the 6 translated
    frames of one DNA read look like to Aligner::Merge (aligner.cpp:697-742).
    """
    rng = np.random.default_rng(seed)
    concat = np.concatenate(db_seqs)
    seqs, names = [], []
    for i in range(n):
        ln = int(rng.integers(min_length, length + 1)) if min_length else length
        if rng.random() < frac_db and concat.shape[0] > ln:
            st = int(rng.integers(0, concat.shape[0] - ln))
            s = mutate(rng, concat[st:st + ln], sub_rate)
        else:
            s = random_residues(rng, ln)
        seqs.append(s)
        names.append(f"{name_prefix}{i // group}")
    return seqs, names


def repeat_db(seed: int, n_residues: int, repeat_frac: float = 0.3, min_len: int = 100,
              max_len: int = 600) -> Tuple[List[np.ndarray], List[str]]:
    """C5: 30 % of each sequence is tandem repeats of period 1-4 over {A,G,S,L}."""
    rng = np.random.default_rng(seed)
    alphabet = np.array([0, 7, 15, 10], dtype=np.uint8)
    seqs, names, total = [], [], 0
    while total < n_residues:
        ln = int(rng.integers(min_len, max_len + 1))
        s = random_residues(rng, ln)
        rl = int(ln * repeat_frac)
        period = int(rng.integers(1, 5))
        unit = alphabet[rng.integers(0, 4, size=period)]
        st = int(rng.integers(0, ln - rl + 1))
        s[st:st + rl] = np.resize(unit, rl)
        seqs.append(s)
        names.append(f"s{len(names)}")
        total += ln
    return seqs, names


def repeat_queries(seed: int, n: int, length: int) -> Tuple[List[np.ndarray], List[str]]:
    """C5 queries: random background carrying a tandem repeat over the same alphabet."""
    rng = np.random.default_rng(seed)
    alphabet = np.array([0, 7, 15, 10], dtype=np.uint8)
    seqs, names = [], []
    for i in range(n):
        s = random_residues(rng, length)
        rl = int(rng.integers(length // 4, length // 2 + 1))
        period = int(rng.integers(1, 5))
        unit = alphabet[rng.integers(0, 4, size=period)]
        st = int(rng.integers(0, length - rl + 1))
        s[st:st + rl] = np.resize(unit, rl)
        seqs.append(s)
        names.append(f"q{i}")
    return seqs, names


def workload(name: str, scale: float = 1.0):
    """Named workloads -> (Db, [QueryChunk]).  `scale` shrinks residue/query counts for
    tests; scale=1 is the BASELINE.json size."""
    if name == "c3":      # 1 M x 75 aa vs 1 G residues, db -l 128
        dbs, dbn = protein_db(1, int(1_000_000_000 * scale))
        q, qn = queries_from_db(2, dbs, max(int(1_000_000 * scale), 16), 75)
        return (formats.make_db(dbs, dbn, 4, 128), formats.make_query_chunks(q, qn, 75, 128))
    if name == "c5":
        dbs, dbn = repeat_db(5, int(64_000_000 * scale))
        q, qn = repeat_queries(6, max(int(10_000 * scale), 16), 75)
        return (formats.make_db(dbs, dbn, 4, 128), formats.make_query_chunks(q, qn, 75, 128))
    raise ValueError(name)
