"""Fast synthetic workloads for bench.py / tools (vectorised numpy, deterministic per seed).

The parity tests use ghostm_b200.synth (sequence lists, FASTA-able); this module generates the
BASELINE.json-sized inputs directly as chunk arrays: a db chunk is a residue array with
SEQUENCE_END separators plus its .pos table, exactly what `ghostm db` writes (db_creator.cpp:130-165),
and its index is built on the device (gm_db_build_index).
"""
from __future__ import annotations

import numpy as np

from .formats import SEQUENCE_END
from .synth import ROBINSON

_BLOSUM62_TEXT = """   A  R  N  D  C  Q  E  G  H  I  L  K  M  F  P  S  T  W  Y  V  B  Z  X  *
A  4 -1 -2 -2  0 -1 -1  0 -2 -1 -1 -1 -1 -2 -1  1  0 -3 -2  0 -2 -1  0 -4
R -1  5  0 -2 -3  1  0 -2  0 -3 -2  2 -1 -3 -2 -1 -1 -3 -2 -3 -1  0 -1 -4
N -2  0  6  1 -3  0  0  0  1 -3 -3  0 -2 -3 -2  1  0 -4 -2 -3  3  0 -1 -4
D -2 -2  1  6 -3  0  2 -1 -1 -3 -4 -1 -3 -3 -1  0 -1 -4 -3 -3  4  1 -1 -4
C  0 -3 -3 -3  9 -3 -4 -3 -3 -1 -1 -3 -1 -2 -3 -1 -1 -2 -2 -1 -3 -3 -2 -4
Q -1  1  0  0 -3  5  2 -2  0 -3 -2  1  0 -3 -1  0 -1 -2 -1 -2  0  3 -1 -4
E -1  0  0  2 -4  2  5 -2  0 -3 -3  1 -2 -3 -1  0 -1 -3 -2 -2  1  4 -1 -4
G  0 -2  0 -1 -3 -2 -2  6 -2 -4 -4 -2 -3 -3 -2  0 -2 -2 -3 -3 -1 -2 -1 -4
H -2  0  1 -1 -3  0  0 -2  8 -3 -3 -1 -2 -1 -2 -1 -2 -2  2 -3  0  0 -1 -4
I -1 -3 -3 -3 -1 -3 -3 -4 -3  4  2 -3  1  0 -3 -2 -1 -3 -1  3 -3 -3 -1 -4
L -1 -2 -3 -4 -1 -2 -3 -4 -3  2  4 -2  2  0 -3 -2 -1 -2 -1  1 -4 -3 -1 -4
K -1  2  0 -1 -3  1  1 -2 -1 -3 -2  5 -1 -3 -1  0 -1 -3 -2 -2  0  1 -1 -4
M -1 -1 -2 -3 -1  0 -2 -3 -2  1  2 -1  5  0 -2 -1 -1 -1 -1  1 -3 -1 -1 -4
F -2 -3 -3 -3 -2 -3 -3 -3 -1  0  0 -3  0  6 -4 -2 -2  1  3 -1 -3 -3 -1 -4
P -1 -2 -2 -1 -3 -1 -1 -2 -2 -3 -3 -1 -2 -4  7 -1 -1 -4 -3 -2 -2 -1 -2 -4
S  1 -1  1  0 -1  0  0  0 -1 -2 -2  0 -1 -2 -1  4  1 -3 -2 -2  0  0  0 -4
T  0 -1  0 -1 -1 -1 -1 -2 -2 -1 -1 -1 -1 -2 -1  1  5 -2 -2  0 -1 -1  0 -4
W -3 -3 -4 -4 -2 -2 -3 -2 -2 -3 -2 -3 -1  1 -4 -3 -2 11  2 -3 -4 -3 -2 -4
Y -2 -2 -2 -3 -2 -1 -2 -3  2 -1 -1 -2 -1  3 -3 -2 -2  2  7 -1 -3 -2 -1 -4
V  0 -3 -3 -3 -1 -2 -2 -3 -3  3  1 -2  1 -1 -2 -2  0 -3 -1  4 -3 -2 -1 -4
B -2 -1  3  4 -3  0  1 -1  0 -3 -4  0 -3 -3 -2  0 -1 -4 -3 -3  4  1 -1 -4
Z -1  0  0  1 -3  3  4 -2  0 -3 -3  1 -1 -3 -1  0 -1 -3 -2 -2  1  4 -1 -4
X  0 -1 -1 -1 -2 -1 -1 -1 -1 -1 -1 -1 -1 -1 -2  0  0 -2 -1 -1 -1 -1 -1 -4
* -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4  1
"""


def blosum62() -> np.ndarray:
    """The NCBI BLOSUM62 table as the 32x32 int matrix `ghostm aln` builds by default
    (score_matrix_reader.cpp:80-113: row = first residue code, column = second)."""
    from .formats import PROTEIN_TO_CODE
    lines = [ln.split() for ln in _BLOSUM62_TEXT.strip("\n").split("\n")]
    cols = lines[0]
    m = np.zeros((32, 32), dtype=np.int32)
    for row in lines[1:]:
        r = PROTEIN_TO_CODE[ord(row[0])]
        for cname, v in zip(cols, row[1:]):
            m[r, PROTEIN_TO_CODE[ord(cname)]] = int(v)
    return m.reshape(-1)


def _residues(rng: np.random.Generator, n: int) -> np.ndarray:
    cdf = np.cumsum(ROBINSON)
    cdf[-1] = 1.0
    return np.searchsorted(cdf, rng.random(n, dtype=np.float32), side="right").astype(np.uint8).clip(0, 19)


def synth_chunk(seed: int, chunk: int, chunk_bytes: int, min_len: int = 100, max_len: int = 600,
                repeats: bool = False):
    """One db chunk of at most chunk_bytes (residues + separators): sequences of length
    U[min_len, max_len], i.i.d. Robinson residues -> (seq uint8[], seq_starts uint32[])."""
    rng = np.random.default_rng([seed, chunk])
    n_est = chunk_bytes // ((min_len + max_len) // 2 + 1) + 16
    lens = rng.integers(min_len, max_len + 1, size=n_est)
    ends = np.cumsum(lens + 1)
    n = int(np.searchsorted(ends, chunk_bytes, side="right"))
    lens, ends = lens[:n], ends[:n]
    total = int(ends[-1])
    seq = _residues(rng, total)
    if repeats:   # C5: 30 % of each sequence is a tandem repeat of period 1-4 over {A,G,S,L}
        alphabet = np.array([0, 7, 15, 10], dtype=np.uint8)
        starts0 = ends - lens - 1
        for s0, ln in zip(starts0, lens):
            rl = int(ln * 0.3)
            period = int(rng.integers(1, 5))
            unit = alphabet[rng.integers(0, 4, size=period)]
            st = int(s0 + rng.integers(0, ln - rl + 1))
            seq[st:st + rl] = np.resize(unit, rl)
    seq[ends - 1] = SEQUENCE_END
    starts = (ends - lens - 1).astype(np.uint32)
    return seq, starts


def synth_queries(seed: int, source: np.ndarray, n: int, length: int, sub_rate: float = 0.15,
                  frac_db: float = 0.5) -> np.ndarray:
    """n x length queries: frac_db of them are substrings of `source` (a db prefix) with
    sub_rate substitutions, the rest random; separators inside a substring become random
    residues."""
    rng = np.random.default_rng([seed, n, length])
    q = _residues(rng, n * length).reshape(n, length)
    from_db = rng.random(n) < frac_db
    st = rng.integers(0, max(source.shape[0] - length, 1), size=n)
    idx = st[:, None] + np.arange(length)[None, :]
    sub = source[idx]
    keep = (rng.random((n, length)) >= sub_rate) & (sub != SEQUENCE_END) & from_db[:, None]
    q[keep] = sub[keep]
    return np.ascontiguousarray(q)
