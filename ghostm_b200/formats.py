"""On-disk formats of GHOSTM (frozen by the reference) as numpy readers/writers.

The files are little-endian raw structs without padding; layouts follow the
reference writers/readers cited on every function (paths under /root/reference).
The product's C++ host reads the same files (csrc/host/formats.h); this module
exists for the test-suite, the synthetic-data generators and bench.py, so that
fixtures can be produced on a box that has neither the reference nor FASTA input.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field
from typing import List, Sequence

import numpy as np

ALPHABET_SIZE = 32   # common.h:31
CHARACTER_SIZE = 5   # common.h:32
SEQUENCE_END = 25    # common.h:34
BASE_X = 23          # common.h:35

# sequence.cpp:63-87 (protein): A0 R1 N2 D3 C4 Q5 E6 G7 H8 I9 L10 K11 M12 F13 P14 S15
# T16 W17 Y18 V19 B20 J21 Z22 X23 *24; anything else -> X.
_LETTERS = "ARNDCQEGHILKMFPSTWYVBJZX*"
PROTEIN_TO_CODE = np.full(256, BASE_X, dtype=np.uint8)
for _i, _c in enumerate(_LETTERS):
    PROTEIN_TO_CODE[ord(_c)] = _i
    PROTEIN_TO_CODE[ord(_c.lower())] = _i
CODE_TO_PROTEIN = np.frombuffer((_LETTERS + "#" * (256 - len(_LETTERS))).encode(), dtype=np.uint8)


def encode_protein(text: str) -> np.ndarray:
    return PROTEIN_TO_CODE[np.frombuffer(text.encode("latin-1"), dtype=np.uint8)]


def decode_protein(codes: np.ndarray) -> str:
    return CODE_TO_PROTEIN[np.asarray(codes, dtype=np.uint8)].tobytes().decode("latin-1")


def seed_length(seed: int) -> int:
    """index.h:137-147."""
    return int(seed).bit_length()


def seed_weight(seed: int) -> int:
    """index.h:149-161."""
    return bin(int(seed)).count("1")


@dataclass
class DbChunk:
    """One `<db>_<i>` chunk: db.h / db_reader.cpp:52-77 / db.cpp:36-122."""
    seq: np.ndarray            # uint8[seq_len], every sequence followed by SEQUENCE_END
    seq_starts: np.ndarray     # uint32[n_seqs]   (.pos)
    names: List[str]           # (.nam)
    seed: int
    keys_count: np.ndarray     # uint32[32^w + 1] (.ind)
    positions: np.ndarray      # uint32[positions_len] (.ind)

    @property
    def n_seqs(self) -> int:
        return int(self.seq_starts.shape[0])


@dataclass
class Db:
    seed: int
    max_chunk_len: int
    sum_residues: int          # db_creator.cpp:419 (residues without separators), uint64 on disk
    chunks: List[DbChunk] = field(default_factory=list)

    @property
    def sum_length_u32(self) -> int:
        """db_reader.h:59-61 returns the uint64 truncated to uint32_t."""
        return self.sum_residues & 0xFFFFFFFF


@dataclass
class QueryChunk:
    """One `<q>_<i>` chunk: query.h:96-116 / query.cpp:36-78."""
    seqs: np.ndarray           # uint8[n, L], X-padded
    names: List[str]

    @property
    def n(self) -> int:
        return int(self.seqs.shape[0])

    @property
    def length(self) -> int:
        return int(self.seqs.shape[1])

    def name_breaks(self) -> np.ndarray:
        """aligner.cpp:697-700: 1 where the name differs from the previous query's."""
        b = np.zeros(self.n, dtype=np.uint8)
        for i in range(1, self.n):
            b[i] = self.names[i] != self.names[i - 1]
        return b


# --------------------------------------------------------------------------- index

def build_index(seq: np.ndarray, seq_starts: np.ndarray, seed: int):
    """db_creator.cpp:167-241 (ConstructIndex), vectorised.

    A position j is indexed iff its sequence is strictly longer than the seed span
    (:197), the span [j, j+len) holds no SEQUENCE_END (loop bound :198) and no X
    (:201-212).  positions are grouped by key, ascending inside a key (counting sort).
    """
    seq = np.ascontiguousarray(seq, dtype=np.uint8)
    n = seq.shape[0]
    slen = seed_length(seed)
    w = seed_weight(seed)
    n_keys = ALPHABET_SIZE ** w
    keys_count = np.zeros(n_keys + 1, dtype=np.uint32)
    if n < slen:
        return keys_count, np.zeros(0, dtype=np.uint32)
    m = n - slen + 1
    ok = np.ones(m, dtype=bool)
    key = np.zeros(m, dtype=np.int64)
    for i in range(slen):
        col = seq[i:i + m]
        ok &= (col != SEQUENCE_END) & (col != BASE_X)
        if (seed >> i) & 1:
            key = (key << CHARACTER_SIZE) | col
    # sequences with length <= seed span are skipped altogether (:197)
    starts = np.asarray(seq_starts, dtype=np.int64)
    ends = np.append(starts[1:], n) - 1           # index of each END separator
    short = (ends - starts) <= slen
    if short.any():
        for s, e in zip(starts[short], ends[short]):
            ok[s:min(e + 1, m)] = False
    pos = np.nonzero(ok)[0].astype(np.uint32)
    k = key[ok]
    order = np.argsort(k, kind="stable")
    positions = pos[order]
    keys_count[1:] = np.cumsum(np.bincount(k, minlength=n_keys)).astype(np.uint32)
    return keys_count, positions


# --------------------------------------------------------------------------- db files

def chunk_sequences(lengths: Sequence[int], max_chunk_len: int) -> List[range]:
    """db_creator.cpp:85-128 (ReadSequences): a chunk takes sequences while the sum of
    (length + 1) stays <= max_chunk_len."""
    out, start, total = [], 0, 0
    for i, ln in enumerate(lengths):
        if total + ln + 1 > max_chunk_len:
            if i == start:
                raise ValueError("error : too small max length.")
            out.append(range(start, i))
            start, total = i, 0
        total += ln + 1
    if start < len(lengths):
        out.append(range(start, len(lengths)))
    return out


def make_db_chunk(seqs: Sequence[np.ndarray], names: Sequence[str], seed: int) -> DbChunk:
    """db_creator.cpp:130-165 (ConvertSequences) + ConstructIndex."""
    lens = np.array([len(s) for s in seqs], dtype=np.int64)
    starts = np.zeros(len(seqs), dtype=np.uint32)
    if len(seqs) > 1:
        starts[1:] = np.cumsum(lens[:-1] + 1)
    total = int((lens + 1).sum())
    data = np.full(total, BASE_X, dtype=np.uint8)
    for s, st, ln in zip(seqs, starts, lens):
        data[st:st + ln] = s
        data[st + ln] = SEQUENCE_END
    keys_count, positions = build_index(data, starts, seed)
    return DbChunk(data, starts, list(names), seed, keys_count, positions)


def make_db(seqs: Sequence[np.ndarray], names: Sequence[str], seed_weight_k: int = 4,
            chunk_mib: float = 128) -> Db:
    """`ghostm db -k K -l MiB` (db_creator.cpp:369-479) in memory."""
    seed = (1 << seed_weight_k) - 1
    max_len = int(chunk_mib * (1 << 20))
    db = Db(seed=seed, max_chunk_len=max_len, sum_residues=0)
    for r in chunk_sequences([len(s) for s in seqs], max_len):
        ch = make_db_chunk([seqs[i] for i in r], [names[i] for i in r], seed)
        db.sum_residues += ch.seq.shape[0] - ch.n_seqs
        db.chunks.append(ch)
    return db


def write_db(prefix: str, db: Db) -> None:
    """db_creator.cpp:243-352 writers."""
    with open(prefix + ".inf", "wb") as f:
        div = np.int32(len(db.chunks))
        f.write(div.tobytes())
        f.write(np.uint32(db.seed).tobytes())
        f.write(np.uint32(db.max_chunk_len).tobytes())
        f.write(np.uint64(db.sum_residues).tobytes())
        f.write(np.full(32, div, dtype=np.int32).tobytes())
    for i, ch in enumerate(db.chunks):
        p = f"{prefix}_{i}"
        with open(p + ".inf", "wb") as f:
            f.write(np.uint32(ch.n_seqs).tobytes())
            f.write(np.uint32(ch.seq.shape[0]).tobytes())
        with open(p + ".nam", "w", encoding="latin-1", newline="\n") as f:
            for nm in ch.names:
                f.write(nm + "\n")
        ch.seq.tofile(p + ".seq")
        ch.seq_starts.astype(np.uint32).tofile(p + ".pos")
        with open(p + ".ind", "wb") as f:
            f.write(np.uint32(ch.seed).tobytes())
            f.write(np.uint32(ch.keys_count.shape[0]).tobytes())
            f.write(np.uint32(ch.positions.shape[0]).tobytes())
            f.write(ch.keys_count.astype(np.uint32).tobytes())
            f.write(ch.positions.astype(np.uint32).tobytes())


def _read_names(path: str, n: int) -> List[str]:
    """db.cpp:36-61 / query.cpp:36-61: getline per record."""
    names = []
    if os.path.exists(path):
        with open(path, "r", encoding="latin-1", newline="\n") as f:
            names = f.read().split("\n")
    names = names[:n]
    return names + [""] * (n - len(names))


def read_db(prefix: str) -> Db:
    """db_reader.cpp:36-77, db.cpp:36-122."""
    with open(prefix + ".inf", "rb") as f:
        raw = f.read(20)
    division = int(np.frombuffer(raw, dtype=np.int32, count=1)[0])
    seed, max_len = (int(x) for x in np.frombuffer(raw, dtype=np.uint32, count=2, offset=4))
    sum_res = int(np.frombuffer(raw, dtype=np.uint64, count=1, offset=12)[0])
    db = Db(seed=seed, max_chunk_len=max_len, sum_residues=sum_res)
    for i in range(division):
        p = f"{prefix}_{i}"
        n_seqs, seq_len = (int(x) for x in np.fromfile(p + ".inf", dtype=np.uint32, count=2))
        seq = np.fromfile(p + ".seq", dtype=np.uint8, count=seq_len)
        pos = np.fromfile(p + ".pos", dtype=np.uint32, count=n_seqs)
        hdr = np.fromfile(p + ".ind", dtype=np.uint32, count=3)
        kc = np.fromfile(p + ".ind", dtype=np.uint32, count=int(hdr[1]), offset=12)
        ps = np.fromfile(p + ".ind", dtype=np.uint32, count=int(hdr[2]), offset=12 + 4 * int(hdr[1]))
        db.chunks.append(DbChunk(seq, pos, _read_names(p + ".nam", n_seqs), int(hdr[0]), kc, ps))
    return db


# --------------------------------------------------------------------------- query files

def make_query_chunks(seqs: Sequence[np.ndarray], names: Sequence[str], length: int = 75,
                      chunk_mib: float = 128) -> List[QueryChunk]:
    """`ghostm qry -t p -l length -L MiB` (query_creator.cpp:189-234, 398-419)."""
    max_len = int(chunk_mib * (1 << 20))
    chunks: List[QueryChunk] = []
    cur: List[int] = []
    total = 0
    for i, s in enumerate(seqs):
        total += len(s)
        if total > max_len:                 # :226-229; the overflowing sequence opens the next chunk
            if not cur:
                raise ValueError("error : too small max length.")
            chunks.append(_query_chunk([seqs[j] for j in cur], [names[j] for j in cur], length))
            cur, total = [], len(s)
        cur.append(i)
    if cur:
        chunks.append(_query_chunk([seqs[j] for j in cur], [names[j] for j in cur], length))
    return chunks


def _query_chunk(seqs, names, length) -> QueryChunk:
    m = np.full((len(seqs), length), BASE_X, dtype=np.uint8)   # :410-412 X padding
    for i, s in enumerate(seqs):
        n = min(len(s), length)                                # :405-408 truncation
        m[i, :n] = s[:n]
    return QueryChunk(m, list(names))


def write_queries(prefix: str, chunks: Sequence[QueryChunk]) -> None:
    """query_creator.cpp:326-419 writers."""
    length = chunks[0].length if chunks else 0
    max_n = max((c.n for c in chunks), default=0)
    with open(prefix + ".inf", "wb") as f:
        div = np.int32(len(chunks))
        f.write(div.tobytes())
        f.write(np.uint32(length).tobytes())
        f.write(np.uint32(max_n).tobytes())
        f.write(np.full(32, div, dtype=np.int32).tobytes())
    for i, c in enumerate(chunks):
        p = f"{prefix}_{i}"
        with open(p + ".inf", "wb") as f:
            f.write(np.uint32(c.n).tobytes())
            f.write(np.uint32(c.length).tobytes())
        with open(p + ".nam", "w", encoding="latin-1", newline="\n") as f:
            for nm in c.names:
                f.write(nm + "\n")
        np.ascontiguousarray(c.seqs, dtype=np.uint8).tofile(p + ".seq")


def read_queries(prefix: str) -> List[QueryChunk]:
    """query_reader.cpp:36-101, query.cpp:36-78."""
    with open(prefix + ".inf", "rb") as f:
        raw = f.read(12)
    division = int(np.frombuffer(raw, dtype=np.int32, count=1)[0])
    out = []
    for i in range(division):
        p = f"{prefix}_{i}"
        n, length = (int(x) for x in np.fromfile(p + ".inf", dtype=np.uint32, count=2))
        seqs = np.fromfile(p + ".seq", dtype=np.uint8, count=n * length).reshape(n, length)
        out.append(QueryChunk(seqs, _read_names(p + ".nam", n)))
    return out


def write_fasta(path: str, names: Sequence[str], seqs: Sequence[np.ndarray], width: int = 0) -> None:
    with open(path, "w", encoding="latin-1", newline="\n") as f:
        for nm, s in zip(names, seqs):
            text = decode_protein(s)
            f.write(f">{nm}\n")
            if width:
                for k in range(0, len(text), width):
                    f.write(text[k:k + width] + "\n")
            else:
                f.write(text + "\n")
