"""Seeded synthetic genomes and the reference's blocking rules (host side, numpy).

No real genome exists offline, so every BASELINE.json config is restated as a seeded synthetic
input.  Blocks are built the way src/main.cpp does it: chromosomes joined by a
single '&' (main.cpp:407-411, :527-531), the block closing right after the chromosome that
pushes it past 500 000 000 bases (:359, :515), the last block losing its trailing '&'
(:414-415, :534-536).
"""
from __future__ import annotations

import numpy as np

ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
DEFAULT_SEQ_BLOCK_SIZE = 500_000_000  # src/graph.h:10
DEFAULT_LASTZ_INTERVAL = 10_000_000   # src/graph.h:11
DEFAULT_WGA_CHUNK = 250_000           # src/graph.h:12


def random_genome(n: int, rng: np.random.Generator) -> np.ndarray:
    """i.i.d. uniform ACGT, upper case, as ASCII bytes."""
    return ACGT[rng.integers(0, 4, size=n, dtype=np.uint8)]


def mutate(seq: np.ndarray, d: float, rng: np.random.Generator) -> np.ndarray:
    """Substitute each upper-case ACGT site with probability d by one of the 3 other bases."""
    out = seq.copy()
    lut = np.full(256, 255, dtype=np.uint8)
    lut[ACGT] = np.arange(4, dtype=np.uint8)
    code = lut[seq]
    sites = np.flatnonzero((rng.random(seq.size) < d) & (code < 4))
    shift = rng.integers(1, 4, size=sites.size, dtype=np.uint8)
    out[sites] = ACGT[(code[sites] + shift) & 3]
    return out


def soft_mask(seq: np.ndarray, frac: float, rng: np.random.Generator, mean_run: int = 300) -> np.ndarray:
    """Lower-case runs (geometric, mean `mean_run`) covering about `frac` of the sequence."""
    out = seq.copy()
    if frac <= 0:
        return out
    n_runs = max(1, int(seq.size * frac / mean_run))
    starts = rng.integers(0, seq.size, size=n_runs)
    lens = rng.geometric(1.0 / mean_run, size=n_runs)
    for s, l in zip(starts, lens):
        e = min(seq.size, s + l)
        out[s:e] |= 0x20
    return out


def insert_runs(seq: np.ndarray, char: bytes, n_runs: int, run_len: int, rng: np.random.Generator) -> np.ndarray:
    out = seq.copy()
    for s in rng.integers(0, max(1, seq.size - run_len), size=n_runs):
        out[s:s + run_len] = ord(char)
    return out


def sprinkle(seq: np.ndarray, chars: bytes, rate: float, rng: np.random.Generator) -> np.ndarray:
    """IUPAC / other letters at the given per-base rate."""
    out = seq.copy()
    sites = np.flatnonzero(rng.random(seq.size) < rate)
    alphabet = np.frombuffer(chars, dtype=np.uint8)
    out[sites] = alphabet[rng.integers(0, alphabet.size, size=sites.size)]
    return out


def revcomp_ascii(seq: np.ndarray) -> np.ndarray:
    """common/ntcoding.cpp:63-105 (alphabet restricted to what that switch handles)."""
    lut = np.arange(256, dtype=np.uint8)
    for a, b in zip(b"ACGTacgt", b"TGCAtgca"):
        lut[a] = b
    return lut[seq[::-1]]


def make_blocks(chroms: list[np.ndarray], block_size: int = DEFAULT_SEQ_BLOCK_SIZE) -> list[np.ndarray]:
    """Concatenate chromosomes into blocks exactly like src/main.cpp:336-415."""
    blocks, cur, cur_len = [], [], 0
    amp = np.frombuffer(b"&", dtype=np.uint8)
    for c in chroms:
        cur.append(c)
        cur_len += c.size
        if cur_len > block_size:
            blocks.append(np.concatenate(cur))
            cur, cur_len = [], 0
        else:
            cur.append(amp)
            cur_len += 1
    if cur_len > 0:
        blk = np.concatenate(cur)
        blocks.append(blk[:-1])  # drop the trailing '&'
    return blocks


def interval_list(block_len: int, seed_size: int, interval: int = DEFAULT_LASTZ_INTERVAL):
    """src/main.cpp:380-393: intervals over [0, block_len - seed_size) (exclusive)."""
    end_pos = block_len - seed_size
    out, cur = [], 0
    while cur < end_pos:
        out.append((cur, min(end_pos, cur + interval)))
        cur += interval
    return out


def chunk_list(block_len: int, seed_size: int, strand: str = "both",
               interval: int = DEFAULT_LASTZ_INTERVAL, chunk: int = DEFAULT_WGA_CHUNK):
    """All SeedAndFilter calls of one query block in the reference's order per interval
    (src/seeder.cpp:48-51, :89-90): list of (rev, j0, j1)."""
    q_block_len = block_len - seed_size  # main.cpp:714
    calls = []
    for (s, e) in interval_list(block_len, seed_size, interval):
        if strand in ("plus", "both"):
            for i in range(s, e, chunk):
                calls.append((0, i, min(i + chunk, e)))
        if strand in ("minus", "both"):
            rs, re_ = q_block_len - e, q_block_len - s
            for i in range(rs, re_, chunk):
                calls.append((1, i, min(i + chunk, re_)))
    return calls


# ---------------------------------------------------------------------------- seed words (host)
def shape_positions(pattern: str):
    pos = [i for i, c in enumerate(pattern) if c in "1T"]
    trans = [1 if pattern[i] == "T" else 0 for i in pos]
    return pos, trans


def chunk_seeds(seq: np.ndarray, j0: int, j1: int, pattern: str, transition: bool) -> np.ndarray:
    """Vectorised src/seeder.cpp:57-74 + common/ntcoding.cpp:43-61 for positions [j0, j1)."""
    span = len(pattern)
    pos, trans = shape_positions(pattern)
    w = len(pos)
    lut = np.full(256, 4, dtype=np.uint8)
    lut[ACGT] = np.arange(4, dtype=np.uint8)
    window = lut[seq[j0:j1 + span - 1]].astype(np.uint64)
    n = j1 - j0
    if window.size < n + span - 1:  # positions whose span runs past the block end are invalid
        window = np.concatenate([window, np.full(n + span - 1 - window.size, 4, dtype=np.uint64)])
    bad = (window > 3).astype(np.int32)
    csum = np.concatenate([[0], np.cumsum(bad)])
    valid = (csum[span:span + n] - csum[:n]) == 0
    kmer = np.zeros(n, dtype=np.uint64)
    for p in pos:
        kmer = (kmer << np.uint64(2)) | (window[p:p + n] & np.uint64(3))
    js = np.arange(j0, j1, dtype=np.uint64)[valid]
    kmer = kmer[valid]
    cols = [(kmer << np.uint64(32)) + js]
    if transition:
        for t in range(w):
            if trans[t]:
                cols.append(((kmer ^ (np.uint64(2) << np.uint64(2 * t))) << np.uint64(32)) + js)
    return np.stack(cols, axis=1).reshape(-1)


def block_tables(block: np.ndarray, prefix: str, block_start: int = 0):
    """Chromosome tables of one block as src/main.cpp keeps them: (names, starts, lens) in block
    order (q_chr_* / r_chr_*, main.cpp:341-344) and for the reverse-complement block
    (rc_q_chr_*, main.cpp:365-370: chromosomes in reverse order, start = 2*block_start + block_len
    - start - len).  Chromosomes are the '&'-separated pieces of the block."""
    amp = np.flatnonzero(block == ord("&"))
    starts = np.concatenate([[0], amp + 1]).astype(np.int64) + block_start
    ends = np.concatenate([amp, [block.size]]).astype(np.int64) + block_start
    lens = ends - starts
    names = [f"{prefix}{i}" for i in range(starts.size)]
    fwd = (names, starts.tolist(), lens.tolist())
    order = range(starts.size - 1, -1, -1)
    rc = ([names[i] for i in order],
          [int(2 * block_start + block.size - starts[i] - lens[i]) for i in order],
          [int(lens[i]) for i in order])
    return fwd, rc
