"""Shared test harness: seeded synthetic cases, the oracle_runner case/dump file formats, and
the three ways a case is run (new CUDA backend through the C ABI, CPU oracle, reference dump).

A case is the unit the reference's pipeline hands to the backend for one (ref block, query
block) pair: src/main.cpp:613-661 (uploads + table), src/seeder.cpp:48-120 (per-chunk seed
vectors, both strands), src/seed_filter.cu:682-828 (SeedAndFilter).
"""
from __future__ import annotations

import hashlib
import struct
import subprocess
import sys
from dataclasses import dataclass, field
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

from segalign_b200 import genome  # noqa: E402

GOLDEN_DIR = ROOT / "tests" / "golden"
ORACLE_RUNNER = ROOT / "oracle" / "_ref" / "oracle_runner"
NEW_RUNNER = ROOT / "oracle" / "_ref" / "new_runner"

SEGMENT_DTYPE = np.dtype([("ref_start", "<u4"), ("query_start", "<u4"), ("len", "<u4"), ("score", "<i4")])


@dataclass
class Case:
    name: str
    gen: str                      # generator name in GENERATORS
    gen_args: dict = field(default_factory=dict)
    rng_seed: int = 20261017
    seed_shape: str = "12of19"
    transition: bool = True
    step: int = 1
    xdrop: int = 910
    hspthresh: int = 3000
    noentropy: bool = False
    wga_chunk: int = 250_000
    lastz_interval: int = 10_000_000
    max_hits_override: int = 0    # >0 forces the multi-iteration path
    strand: str = "both"
    ambiguous: str = ""

    def inputs(self):
        rng = np.random.default_rng(self.rng_seed)
        ref, query = GENERATORS[self.gen](rng, **self.gen_args)
        return np.ascontiguousarray(ref, dtype=np.uint8), np.ascontiguousarray(query, dtype=np.uint8)


# ------------------------------------------------------------------------------ generators
def _indels(seq, rng, every):
    """Short insertions/deletions about every `every` bases: bounds HSP length the way real
    genomes do (a substitution-only homolog is one genome-long diagonal)."""
    cuts = np.sort(rng.choice(np.arange(1, seq.size - 1), size=max(1, seq.size // every), replace=False))
    parts, prev = [], 0
    for c in cuts:
        parts.append(seq[prev:c])
        k = int(rng.integers(1, 12))
        if rng.random() < 0.5:
            parts.append(genome.random_genome(k, rng))
            prev = c
        else:
            prev = min(seq.size, c + k)
    parts.append(seq[prev:])
    return np.concatenate(parts)


def _homolog(rng, n, d, inversions=0, inv_len=0, indel_every=1500):
    ref = genome.random_genome(n, rng)
    q = genome.mutate(ref, d, rng)
    if indel_every:
        q = _indels(q, rng, indel_every)
        n = min(n, q.size)
    for s in (rng.integers(0, max(1, n - inv_len), size=inversions) if inversions else []):
        q[s:s + inv_len] = genome.revcomp_ascii(q[s:s + inv_len])
    return ref, q


def gen_diverged(rng, n=200_000, d=0.25, inversions=2, inv_len=20_000):
    return _homolog(rng, n, d, inversions, inv_len)


def gen_self(rng, n=30_000):
    ref = genome.random_genome(n, rng)
    return ref, ref.copy()


def gen_masked_multichrom(rng, n=300_000, d=0.2, chroms=4, f_mask=0.15, n_runs=3, n_len=500):
    """Several chromosomes joined by '&', soft-masked runs, N runs (A.1, App. C L/N/E rows)."""
    ref, q = _homolog(rng, n, d, 2, 15_000)
    ref = genome.soft_mask(ref, f_mask, rng)
    q = genome.soft_mask(q, f_mask, rng)
    ref = genome.insert_runs(ref, b"N", n_runs, n_len, rng)
    q = genome.insert_runs(q, b"N", n_runs, n_len, rng)
    cut_r = np.sort(rng.integers(1, n - 1, size=chroms - 1))
    cut_q = np.sort(rng.integers(1, n - 1, size=chroms - 1))
    rb = genome.make_blocks(np.split(ref, cut_r), block_size=10 ** 12)
    qb = genome.make_blocks(np.split(q, cut_q), block_size=10 ** 12)
    return rb[0], qb[0]


def gen_shared_ambiguous(rng, n=150_000, d=0.18, runs=60, run_len=6, query_iupac=False):
    """Short N / IUPAC runs at the SAME places in ref and query so that N-N and X-X cells sit
    inside HSPs under --ambiguous=iupac (the count[] aliasing of SURVEY A.6).  IUPAC letters in
    the query require strand="plus": the reference's host RevComp drops them (ntcoding.cpp:63-105)
    so its minus-strand seed positions no longer match the device's rev-comp block."""
    ref = genome.random_genome(n, rng)
    q = genome.mutate(ref, d, rng)
    for s in rng.integers(0, n - run_len, size=runs):
        ref[s:s + run_len] = ord("N")
        q[s:s + run_len] = ord("N")
    for s in rng.integers(0, n - run_len, size=runs):
        ref[s:s + 2] = ord("R")
        q[s:s + 2] = ord("R") if query_iupac else ord("N")
    return ref, _indels(q, rng, 1500)


def gen_repeats(rng, n=200_000, d=0.3, copies=150, elem=300, elem_div=0.1, low_complexity=40):
    """A repeat family (heavy buckets, many HSPs per diagonal neighbourhood) plus AT-rich
    low-complexity islands whose HSPs fall in [hspthresh, 3*hspthresh] and get entropy-scaled."""
    ref, q = _homolog(rng, n, d)
    element = genome.random_genome(elem, rng)
    for arr in (ref, q):
        for s in rng.integers(0, n - elem, size=copies):
            arr[s:s + elem] = genome.mutate(element, elem_div, rng)
    at = np.frombuffer(b"AT", dtype=np.uint8)
    for s in rng.integers(0, n - 200, size=low_complexity):
        island = at[(rng.random(120) < 0.5).astype(np.uint8)]
        ref[s:s + 120] = island
        q[s:s + 120] = genome.mutate(island, 0.05, rng)
    return ref, q


def gen_inverted_repeats(rng, n=30_000, copies=30, elem=300, elem_div=0.08):
    """One sequence with a repeat family whose copies sit on BOTH strands (every other copy is
    reverse-complemented): a self-alignment then has minus-strand HSPs (repeat-masker variant)."""
    seq = genome.random_genome(n, rng)
    element = genome.random_genome(elem, rng)
    for k, s in enumerate(rng.integers(0, n - elem, size=copies)):
        copy = genome.mutate(element, elem_div, rng)
        seq[s:s + elem] = genome.revcomp_ascii(copy) if k & 1 else copy
    return seq, seq.copy()


def gen_random_pair(rng, n_ref=400_000, n_query=100_000):
    """Unrelated sequences: only random hits, nothing passes (zero-anchor iterations)."""
    return genome.random_genome(n_ref, rng), genome.random_genome(n_query, rng)


GENERATORS = {
    "diverged": gen_diverged,
    "self": gen_self,
    "masked_multichrom": gen_masked_multichrom,
    "shared_ambiguous": gen_shared_ambiguous,
    "repeats": gen_repeats,
    "random_pair": gen_random_pair,
    "inverted_repeats": gen_inverted_repeats,
}

# Small cases: the CPU oracle finishes each in seconds.  Golden dumps of the UNMODIFIED
# reference for all of them are produced on a B200 by tests/golden/make_golden.py.
CASES = [
    Case("diverged_default", "diverged"),
    Case("diverged_chunked", "diverged", dict(n=120_000, d=0.2), rng_seed=11, wga_chunk=20_000,
         lastz_interval=50_000),
    Case("diverged_multi_iter", "diverged", dict(n=150_000, d=0.22), rng_seed=12, max_hits_override=3000),
    Case("self_align", "self"),
    Case("masked_multichrom", "masked_multichrom"),
    Case("iupac_notransition", "shared_ambiguous", dict(query_iupac=True), transition=False,
         ambiguous="iupac", hspthresh=2200, strand="plus"),
    Case("iupac_both_strands", "shared_ambiguous", rng_seed=6, ambiguous="iupac", hspthresh=2200),
    Case("ambiguous_n", "shared_ambiguous", rng_seed=5, ambiguous="n", hspthresh=2500),
    Case("repeats_entropy", "repeats"),
    Case("repeats_noentropy", "repeats", rng_seed=7, noentropy=True),
    Case("seed_14of22_notransition", "diverged", dict(n=250_000, d=0.15), rng_seed=21,
         seed_shape="14of22", transition=False),
    Case("step2", "diverged", dict(n=150_000, d=0.15), rng_seed=22, step=2),
    Case("custom_seed_plus_only", "diverged", dict(n=100_000, d=0.2), rng_seed=23,
         seed_shape="1110100110010101111", strand="plus"),
    Case("minus_only_low_thresh", "diverged", dict(n=80_000, d=0.3), rng_seed=24, strand="minus",
         hspthresh=1800, xdrop=600),
    Case("random_pair", "random_pair"),
]
CASES_BY_NAME = {c.name: c for c in CASES}


# ------------------------------------------------------------------------------ file formats
def matrix_for(case: Case) -> np.ndarray:
    from oracle import sa_oracle_py as sao
    return sao.build_matrix(case.ambiguous, case.xdrop)


def write_case_file(case: Case, path: Path, ref=None, query=None) -> None:
    """SACASE01, read by oracle/ref_driver.cpp:read_case."""
    if ref is None:
        ref, query = case.inputs()
    strand = {"both": 0, "plus": 1, "minus": 2}[case.strand]
    shape = case.seed_shape.encode()
    with open(path, "wb") as f:
        f.write(b"SACASE01")
        f.write(struct.pack("<I", len(shape)) + shape)
        f.write(struct.pack("<iIiiiIIii", int(case.transition), case.step, case.xdrop, case.hspthresh,
                            int(case.noentropy), case.wga_chunk, case.lastz_interval,
                            case.max_hits_override, strand))
        f.write(matrix_for(case).astype("<i4").tobytes())
        f.write(struct.pack("<Q", ref.size) + ref.tobytes())
        f.write(struct.pack("<Q", query.size) + query.tobytes())


@dataclass
class Dump:
    calls: list           # [(rev, chunk_start, chunk_end, num_seeds, total_anchors, num_hits, segs)]
    times: np.ndarray     # ref_upload, table, query_upload, seedgen, seed_and_filter (s)
    counters: np.ndarray  # seeds, hits, hsps, device MAX_HITS
    table: tuple | None = None


def read_dump(path: Path) -> Dump:
    """SAOUT001, written by oracle/ref_driver.cpp."""
    b = Path(path).read_bytes()
    assert b[:8] == b"SAOUT001", "bad dump magic"
    off = 8
    (ncalls,) = struct.unpack_from("<I", b, off); off += 4
    calls = []
    for _ in range(ncalls):
        rev, cs, ce, ns, nseg, tot, nh = struct.unpack_from("<7I", b, off); off += 28
        segs = np.frombuffer(b, dtype=SEGMENT_DTYPE, count=nseg, offset=off).copy(); off += 16 * nseg
        calls.append((rev, cs, ce, ns, tot, nh, segs))
    times = np.frombuffer(b, dtype="<f8", count=5, offset=off).copy(); off += 40
    counters = np.frombuffer(b, dtype="<u8", count=4, offset=off).copy(); off += 32
    (has_table,) = struct.unpack_from("<I", b, off); off += 4
    table = None
    if has_table:
        isz, npos = struct.unpack_from("<II", b, off); off += 8
        idx = np.frombuffer(b, dtype="<u4", count=isz, offset=off).copy(); off += 4 * isz
        pos = np.frombuffer(b, dtype="<u4", count=npos, offset=off).copy(); off += 4 * npos
        table = (idx, pos)
    return Dump(calls, times, counters, table)


def run_runner(binary: Path, case: Case, workdir: Path, extra=()) -> Dump:
    workdir.mkdir(parents=True, exist_ok=True)
    cf, of = workdir / f"{case.name}.case", workdir / f"{case.name}.{binary.name}.out"
    write_case_file(case, cf)
    subprocess.run([str(binary), str(cf), str(of), *extra], check=True, stderr=subprocess.PIPE)
    return read_dump(of)


# ------------------------------------------------------------------------------ golden fixtures
def golden_path(case: Case) -> Path:
    return GOLDEN_DIR / f"{case.name}.npz"


def inputs_digest(ref: np.ndarray, query: np.ndarray) -> str:
    h = hashlib.sha256()
    h.update(ref.tobytes()); h.update(b"|"); h.update(query.tobytes())
    return h.hexdigest()


def save_golden(case: Case, dump: Dump, ref, query) -> None:
    """Compact fixture: call headers + concatenated segments + input digest.  The inputs are
    re-generated from the case's RNG seed and checked against the digest."""
    hdr = np.array([(c[0], c[1], c[2], c[3], c[4], c[5], c[6].size) for c in dump.calls], dtype=np.uint32).reshape(-1, 7)
    segs = np.concatenate([c[6] for c in dump.calls]) if dump.calls else np.empty(0, SEGMENT_DTYPE)
    np.savez_compressed(golden_path(case), hdr=hdr, segs=segs.view(np.uint32).reshape(-1, 4),
                        digest=np.array(inputs_digest(ref, query)), max_hits_device=dump.counters[3])


def load_golden(case: Case):
    z = np.load(golden_path(case))
    hdr = z["hdr"]
    segs = np.ascontiguousarray(z["segs"]).view(SEGMENT_DTYPE).reshape(-1)
    calls, off = [], 0
    for rev, cs, ce, ns, tot, nh, nseg in hdr:
        calls.append((int(rev), int(cs), int(ce), int(ns), int(tot), int(nh), segs[off:off + nseg]))
        off += int(nseg)
    return calls, str(z["digest"])


# ------------------------------------------------------------------------------ runners
def chunk_calls(case: Case, q_len: int, span: int):
    return genome.chunk_list(q_len, span, case.strand, case.lastz_interval, case.wga_chunk)


def run_cpu_oracle(case: Case, ref=None, query=None, max_hits_device: int = 0xFFFFFFFF):
    """The whole case through oracle/sa_oracle.c.  Returns [(rev, j0, j1, num_seeds, segs_with_header)]."""
    from oracle import sa_oracle_py as sao
    if ref is None:
        ref, query = case.inputs()
    shape = sao.Shape(case.seed_shape)
    table = sao.Table(shape, ref, ref.size, case.step)
    ref_enc = sao.encode(ref)
    q_fwd, q_rc = sao.encode_rc(query)
    q_rc_ascii = sao.revcomp_ascii(query)
    mh = case.max_hits_override if case.max_hits_override > 0 else max_hits_device
    params = sao.make_params(matrix_for(case), case.xdrop, case.hspthresh, case.noentropy, shape.span, mh)
    out = []
    for rev, j0, j1 in chunk_calls(case, query.size, shape.span):
        seeds = shape.chunk_seeds(q_rc_ascii if rev else query, j0, j1, case.transition)
        if seeds.size == 0:
            continue
        res = sao.seed_and_filter(params, table, ref_enc, q_rc if rev else q_fwd, seeds)
        out.append((rev, j0, j1, seeds.size, res))
    return out


def setup_backend(be, case: Case, ref, query, buffer: int = 0):
    """InitializeProcessor ... SendQueryWriteRequest in the order of src/main.cpp:297-298,:613-661."""
    from segalign_b200.backend import shape_pattern
    weight = be.GenerateShapePos(case.seed_shape)
    span = len(shape_pattern(case.seed_shape))
    be.InitializeProcessor(case.transition, case.wga_chunk, span, matrix_for(case), case.xdrop,
                           case.hspthresh, case.noentropy)
    if case.max_hits_override > 0:
        be.set_max_hits(case.max_hits_override)
    be.SendRefWriteRequest(ref, 0, ref.size)
    be.GenerateSeedPosTable(ref, 0, ref.size, case.step)
    be.SendQueryWriteRequest(query, 0, query.size, buffer)
    return span, weight


def run_backend(be, case: Case, ref=None, query=None, device_seeding: bool = False):
    """The whole case through the CUDA backend (C ABI).  Same return layout as run_cpu_oracle."""
    from segalign_b200.backend import shape_pattern
    if ref is None:
        ref, query = case.inputs()
    span, _ = setup_backend(be, case, ref, query)
    pattern = shape_pattern(case.seed_shape)
    q_rc_ascii = genome.revcomp_ascii(query)
    out = []
    try:
        for rev, j0, j1 in chunk_calls(case, query.size, span):
            if device_seeding:
                res, ns = be.SeedAndFilterRange(j0, j1, case.transition, bool(rev), 0)
                if ns == 0:
                    continue
            else:
                seeds = genome.chunk_seeds(q_rc_ascii if rev else query, j0, j1, pattern, case.transition)
                if seeds.size == 0:
                    continue
                ns = seeds.size
                res = be.SeedAndFilter(seeds, bool(rev), 0)
            out.append((rev, j0, j1, ns, res))
    finally:
        be.ClearQuery(0)
        be.ClearRef()
        be.ShutdownProcessor()
    return out


def assert_calls_equal(got, want, what: str):
    """got/want: lists of (rev, j0, j1, num_seeds, segs_with_header)."""
    assert len(got) == len(want), f"{what}: {len(got)} calls vs {len(want)}"
    for g, w in zip(got, want):
        assert tuple(g[:4]) == tuple(w[:4]), f"{what}: call key {g[:4]} vs {w[:4]}"
        gs, ws = g[4], w[4]
        assert gs[0]["len"] == ws[0]["len"] and gs[0]["score"] == ws[0]["score"], \
            f"{what}: header of call {g[:3]}: anchors/hits {gs[0]['len']}/{gs[0]['score']} vs {ws[0]['len']}/{ws[0]['score']}"
        assert gs.size == ws.size, f"{what}: call {g[:3]}: {gs.size - 1} HSPs vs {ws.size - 1}"
        if not np.array_equal(gs[1:], ws[1:]):
            bad = np.flatnonzero(gs[1:] != ws[1:])[:5]
            raise AssertionError(f"{what}: call {g[:3]} differs at {bad}: {gs[1:][bad]} vs {ws[1:][bad]}")


def golden_as_calls(case: Case):
    """Golden dump -> the (rev, j0, j1, num_seeds, segs_with_header) layout."""
    calls, digest = load_golden(case)
    out = []
    for rev, cs, ce, ns, tot, nh, segs in calls:
        res = np.zeros(segs.size + 1, dtype=SEGMENT_DTYPE)
        res[0]["len"], res[0]["score"] = tot, np.int32(np.uint32(nh).view(np.int32))
        res[1:] = segs
        out.append((rev, cs, ce, ns, res))
    return out, digest


# ------------------------------------------------------------------------------ the reference's segment printer
SEGPRINT_RUNNER = ROOT / "oracle" / "_ref" / "segprint_runner"


def run_reference_printer(workdir: Path, r_table, q_table, rc_table, block, interval, fw, rc, *, data_folder="",
                          output_format="maf-", ambiguous="", scoring_file="", gapped=True, ydrop=9430,
                          gappedthresh=3000, notrivial=False):
    """One printer_input through the UNMODIFIED src/segment_printer.cpp (oracle/_ref/segprint_runner).

    tables = (names, starts, lens) with buffer offsets as src/main.cpp keeps them; block =
    (r_index [ref block index + 1, main.cpp:611], q_index, r_start, q_start, r_len, q_len [block length -
    seed size, main.cpp:714]); interval = (start, end, num_invoked).  Returns ({file name: text},
    [LASTZ command lines])."""
    workdir.mkdir(parents=True, exist_ok=True)
    out_dir = workdir / "printer_out"
    out_dir.mkdir(exist_ok=True)
    for f in out_dir.glob("*"):
        f.unlink()

    def s(x):
        b = x.encode()
        return struct.pack("<I", len(b)) + b

    def table(t):
        names, starts, lens = t
        out = struct.pack("<I", len(names))
        for n, st, ln in zip(names, starts, lens):
            out += s(n) + struct.pack("<QI", int(st), int(ln))
        return out
    inp = workdir / "printer.in"
    with open(inp, "wb") as f:
        f.write(b"SASEG001" + s(data_folder) + s(output_format) + s(ambiguous) + s(scoring_file))
        f.write(struct.pack("<iiii", int(gapped), ydrop, gappedthresh, int(notrivial)))
        f.write(table(r_table) + table(q_table) + table(rc_table))
        f.write(struct.pack("<iiQQII", *[int(v) for v in block]))
        f.write(struct.pack("<III", *[int(v) for v in interval]))
        for h in (fw, rc):
            h = np.ascontiguousarray(h, dtype=SEGMENT_DTYPE)
            f.write(struct.pack("<I", h.size) + h.tobytes())
    p = subprocess.run([str(SEGPRINT_RUNNER), str(inp), str(out_dir)], check=True, capture_output=True, text=True)
    return {q.name: q.read_text() for q in out_dir.glob("*.segments")}, p.stdout.splitlines()


# ------------------------------------------------------------------------------ repeat-masker variant
RM_ORACLE_RUNNER = ROOT / "oracle" / "_ref" / "rm_oracle_runner"
RM_NEW_RUNNER = ROOT / "oracle" / "_ref" / "rm_new_runner"
RM_GOLDEN_DIR = GOLDEN_DIR / "rm"

# The repeat masker aligns ONE sequence against itself: a case's reference is the sequence, its query
# is ignored.  neigh_prop = --neighbor_proportion (repeat_masker_src/main.cpp:50): which part of the
# sequence around a query interval its hits may come from.
RM_CASES = [
    (Case("rm_repeats", "repeats", dict(n=36_000, copies=40, low_complexity=10), lastz_interval=12_000, wga_chunk=5_000), 0.4),
    (Case("rm_repeats_whole", "repeats", dict(n=30_000, copies=30, low_complexity=8), rng_seed=31, wga_chunk=11_000), 1.0),
    (Case("rm_masked_multichrom", "masked_multichrom", dict(n=40_000, n_len=200), rng_seed=32, lastz_interval=9_000,
          wga_chunk=4_000), 0.5),
    (Case("rm_multi_iter", "repeats", dict(n=28_000, copies=25, low_complexity=6), rng_seed=33, lastz_interval=15_000,
          wga_chunk=6_000, max_hits_override=900), 0.6),
    (Case("rm_plus_only_noentropy", "repeats", dict(n=26_000, copies=24, low_complexity=12), rng_seed=34, strand="plus",
          noentropy=True, lastz_interval=8_000), 0.3),
    (Case("rm_inverted_repeats", "inverted_repeats", rng_seed=36, lastz_interval=11_000, wga_chunk=4_000), 0.6),
    (Case("rm_inverted_multi_iter", "inverted_repeats", dict(n=24_000, copies=24), rng_seed=37, lastz_interval=13_000,
          wga_chunk=6_500, max_hits_override=700), 1.0),
    (Case("rm_iupac_notransition", "shared_ambiguous", dict(n=30_000, runs=20), rng_seed=35, transition=False, strand="plus",
          ambiguous="iupac", hspthresh=2200, lastz_interval=10_000, wga_chunk=4_000), 0.7),
]
RM_CASES_BY_NAME = {c.name: (c, p) for c, p in RM_CASES}


def rm_intervals(seq_len: int, seed_size: int, interval: int, neigh_prop: float):
    """repeat_masker_src/main.cpp:323-420 for one block holding the whole sequence:
    [(start, end, ref_start, ref_end)]."""
    import math
    f32 = np.float32
    total = int(math.ceil(f32(seq_len) / f32(interval)))
    num_neigh = int(math.ceil(f32(f32(neigh_prop) * f32(total))))
    left_n = int(math.ceil(f32(num_neigh - 1) / f32(2)))
    right_n = num_neigh - 1 - left_n
    left_ov, right_ov = left_n * interval, right_n * interval
    max_len = left_ov + interval + right_ov
    out, start, end_pos = [], 0, seq_len - seed_size
    while start < end_pos:
        end = min(end_pos, start + interval)
        left_limit, right_limit = start < left_ov, end + right_ov > seq_len
        if left_limit:
            rs, re_ = 0, (seq_len if right_limit else min(seq_len, max_len))
        elif right_limit:
            rs, re_ = (0 if seq_len < max_len else seq_len - max_len), seq_len
        else:
            rs, re_ = start - left_ov, end + right_ov
        out.append((start, end, rs, re_))
        start += interval
    return out


def rm_calls(case: Case, seq_len: int, seed_size: int, neigh_prop: float):
    """Every SeedAndFilter call of the repeat masker's seeder (repeat_masker_src/seeder.cpp:69-150) in
    order: [(rev, j0, j1, ref_start, ref_end)]."""
    calls = []
    for (s, e, rs, re_) in rm_intervals(seq_len, seed_size, case.lastz_interval, neigh_prop):
        end_pos_rc = seq_len - 1 - s
        for i in range(s, e, case.wga_chunk):
            j0, j1 = i, min(i + case.wga_chunk, e)
            if case.strand in ("plus", "both"):
                calls.append((0, j0, j1, rs, re_))
            if case.strand in ("minus", "both"):
                r0 = seq_len - 1 - j1
                calls.append((1, r0, min(r0 + case.wga_chunk, end_pos_rc), rs, re_))
    return calls


@dataclass
class RmDump:
    calls: list           # [(rev, j0, j1, num_seeds, ref_start, ref_end, header(4 x u32), segs)]
    times: np.ndarray     # upload, table, seed_and_filter (s)
    counters: np.ndarray  # seeds, hits, hsps, device MAX_HITS


def read_rm_dump(path: Path) -> RmDump:
    """SARMO001, written by oracle/rm_driver.cpp."""
    b = Path(path).read_bytes()
    assert b[:8] == b"SARMO001", "bad dump magic"
    off = 8
    (ncalls,) = struct.unpack_from("<I", b, off); off += 4
    calls = []
    for _ in range(ncalls):
        rev, cs, ce, ns, rs, re_, nseg, h0, h1, h2, h3 = struct.unpack_from("<11I", b, off); off += 44
        segs = np.frombuffer(b, dtype=SEGMENT_DTYPE, count=nseg, offset=off).copy(); off += 16 * nseg
        calls.append((rev, cs, ce, ns, rs, re_, (h0, h1, h2, h3), segs))
    times = np.frombuffer(b, dtype="<f8", count=3, offset=off).copy(); off += 24
    counters = np.frombuffer(b, dtype="<u8", count=4, offset=off).copy()
    return RmDump(calls, times, counters)


def run_rm_runner(binary: Path, case: Case, neigh_prop: float, workdir: Path) -> RmDump:
    workdir.mkdir(parents=True, exist_ok=True)
    cf, of = workdir / f"{case.name}.case", workdir / f"{case.name}.{binary.name}.out"
    ref, _ = case.inputs()
    write_case_file(case, cf, ref, ref[:1])
    subprocess.run([str(binary), str(cf), str(of), "--neigh-prop", repr(float(neigh_prop))], check=True, stderr=subprocess.PIPE)
    return read_rm_dump(of)


def rm_golden_path(case: Case) -> Path:
    return RM_GOLDEN_DIR / f"{case.name}.npz"


def save_rm_golden(case: Case, dump: RmDump, seq) -> None:
    hdr = np.array([(c[0], c[1], c[2], c[3], c[4], c[5], *c[6], c[7].size) for c in dump.calls], dtype=np.uint32).reshape(-1, 11)
    segs = np.concatenate([c[7] for c in dump.calls]) if dump.calls else np.empty(0, SEGMENT_DTYPE)
    RM_GOLDEN_DIR.mkdir(exist_ok=True)
    np.savez_compressed(rm_golden_path(case), hdr=hdr, segs=segs.view(np.uint32).reshape(-1, 4),
                        digest=np.array(inputs_digest(seq, seq[:0])), max_hits_device=dump.counters[3])


def load_rm_golden(case: Case):
    """-> ([(rev, j0, j1, num_seeds, ref_start, ref_end, segs_with_header)], digest)"""
    z = np.load(rm_golden_path(case))
    segs = np.ascontiguousarray(z["segs"]).view(SEGMENT_DTYPE).reshape(-1)
    calls, off = [], 0
    for rev, cs, ce, ns, rs, re_, h0, h1, h2, h3, nseg in z["hdr"]:
        res = np.zeros(int(nseg) + 1, dtype=SEGMENT_DTYPE)
        res[0] = (h0, h1, h2, np.uint32(h3).view(np.int32))
        res[1:] = segs[off:off + int(nseg)]
        off += int(nseg)
        calls.append((int(rev), int(cs), int(ce), int(ns), int(rs), int(re_), res))
    return calls, str(z["digest"])


def run_rm_cpu_oracle(case: Case, neigh_prop: float, seq=None, max_hits_device: int = 0xFFFFFFFF):
    """The whole case through oracle/sa_oracle.c's repeat-masker restatement; same layout as load_rm_golden."""
    from oracle import sa_oracle_py as sao
    if seq is None:
        seq, _ = case.inputs()
    shape = sao.Shape(case.seed_shape)
    table = sao.Table(shape, seq, seq.size, case.step)
    enc = sao.encode(seq)
    enc_rc = sao.rm_revcomp_codes(enc)
    # host RevComp: what the seeder reads minus-strand seeds from.  The last minus-strand chunk runs up to
    # position len-1 (seeder.cpp:110-111), i.e. its seed spans reach past the sequence into the zero-filled
    # DRAM arena: both host buffers get that zero tail here
    tail = np.zeros(64, dtype=np.uint8)
    rc_ascii = np.concatenate([sao.revcomp_ascii(seq), tail])
    seq_padded = np.concatenate([seq, tail])
    mh = case.max_hits_override if case.max_hits_override > 0 else max_hits_device
    params = sao.make_params(matrix_for(case), case.xdrop, case.hspthresh, case.noentropy, shape.span, mh)
    out = []
    for rev, j0, j1, rs, re_ in rm_calls(case, seq.size, shape.span, neigh_prop):
        seeds = shape.chunk_seeds(rc_ascii if rev else seq_padded, j0, j1, case.transition)
        if seeds.size == 0:
            continue
        res = sao.rm_seed_and_filter(params, table, enc, enc_rc, seeds, bool(rev), rs, re_)
        out.append((rev, j0, j1, seeds.size, rs, re_, res))
    return out


def assert_rm_calls_equal(got, want, what: str):
    assert len(got) == len(want), f"{what}: {len(got)} calls vs {len(want)}"
    for g, w in zip(got, want):
        assert tuple(g[:6]) == tuple(w[:6]), f"{what}: call key {g[:6]} vs {w[:6]}"
        gs, ws = g[6], w[6]
        assert gs[0] == ws[0], f"{what}: header of call {g[:3]}: {gs[0]} vs {ws[0]}"
        assert gs.size == ws.size, f"{what}: call {g[:3]}: {gs.size - 1} HSPs vs {ws.size - 1}"
        if not np.array_equal(gs[1:], ws[1:]):
            bad = np.flatnonzero(gs[1:] != ws[1:])[:5]
            raise AssertionError(f"{what}: call {g[:3]} differs at {bad}: {gs[1:][bad]} vs {ws[1:][bad]}")


def setup_rm_backend(be, case: Case, seq):
    """repeat_masker_src/main.cpp:256-257, :498-505: InitializeProcessor, SendRefWriteRequest,
    SendQueryWriteRequest(), GenerateSeedPosTable."""
    from segalign_b200.backend import shape_pattern
    be.GenerateShapePos(case.seed_shape)
    span = len(shape_pattern(case.seed_shape))
    be.InitializeProcessor(case.transition, case.wga_chunk, span, matrix_for(case), case.xdrop, case.hspthresh,
                           case.noentropy)
    if case.max_hits_override > 0:
        be.set_max_hits(case.max_hits_override)
    be.SendRefWriteRequest(seq, 0, seq.size)
    be.RmSendQueryWriteRequest()
    be.GenerateSeedPosTable(seq, 0, seq.size, case.step)
    return span


def run_rm_backend(be, case: Case, neigh_prop: float, seq=None, device_seeding: bool = False):
    """The whole repeat-masker case through the CUDA backend (C ABI); layout of load_rm_golden."""
    from segalign_b200.backend import shape_pattern
    if seq is None:
        seq, _ = case.inputs()
    span = setup_rm_backend(be, case, seq)
    pattern = shape_pattern(case.seed_shape)
    rc_ascii = genome.revcomp_ascii(seq)
    out = []
    try:
        for rev, j0, j1, rs, re_ in rm_calls(case, seq.size, span, neigh_prop):
            if device_seeding:
                res, ns = be.RmSeedAndFilterRange(j0, j1, case.transition, bool(rev), rs, re_)
                if ns == 0:
                    continue
            else:
                seeds = genome.chunk_seeds(rc_ascii if rev else seq, j0, j1, pattern, case.transition)
                if seeds.size == 0:
                    continue
                ns = seeds.size
                res = be.RmSeedAndFilter(seeds, bool(rev), rs, re_)
            out.append((rev, j0, j1, ns, rs, re_, res))
    finally:
        be.RmClearQuery()
        be.ClearRef()
        be.ShutdownProcessor()
    return out
