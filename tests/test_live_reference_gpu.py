"""Live parity on the GPU box: the UNMODIFIED reference kernels (oracle/_ref/oracle_runner: the
reference's seed_filter.cu / seed_pos_table.cu / seed_filter_interface.cu rebuilt for sm_100a by
oracle/Makefile) and this backend run the same SeedAndFilter calls on inputs that are far too large
to commit as golden fixtures -- one case per BASELINE.json config, plus one case whose inputs are
drawn from a FRESH seed on every run, so that a stale golden can never hide a regression.
Every call's records must be byte-identical through both entry points (seed vectors and device
seeding).  Skipped only where the runner was not built (no /root/reference at build time)."""
import os
import time

import numpy as np
import pytest

from segalign_b200 import genome
from tests import harness as H

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not H.ORACLE_RUNNER.exists(),
                                 reason="oracle/_ref/oracle_runner not built (needs /root/reference at build time)")]


# ------------------------------------------------------------------------------ inputs (seeded, synthetic)
def gen_ecoli_self(rng, n=4_641_652):
    g = genome.random_genome(n, rng)
    return g, g.copy()


def gen_ecoli_mut40(rng):
    g = genome.random_genome(4_641_652, rng)
    return g, genome.mutate(g, 0.40, rng)


def gen_worm_piece(rng):
    import bench
    chroms = bench.make_ce11_ref(bench.scaled_records(20))
    ref = genome.make_blocks(chroms)[0]
    q = genome.make_blocks(bench.make_ce11_query(chroms, 0))[0][:3_000_000]
    return ref, q


def gen_syn500_piece(rng, ref_mb=100, query_mb=2):
    """BASELINE configs[2] at 1/5 of the reference block: 5 unmasked records, query = mutate(0.40)."""
    chroms = [genome.random_genome(ref_mb * 200_000, rng) for _ in range(5)]
    ref = genome.make_blocks(chroms)[0]
    q = genome.mutate(chroms[1][7_000_000:7_000_000 + query_mb * 1_000_000], 0.40, rng)
    return ref, q


def gen_chr1_like(rng):
    """BASELINE configs[3] at reduced size: half soft-masked reference with long N runs, a 30 %-diverged
    query with shared short N / IUPAC runs (they sit inside HSPs under --ambiguous=iupac) and its own
    soft-masking; run with --notransition --ambiguous=iupac on the plus strand (IUPAC letters in the
    query: the reference's host RevComp drops them, see harness.gen_shared_ambiguous)."""
    n = 6_000_000
    ref = genome.random_genome(n, rng)
    q = genome.mutate(ref, 0.30, rng)
    for s in rng.integers(0, n - 8, size=3000):
        ref[s:s + 6] = ord("N"); q[s:s + 6] = ord("N")
    for s in rng.integers(0, n - 8, size=3000):
        ref[s:s + 2] = ord("R"); q[s:s + 2] = ord("R")
    ref = genome.insert_runs(genome.soft_mask(ref, 0.5, rng), b"N", 3, 150_000, rng)
    q = genome.soft_mask(q, 0.45, rng)
    return ref, q[:4_000_000]


def gen_masked_x_ambiguous(rng, n=400_000):
    """Lower case opposite N, N opposite lower case, lower case opposite lower case -- inside homologous sequence.
    Under --ambiguous=n|iupac the first two pairs score 0 and HSPs run through them, the third is -1000 and stops
    them: lower case is a terminator code for the filter stage only while the opposite cell is not N
    (screen_bound.h: screen_terminator_codes; kernels_filter.cuh: tile_walk's inclusive soft test)."""
    ref = genome.random_genome(n, rng)
    q = genome.mutate(ref, 0.15, rng)
    sites = rng.permutation(n - 2)[:n // 40]
    a, b, c = np.array_split(sites, 3)
    ref[a] |= 0x20; q[a] = ord("N")
    ref[b] = ord("N"); q[b] |= 0x20
    ref[c] |= 0x20; q[c] |= 0x20
    return ref, q


def gen_fresh(rng):
    """Everything at once, small: several chromosomes, soft-masking, N runs, inversions, a repeat family."""
    ref, q = H.gen_masked_multichrom(rng, n=400_000, d=0.22, chroms=3, f_mask=0.2)
    rr, qq = H.gen_repeats(rng, n=150_000, copies=60)
    amp = np.frombuffer(b"&", dtype=np.uint8)
    return np.concatenate([ref, amp, rr]), np.concatenate([q, amp, qq])


H.GENERATORS.update(ecoli_self=gen_ecoli_self, ecoli_mut40=gen_ecoli_mut40, worm_piece=gen_worm_piece,
                    chr1_like=gen_chr1_like, syn500_piece=gen_syn500_piece, fresh=gen_fresh,
                    masked_x_ambiguous=gen_masked_x_ambiguous)

# SEGALIGN_LIVE_FULL=1 (tests/golden/check_large.py sets it) runs the self-alignment at E. coli size:
# the reference kernels need ~270 s for it on a B200 (every main-diagonal hit re-walks the diagonal).
SELF_N = 4_641_652 if os.environ.get("SEGALIGN_LIVE_FULL") else 1_200_000
FRESH_SEED = int(os.environ.get("SEGALIGN_LIVE_SEED", "0")) or (int(time.time()) & 0x7FFFFFFF)

LIVE_CASES = [
    H.Case("ecoli_mut40", "ecoli_mut40"),                      # BASELINE configs[0], throughput variant
    H.Case("ecoli_self", "ecoli_self", dict(n=SELF_N)),        # configs[0]: main-diagonal blow-up
    H.Case("worm_piece_20Mb_x_3Mb", "worm_piece"),             # configs[1] at reduced size, soft-masked
    H.Case("syn500_piece_100Mb_x_2Mb", "syn500_piece"),        # configs[2] at reduced size
    H.Case("chr1_like_6Mb_x_4Mb_iupac_notransition", "chr1_like", transition=False, ambiguous="iupac",
           strand="plus"),                                    # configs[3] flags at reduced size
    H.Case("masked_x_ambiguous_iupac", "masked_x_ambiguous", ambiguous="iupac", rng_seed=41),
    H.Case("masked_x_ambiguous_n_notransition", "masked_x_ambiguous", ambiguous="n", transition=False, rng_seed=42),
    H.Case("fresh_seed", "fresh", rng_seed=FRESH_SEED, wga_chunk=100_000),
]


def reference_calls(case, workdir):
    """The case through oracle_runner -> [(rev, j0, j1, num_seeds, segs_with_header)], and its dump."""
    dump = H.run_runner(H.ORACLE_RUNNER, case, workdir)
    want = []
    for rev, cs, ce, ns, tot, nh, segs in dump.calls:
        res = np.zeros(segs.size + 1, dtype=H.SEGMENT_DTYPE)
        res[0]["len"], res[0]["score"] = tot, np.uint32(nh).view(np.int32)
        res[1:] = segs
        want.append((rev, cs, ce, ns, res))
    return want, dump


def backend_calls(be, case, ref, query, device_seeding):
    """Same calls through the C ABI; returns the calls and the seconds spent inside SeedAndFilter."""
    span, _ = H.setup_backend(be, case, ref, query)
    q_rc = genome.revcomp_ascii(query)
    got, t_calls = [], 0.0
    try:
        for rev, j0, j1 in H.chunk_calls(case, query.size, span):
            if device_seeding:
                t1 = time.perf_counter()
                res, ns = be.SeedAndFilterRange(j0, j1, case.transition, bool(rev), 0)
                t_calls += time.perf_counter() - t1
            else:
                seeds = be.host_chunk_seeds(q_rc if rev else query, j0, j1, case.transition)
                ns = seeds.size
                if ns == 0:
                    continue
                t1 = time.perf_counter()
                res = be.SeedAndFilter(seeds, bool(rev), 0)
                t_calls += time.perf_counter() - t1
            if ns:
                got.append((rev, j0, j1, ns, res))
    finally:
        be.ClearQuery(0); be.ClearRef(); be.ShutdownProcessor()
    return got, t_calls


@pytest.mark.parametrize("case", LIVE_CASES, ids=lambda c: c.name)
def test_backend_matches_live_reference_kernels(case, tmp_path, built):
    from segalign_b200.backend import Backend
    if case.name == "fresh_seed":
        print(f"fresh_seed: rng_seed={case.rng_seed} (SEGALIGN_LIVE_SEED reproduces it)")
    ref, query = case.inputs()
    want, dump = reference_calls(case, tmp_path)
    assert len(want) > 0 and int(dump.counters[1]) > 0, "the reference made no SeedAndFilter call / found no hit"
    for f in tmp_path.glob("*"):
        f.unlink()
    for device_seeding in (False, True):
        be = Backend()
        be.InitializeInterface(1)
        got, _ = backend_calls(be, case, ref, query, device_seeding)
        H.assert_calls_equal(got, want, f"{case.name}: backend ({'device seeding' if device_seeding else 'seed vectors'}) "
                                        f"vs the reference's own kernels (seed {case.rng_seed})")
