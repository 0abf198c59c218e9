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


def gen_n_runs(rng, n=600_000, d=0.12):
    """Runs of N from 33 bases to 70 kb in either sequence, overlapping ones (N opposite N), a run opposite the other
    sequence's chromosome separator, frequent short soft-masked stretches (HSPs of a few hundred bases: the entropy
    path).  No indels: both sequences stay on one diagonal, so under --ambiguous=n|iupac the walk crosses a run at 0 per
    cell and the HSP resumes behind it -- what stage B's zero-run planes skip (zero_runs.h: zero_tile / zero_jump)."""
    amp = np.frombuffer(b"&", dtype=np.uint8)
    ref = np.concatenate([genome.random_genome(n // 2, rng), amp, genome.random_genome(n - n // 2 - 1, rng)])
    q = genome.mutate(ref, d, rng)
    ref = genome.soft_mask(ref, 0.25, rng, mean_run=80)
    q = genome.soft_mask(q, 0.25, rng, mean_run=80)
    # (start, length, flank): clean, closely related flanks on both sides of a run make HSPs cross it -- 400 bases
    # (score far above 3 x hspthresh: decided by the warp-per-hit kernel) or 45 bases fenced by lower case (hspthresh ..
    # 3 x hspthresh: the entropy path of the lane-pair kernel; behind a long run the entropy factor rejects the HSP, across
    # a run of 70 - 150 bases it passes)
    runs_r = [(20_000, 40, 400), (25_000, 70, 45), (27_000, 150, 45), (40_000, 700, 45), (60_000, 1_100, 400),
              (90_000, 2_100, 45), (120_000, 5_000, 400), (150_000, 40_000, 45), (n // 2 + 50_000, 70_000, 400),
              (n - 3_000, 3_000, 45)]
    runs_q = [(30_000, 33, 400), (33_000, 100, 45), (35_000, 130, 45), (175_000, 10_000, 45), (200_000, 1_500, 400),
              (215_000, 3_000, 45), (230_000, 36_000, 400), (n // 2 - 9_000, 20_000, 45), (0, 2_500, 400)]
    for s0, l, w in runs_r + runs_q:
        for lo, hi in ((max(0, s0 - w), s0), (s0 + l, min(n, s0 + l + w))):
            if hi <= lo:
                continue
            ref[lo:hi] &= 0xDF
            q[lo:hi] = genome.mutate(ref[lo:hi], 0.04, rng)
            if w == 45:
                if lo >= 4: ref[lo - 4:lo] |= 0x20
                if hi + 4 <= n: ref[hi:hi + 4] |= 0x20
    keep = ref == ord("&")
    for s0, l, _ in runs_r:
        ref[s0:s0 + l] = ord("N")
    ref[keep] = ord("&")
    for s0, l, _ in runs_q:
        q[s0:s0 + l] = ord("N")
    return ref, q


def gen_fresh(rng):
    """Everything at once, small: several chromosomes, soft-masking, N runs, inversions, a repeat family."""
    ref, q = H.gen_masked_multichrom(rng, n=400_000, d=0.22, chroms=3, f_mask=0.2)
    rr, qq = H.gen_repeats(rng, n=150_000, copies=60)
    amp = np.frombuffer(b"&", dtype=np.uint8)
    return np.concatenate([ref, amp, rr]), np.concatenate([q, amp, qq])


H.GENERATORS.update(ecoli_self=gen_ecoli_self, ecoli_mut40=gen_ecoli_mut40, worm_piece=gen_worm_piece,
                    chr1_like=gen_chr1_like, syn500_piece=gen_syn500_piece, fresh=gen_fresh,
                    masked_x_ambiguous=gen_masked_x_ambiguous, n_runs=gen_n_runs)

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

# long runs of N under --ambiguous: stage B's zero-run skipping, through each of its code paths
N_RUN_CASES = [
    H.Case("n_runs_iupac", "n_runs", ambiguous="iupac", rng_seed=51),
    H.Case("n_runs_n_notransition", "n_runs", ambiguous="n", transition=False, rng_seed=52),
]
N_RUN_KNOBS = ["", "SEGALIGN_B200_WIDE=0", "SEGALIGN_B200_ZERO_RUNS=0", "SEGALIGN_B200_FILTER=0", "SEGALIGN_B200_FUSED=0"]


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


@pytest.mark.parametrize("knob", N_RUN_KNOBS, ids=[k or "default" for k in N_RUN_KNOBS])
@pytest.mark.parametrize("case", N_RUN_CASES, ids=lambda c: c.name)
def test_n_runs_match_live_reference_kernels(case, knob, tmp_path, built, monkeypatch):
    """HSPs that cross runs of N (0 per cell under --ambiguous) and hits that walk into such runs: the backend skips
    the runs (zero-run planes), the reference walks them tile by tile; records must be byte-identical with the skipping on,
    off, and with every survivor on the lane-pair kernel."""
    from segalign_b200.backend import Backend
    if knob:
        k, v = knob.split("=")
        monkeypatch.setenv(k, v)
    ref, query = case.inputs()
    want, dump = reference_calls(case, tmp_path)
    assert int(dump.counters[2]) > 100, "the reference found too few HSPs for this to test anything"
    # the case must contain what it is there for: HSPs longer than the shortest long run, i.e. crossing one
    assert max(int(w[4][1:]["len"].max()) for w in want if w[4].size > 1) > 5_000
    for device_seeding in (False, True):
        be = Backend()
        be.InitializeInterface(1)
        got, _ = backend_calls(be, case, ref, query, device_seeding)
        H.assert_calls_equal(got, want, f"{case.name} [{knob or 'default'}]: backend vs the reference's own kernels")
