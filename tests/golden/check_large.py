"""Large-input parity + timing against the UNMODIFIED reference backend, on a GPU box.

    gpurun -- 'python tests/golden/check_large.py > gpurun_out/check_large.json'

For each case: oracle/_ref/oracle_runner (the reference's own CUDA kernels rebuilt for sm_100a,
SURVEY 8d "reference GPU timing") and the new backend run the same SeedAndFilter calls; every
call's records must be byte-identical; the time both spend inside SeedAndFilter is reported.
Too slow for the CPU oracle and too big to commit as fixtures -- hence a script, not a test.
"""
from __future__ import annotations

import json
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))

from segalign_b200 import genome  # noqa: E402
from tests import harness as H  # noqa: E402


def gen_ecoli_self(rng):
    g = genome.random_genome(4_641_652, rng)
    return g, g.copy()


def gen_ecoli_mut40(rng):
    g = genome.random_genome(4_641_652, rng)
    return g, genome.mutate(g, 0.40, rng)


def gen_worm_piece(rng):
    import bench
    chroms = bench.make_ref(bench.scaled_records(20))
    ref = genome.make_blocks(chroms)[0]
    q = genome.make_blocks(bench.make_query(chroms, 0))[0][:3_000_000]
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


H.GENERATORS.update(ecoli_self=gen_ecoli_self, ecoli_mut40=gen_ecoli_mut40, worm_piece=gen_worm_piece,
                    chr1_like=gen_chr1_like)
CASES = [
    H.Case("ecoli_mut40", "ecoli_mut40"),                      # BASELINE configs[0], throughput variant
    H.Case("worm_piece_20Mb_x_3Mb", "worm_piece"),             # configs[1] at reduced size, soft-masked
    H.Case("ecoli_self", "ecoli_self"),                        # configs[0]: main-diagonal blow-up
    H.Case("chr1_like_6Mb_x_4Mb_iupac_notransition", "chr1_like", transition=False, ambiguous="iupac",
           strand="plus"),                                    # configs[3] flags at reduced size
]


def main():
    from segalign_b200.backend import Backend
    only = sys.argv[1:]
    out = {}
    work = Path(tempfile.mkdtemp(prefix="sa_large_"))
    for case in CASES:
        if only and case.name not in only:
            continue
        ref, query = case.inputs()
        t0 = time.time()
        dump = H.run_runner(H.ORACLE_RUNNER, case, work)
        t_ref_total = time.time() - t0
        want = []
        for rev, cs, ce, ns, tot, nh, segs in dump.calls:
            res = np.zeros(segs.size + 1, dtype=H.SEGMENT_DTYPE)
            res[0]["len"], res[0]["score"] = tot, np.uint32(nh).view(np.int32)
            res[1:] = segs
            want.append((rev, cs, ce, ns, res))
        row = {"ref_bp": int(ref.size), "query_bp": int(query.size), "calls": len(want),
               "hits": int(dump.counters[1]), "hsps": int(dump.counters[2]),
               "reference_seed_and_filter_s": round(float(dump.times[4]), 3),
               "reference_table_build_s": round(float(dump.times[1]), 3),
               "reference_host_seedgen_s": round(float(dump.times[3]), 3)}
        for mode, dev in (("vector_abi", False), ("device_seeding", True)):
            be = Backend()
            be.InitializeInterface(1)
            span, _ = H.setup_backend(be, case, ref, query)
            from segalign_b200.backend import shape_pattern
            pattern = shape_pattern(case.seed_shape)
            q_rc = genome.revcomp_ascii(query)
            got, t_calls = [], 0.0
            for rev, j0, j1 in H.chunk_calls(case, query.size, span):
                if dev:
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
            be.ClearQuery(0); be.ClearRef(); be.ShutdownProcessor()
            try:
                H.assert_calls_equal(got, want, f"{case.name}/{mode}")
                row[mode] = "bit-identical"
            except AssertionError as e:
                row[mode] = f"MISMATCH: {e}"
            row[mode + "_seed_and_filter_s"] = round(t_calls, 3)
        row["speedup_vs_reference_kernels_vector_abi"] = round(row["reference_seed_and_filter_s"] / max(1e-9, row["vector_abi_seed_and_filter_s"]), 1)
        out[case.name] = row
        print(json.dumps({case.name: row}), flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
