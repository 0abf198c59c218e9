"""Parity tests proper: the CUDA backend, called through the C ABI, against (i) the golden dumps
of the UNMODIFIED reference backend, (ii) the CPU oracle on fresh seeded inputs, and (iii)
size-independent properties at sizes the oracle cannot reach.  Bit-exact everywhere."""
import threading

import numpy as np
import pytest

from tests import harness as H

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", H.CASES, ids=lambda c: c.name)
def test_backend_matches_reference_golden(backend, case):
    want, digest = H.golden_as_calls(case)
    ref, query = case.inputs()
    assert H.inputs_digest(ref, query) == digest
    got = H.run_backend(backend, case, ref, query)
    H.assert_calls_equal(got, want, "backend vs reference golden")


@pytest.mark.parametrize("case", H.CASES, ids=lambda c: c.name)
def test_exact_stage_alone_matches_reference_golden(backend, case, monkeypatch):
    """SEGALIGN_B200_FILTER=0 sends every hit to the exact kernel (stage B): it must reproduce the
    reference on its own, so that the filter (stage A) can only ever remove non-HSPs."""
    monkeypatch.setenv("SEGALIGN_B200_FILTER", "0")
    want, _ = H.golden_as_calls(case)
    got = H.run_backend(backend, case)
    H.assert_calls_equal(got, want, "exact stage alone vs reference golden")


@pytest.mark.parametrize("name", ["repeats_entropy", "diverged_multi_iter", "self_align", "iupac_both_strands"])
def test_without_duplicate_table_matches_reference_golden(backend, name, monkeypatch):
    """SEGALIGN_B200_DEDUP=0: every passing record reaches the sort (the reference's own flow)."""
    monkeypatch.setenv("SEGALIGN_B200_DEDUP", "0")
    case = H.CASES_BY_NAME[name]
    want, _ = H.golden_as_calls(case)
    got = H.run_backend(backend, case)
    H.assert_calls_equal(got, want, "no duplicate table vs reference golden")


@pytest.mark.parametrize("case", H.CASES, ids=lambda c: c.name)
@pytest.mark.parametrize("device_seeding", [False, True], ids=["vector", "range"])
def test_general_path_matches_reference_golden(backend, case, device_seeding, monkeypatch):
    """SEGALIGN_B200_FUSED=0: hit counts -> scan -> iteration plan -> materialised hit list ->
    filter -> exact (the path every call with num_hits >= MAX_HITS takes)."""
    monkeypatch.setenv("SEGALIGN_B200_FUSED", "0")
    want, _ = H.golden_as_calls(case)
    got = H.run_backend(backend, case, device_seeding=device_seeding)
    H.assert_calls_equal(got, want, "general path vs reference golden")


@pytest.mark.parametrize("case", H.CASES, ids=lambda c: c.name)
@pytest.mark.parametrize("device_seeding", [False, True], ids=["vector", "range"])
def test_tile_walk_kernel_matches_reference_golden(backend, case, device_seeding, monkeypatch):
    """SEGALIGN_B200_FILTER_KERNEL=2: the tile-walk-only filter kernel (no popcount screen) that
    the default kernel (3) falls back to for matrices the screen does not admit."""
    kernel = "2"
    monkeypatch.setenv("SEGALIGN_B200_FILTER_KERNEL", kernel)
    want, _ = H.golden_as_calls(case)
    got = H.run_backend(backend, case, device_seeding=device_seeding)
    H.assert_calls_equal(got, want, f"filter kernel {kernel} vs reference golden")


@pytest.mark.parametrize("case", H.CASES, ids=lambda c: c.name)
def test_lane_pair_exact_kernel_alone_matches_reference_golden(backend, case, monkeypatch):
    """SEGALIGN_B200_WIDE=0: every survivor goes to k_extend_hits (two lanes per hit) instead of
    the warp-per-hit kernel k_extend_wide, which otherwise handles all hits that need no entropy
    factor.  Both must give the reference's bytes."""
    monkeypatch.setenv("SEGALIGN_B200_WIDE", "0")
    want, _ = H.golden_as_calls(case)
    got = H.run_backend(backend, case, device_seeding=True)
    H.assert_calls_equal(got, want, "lane-pair exact kernel alone vs reference golden")


@pytest.mark.parametrize("name", ["diverged_default", "masked_multichrom", "repeats_entropy"])
def test_plain_seed_vector_copy_matches_reference_golden(backend, name, monkeypatch):
    """SEGALIGN_B200_COMPACT_SEEDS=0: the seed vector is copied as handed over instead of as base
    words + device-side rebuild of the transition variants."""
    monkeypatch.setenv("SEGALIGN_B200_COMPACT_SEEDS", "0")
    case = H.CASES_BY_NAME[name]
    want, _ = H.golden_as_calls(case)
    got = H.run_backend(backend, case)
    H.assert_calls_equal(got, want, "plain seed vector copy vs reference golden")


def test_non_canonical_seed_vectors_match_cpu_oracle(backend):
    """Seed vectors that are NOT in the seeder's form (shuffled words, a truncated vector, words of one
    variant only) must take the plain copy and still give what the oracle gives for the same vector."""
    from oracle import sa_oracle_py as sao
    case = H.CASES_BY_NAME["diverged_default"]
    ref, query = case.inputs()
    span, _ = H.setup_backend(backend, case, ref, query)
    shape = sao.Shape(case.seed_shape)
    table = sao.Table(shape, ref, ref.size, case.step)
    ref_enc = sao.encode(ref)
    q_fwd, _ = sao.encode_rc(query)
    params = sao.make_params(H.matrix_for(case), case.xdrop, case.hspthresh, case.noentropy, shape.span, 748058112)
    seeds = shape.chunk_seeds(query, 0, min(60_000, query.size - span), case.transition)
    assert seeds.size % 13 == 0 and seeds.size > 13 * 4096
    rng = np.random.default_rng(9)
    variants = {"shuffled": rng.permutation(seeds), "truncated": seeds[:-5], "one_variant": seeds[3::13].copy(),
                "swapped_pair": np.concatenate([seeds[:13][::-1], seeds[13:]])}
    backend.reset_stats()
    for tag, vec in variants.items():
        vec = np.ascontiguousarray(vec)
        want = sao.seed_and_filter(params, table, ref_enc, q_fwd, vec)
        got = backend.SeedAndFilter(vec, False, 0)
        assert got[0]["len"] == want[0]["len"] and got[0]["score"] == want[0]["score"], tag
        assert np.array_equal(got[1:], want[1:]), tag
    # none of them may have gone through the compact upload
    assert backend.stats()["h2d_bytes"] == sum(v.size * 8 for v in variants.values())
    # ... while the canonical vector does
    backend.reset_stats()
    want = sao.seed_and_filter(params, table, ref_enc, q_fwd, seeds)
    got = backend.SeedAndFilter(seeds, False, 0)
    assert np.array_equal(got[1:], want[1:])
    assert backend.stats()["h2d_bytes"] == seeds.size * 8 // 13


@pytest.mark.parametrize("n_ref,n_query,off", [(64, 40, 3), (257, 64, 100), (1000, 333, 0), (5000, 20, 4000),
                                                (4097, 4097, 0), (100_000, 96, 99_904), (100_000, 700, 0)])
@pytest.mark.parametrize("device_seeding", [False, True], ids=["vector", "range"])
def test_tiny_and_ragged_blocks_match_cpu_oracle(backend, n_ref, n_query, off, device_seeding):
    """Blocks far smaller than the screen's 160-cell window, queries that end at the block end, a
    query of one seed span + 1: every window then runs into the terminator padding on one or both
    sides.  The query is a (lightly mutated, partly soft-masked) slice of the reference so that the
    calls do have hits."""
    from segalign_b200 import genome
    rng = np.random.default_rng(n_ref * 7 + n_query)
    ref = genome.random_genome(n_ref, rng)
    query = genome.mutate(ref[off:off + n_query].copy(), 0.05, rng)
    if n_query > 200:
        query = genome.soft_mask(query, 0.1, rng, mean_run=20)
    H.GENERATORS["_tiny"] = lambda _rng: (ref, query)
    case = H.Case(f"tiny_{n_ref}_{n_query}", "_tiny", hspthresh=1500 if n_query < 100 else 3000, wga_chunk=300)
    want = H.run_cpu_oracle(case, ref, query, max_hits_device=748058112)
    got = H.run_backend(backend, case, ref, query, device_seeding=device_seeding)
    H.assert_calls_equal(got, want, "tiny / ragged blocks vs cpu oracle")
    assert sum(w[4].size - 1 for w in want) > 0 or n_query <= 64


def test_screen_decides_most_random_hits(backend):
    """The popcount screen (default kernel) must be live -- few hits reach the tile walk on a
    diverged random pair -- and the tile-walk-only kernel must report none."""
    case = H.CASES_BY_NAME["random_pair"]
    ref, query = case.inputs()
    backend.reset_stats()
    H.run_backend(backend, case, ref, query, device_seeding=True)
    st = backend.stats()
    assert st["hits"] > 1000
    assert 0 < st["walked"] < 0.15 * st["hits"], st
    assert st["survivors"] <= st["walked"]


def test_filter_keeps_a_small_superset(backend):
    """The filter's survivors are few (it is the point of the stage) and contain every HSP."""
    case = H.CASES_BY_NAME["masked_multichrom"]
    ref, query = case.inputs()
    backend.reset_stats()
    got = H.run_backend(backend, case, ref, query)
    st = backend.stats()
    assert st["hits"] > 100_000
    # this small, closely related pair has ~23 % true (homologous) hits; random hits must be gone
    assert st["anchors_pre_dedupe"] <= st["survivors"] < 0.3 * st["hits"]
    assert sum(g[4].size - 1 for g in got) == st["hsps"]


@pytest.mark.parametrize("case", H.CASES, ids=lambda c: c.name)
def test_device_seeding_matches_reference_golden(backend, case):
    """sa_seed_and_filter_range (seed words generated on the GPU) returns exactly what
    the reference returns for the host-built seed vector of the same chunk."""
    want, _ = H.golden_as_calls(case)
    got = H.run_backend(backend, case, device_seeding=True)
    H.assert_calls_equal(got, want, "device seeding vs reference golden")


@pytest.mark.parametrize("seed", [101, 202, 303])
def test_backend_matches_cpu_oracle_fresh_inputs(backend, seed):
    case = H.Case(f"fresh_{seed}", "masked_multichrom", dict(n=120_000, d=0.22), rng_seed=seed,
                  wga_chunk=50_000, max_hits_override=40_000 if seed == 202 else 0)
    ref, query = case.inputs()
    want = H.run_cpu_oracle(case, ref, query, max_hits_device=748058112)
    got = H.run_backend(backend, case, ref, query)
    H.assert_calls_equal(got, want, "backend vs cpu oracle")


def test_encoding_matches_oracle(backend):
    from oracle import sa_oracle_py as sao
    rng = np.random.default_rng(5)
    for n in (1, 15, 16, 17, 31, 32, 33, 1000, 100_003):
        seq = rng.choice(np.frombuffer(b"ACGTacgtNn&RYKM-", dtype=np.uint8), size=n)
        case = H.Case("enc", "random_pair")
        backend.GenerateShapePos("12of19")
        backend.InitializeProcessor(True, 1000, 19, H.matrix_for(case), 910, 3000, False)
        backend.SendRefWriteRequest(seq, 0, n)
        backend.SendQueryWriteRequest(seq, 0, n, 1)
        fwd, rc = sao.encode_rc(seq)
        assert np.array_equal(backend.get_encoded(0, 0, n), sao.encode(seq))
        assert np.array_equal(backend.get_encoded(1, 1, n), fwd)
        assert np.array_equal(backend.get_encoded(2, 1, n), rc)
        backend.ClearQuery(1)
        backend.ClearRef()
        backend.ShutdownProcessor()
        backend.InitializeInterface(1)


@pytest.mark.parametrize("shape,step", [("12of19", 1), ("12of19", 3), ("14of22", 1), ("110101011", 2)])
def test_seed_position_table_matches_oracle(backend, shape, step):
    """index_table identical; pos_table identical as per-bucket multisets (order inside a bucket
    is scheduling-dependent in the reference)."""
    from oracle import sa_oracle_py as sao
    from tests.harness import genome
    rng = np.random.default_rng(8)
    ref = genome.soft_mask(genome.random_genome(300_000, rng), 0.2, rng)
    ref = genome.insert_runs(ref, b"N", 3, 400, rng)
    ref[1234] = ord("&")
    sh = sao.Shape(shape)
    backend.GenerateShapePos(shape)
    backend.InitializeProcessor(True, 1000, sh.span, H.matrix_for(H.Case("t", "random_pair")), 910, 3000, False)
    backend.SendRefWriteRequest(ref, 0, ref.size)
    backend.GenerateSeedPosTable(ref, 0, ref.size, step)
    idx, pos = backend.get_table()
    tab = sao.Table(sh, ref, ref.size, step)
    assert np.array_equal(idx, tab.index)
    assert pos.size == tab.pos.size
    starts = np.concatenate([[0], idx[:-1]]).astype(np.int64)
    bucket = np.repeat(np.arange(idx.size, dtype=np.int64), idx.astype(np.int64) - starts)
    assert np.array_equal(pos[np.lexsort((pos, bucket))], tab.pos[np.lexsort((tab.pos, bucket))])
    backend.ClearRef()


def _big_case():
    return H.Case("big", "masked_multichrom", dict(n=3_000_000, d=0.3, chroms=3, f_mask=0.1), rng_seed=77)


def test_properties_at_scale(backend):
    """3 Mb x 3 Mb (beyond what the CPU oracle does in seconds): per-call output is sorted by
    hspCompLastz within each iteration, survives its own sort+dedupe unchanged (idempotence, via
    the oracle's sort_dedupe), every HSP re-extends to itself from its own start, and repeated /
    concurrent calls return identical bytes."""
    from oracle import sa_oracle_py as sao
    case = _big_case()
    ref, query = case.inputs()
    span, _ = H.setup_backend(backend, case, ref, query)
    units = H.chunk_calls(case, query.size, span)
    outs = [backend.SeedAndFilterRange(j0, j1, True, bool(rev), 0)[0] for rev, j0, j1 in units]
    total = sum(o.size - 1 for o in outs)
    assert total > 500
    for o in outs:
        assert o[0]["len"] == o.size - 1
        segs = o[1:]
        if segs.size < 2:
            continue
        # at most two iterations (A.7): a sorted run, optionally followed by a second sorted run
        key = segs["query_start"].astype(np.int64)
        assert (np.diff(key) < 0).sum() <= 1
        # dedupe idempotence on the first run
        cut = int(np.argmax(np.diff(key) < 0)) + 1 if (np.diff(key) < 0).any() else segs.size
        assert np.array_equal(sao.sort_dedupe(segs[:cut]), segs[:cut])
    # checksum of checksums: repeat sequentially and from 4 threads
    def checksum(arrs):
        return [hash(a[1:].tobytes()) for a in arrs]
    base = checksum(outs)
    again = [None] * len(units)

    def work(tid):
        for i in range(tid, len(units), 4):
            rev, j0, j1 = units[i]
            again[i] = backend.SeedAndFilterRange(j0, j1, True, bool(rev), 0)[0]
    threads = [threading.Thread(target=work, args=(t,)) for t in range(4)]
    [t.start() for t in threads]
    [t.join() for t in threads]
    assert checksum(again) == base
    # spot-check 200 HSPs against the CPU oracle's single-hit extension: extending from the
    # HSP's own last cell + 1 ... is not an invariant, but scoring the segment is:
    sub = H.matrix_for(case).reshape(8, 8)
    ref_enc = sao.encode(ref)
    q_fwd, q_rc = sao.encode_rc(query)
    rng = np.random.default_rng(0)
    checked = 0
    for (rev, j0, j1), o in zip(units, outs):
        q = q_rc if rev else q_fwd
        for s in o[1:][rng.permutation(o.size - 1)[:20]]:
            r0, q0, n = int(s["ref_start"]), int(s["query_start"]), int(s["len"]) + 1
            raw = int(sub[ref_enc[r0:r0 + n], q[q0:q0 + n]].sum())
            assert raw >= case.hspthresh and s["score"] <= raw
            assert s["score"] == raw or raw <= 3 * case.hspthresh  # only entropy may lower it
            checked += 1
    assert checked > 100
    backend.ClearQuery(0)
    backend.ClearRef()


def test_vector_abi_equals_device_seeding_at_scale(backend):
    case = H.Case("big2", "diverged", dict(n=1_500_000, d=0.3), rng_seed=78)
    ref, query = case.inputs()
    a = H.run_backend(backend, case, ref, query, device_seeding=False)
    backend.InitializeInterface(1)
    b = H.run_backend(backend, case, ref, query, device_seeding=True)
    H.assert_calls_equal(a, b, "vector ABI vs device seeding")


def test_state_errors(backend):
    from segalign_b200.backend import BackendError
    with pytest.raises(BackendError) as ei:
        backend.SeedAndFilter(np.array([1], dtype=np.uint64), False, 0)
    assert ei.value.code == -21
    backend.GenerateShapePos("12of19")
    backend.InitializeProcessor(True, 100, 19, H.matrix_for(H.Case("t", "random_pair")), 910, 3000, False)
    seq = np.frombuffer(b"ACGT" * 100, dtype=np.uint8)
    backend.SendRefWriteRequest(seq, 0, seq.size)
    backend.GenerateSeedPosTable(seq, 0, seq.size, 1)
    backend.SendQueryWriteRequest(seq, 0, seq.size, 0)
    with pytest.raises(BackendError) as ei:  # MAX_SEEDS exceeded (seed_filter.cu:688-692)
        backend.SeedAndFilter(np.zeros(13 * 100 + 1, dtype=np.uint64), False, 0)
    assert ei.value.code == -20
    with pytest.raises(BackendError):
        backend.SendQueryWriteRequest(seq, 0, seq.size, 2)  # BUFFER_DEPTH = 2
    # empty and header-only results
    out = backend.SeedAndFilter(np.empty(0, dtype=np.uint64), False, 0)
    assert out.size == 1 and out[0]["len"] == 0 and out[0]["score"] == 0


@pytest.mark.parametrize("name", ["diverged_chunked", "masked_multichrom", "diverged_multi_iter", "repeats_entropy"])
def test_dropin_runner_with_reference_host_objects(name, tmp_path, built):
    """oracle/_ref/new_runner = the reference's own ntcoding.o, DRAM.o and seeder.o + the driver
    that produced the golden dumps, linked against segalign_b200/csrc/shim.cpp (g_* symbols,
    GenerateSeedPosTable) + libsegalign_b200.so instead of the reference's three .cu files.
    --check-seeder additionally runs the reference's unmodified seeder_body functor on it."""
    if not H.NEW_RUNNER.exists():
        pytest.skip("oracle/_ref/new_runner not built (needs /root/reference at build time)")
    case = H.CASES_BY_NAME[name]
    dump = H.run_runner(H.NEW_RUNNER, case, tmp_path, extra=("--check-seeder", "--dump-table"))
    want, _ = H.golden_as_calls(case)
    got = []
    for rev, cs, ce, ns, tot, nh, segs in dump.calls:
        res = np.zeros(segs.size + 1, dtype=H.SEGMENT_DTYPE)
        res[0]["len"], res[0]["score"] = tot, np.uint32(nh).view(np.int32)
        res[1:] = segs
        got.append((rev, cs, ce, ns, res))
    H.assert_calls_equal(got, want, "new_runner (reference host objects + shim) vs reference golden")


@pytest.mark.parametrize("case", H.CASES, ids=lambda c: c.name)
@pytest.mark.parametrize("device_seeding", [False, True], ids=["vector", "range"])
def test_merge_pass_matches_reference_golden(backend, case, device_seeding, monkeypatch):
    """SEGALIGN_B200_MERGE_MIN=8: (almost) every call takes the merge pass of kernels_merge.cuh -- survivors
    sorted by (diagonal, anchor), those joined to their predecessor by an all-match stretch dropped as
    provable copies, stage B replayed on the rest.  Must be result-neutral on every golden case."""
    monkeypatch.setenv("SEGALIGN_B200_MERGE_MIN", "8")
    want, _ = H.golden_as_calls(case)
    got = H.run_backend(backend, case, device_seeding=device_seeding)
    H.assert_calls_equal(got, want, "merge pass vs reference golden")


def test_merge_pass_collapses_a_self_alignment(backend, monkeypatch):
    """The main diagonal of a self-alignment is one all-match chain: with the merge pass its hits are
    extended once per call instead of once per hit, and the records stay those of the reference."""
    monkeypatch.setenv("SEGALIGN_B200_MERGE_MIN", "1000")
    case = H.CASES_BY_NAME["self_align"]
    want, _ = H.golden_as_calls(case)
    ref, query = case.inputs()
    span, _ = H.setup_backend(backend, case, ref, query)
    backend.reset_stats()
    got = []
    for rev, j0, j1 in H.chunk_calls(case, query.size, span):
        res, ns = backend.SeedAndFilterRange(j0, j1, case.transition, bool(rev), 0)
        if ns:
            got.append((rev, j0, j1, ns, res))
    st = backend.stats()
    H.assert_calls_equal(got, want, "self-alignment through the merge pass vs reference golden")
    assert st["merge_calls"] >= 1
    # ~30 k main-diagonal survivors on the plus strand collapse to a handful of representatives
    assert st["merge_dropped"] > 0.9 * (query.size - span)


@pytest.mark.parametrize("case", H.CASES, ids=lambda c: c.name)
def test_device_wide_radix_sort_path_matches_reference_golden(backend, case, monkeypatch):
    """SEGALIGN_B200_FINALIZE_CAP=8: calls with more than 8 anchors skip the one-block bitonic finalisation
    and take the device-wide path -- stable 64-bit radix passes over the composite keys of
    src/seed_filter.cu:54-108, predecessor dedupe, radix passes for the final order."""
    monkeypatch.setenv("SEGALIGN_B200_FINALIZE_CAP", "8")
    want, _ = H.golden_as_calls(case)
    got = H.run_backend(backend, case, device_seeding=True)
    H.assert_calls_equal(got, want, "device-wide radix sort path vs reference golden")
