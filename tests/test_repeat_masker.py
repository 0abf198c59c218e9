"""the repeat-masker variant of the backend (segalign_repeat_masker links
repeat_masker_src/seed_filter.cu instead of src/seed_filter.cu).

Pin: tests/golden/rm/*.npz are dumps of oracle/_ref/rm_oracle_runner = the reference's UNMODIFIED
repeat_masker_src/seed_filter.cu + common/*.cu run on a B200 (tests/golden/make_golden_rm.py).
CPU tests: the restatement (oracle/sa_oracle.c: sao_rm_*) against those dumps + unit tests of its
sort / unique chain.  GPU tests: the CUDA backend through the C ABI (sa_rm_*), bit-exact against the dumps
through both entry points and every code path, and through the reference's own symbols (rm_new_runner)."""
import numpy as np
import pytest

from tests import harness as H

HAVE_GOLDEN = all(H.rm_golden_path(c).exists() for c, _ in H.RM_CASES)
needs_golden = pytest.mark.skipif(not HAVE_GOLDEN, reason="tests/golden/rm/*.npz not generated yet (make_golden_rm.py on a GPU box)")
RM_IDS = [c.name for c, _ in H.RM_CASES]


# ------------------------------------------------------------------------------ CPU
@needs_golden
@pytest.mark.parametrize("case,prop", H.RM_CASES, ids=RM_IDS)
def test_rm_cpu_oracle_matches_reference_dump(case, prop):
    want, digest = H.load_rm_golden(case)
    seq, _ = case.inputs()
    assert H.inputs_digest(seq, seq[:0]) == digest, "synthetic inputs are not reproducible from the seed"
    got = H.run_rm_cpu_oracle(case, prop, seq, max_hits_device=748058112)
    H.assert_rm_calls_equal(got, want, "rm cpu-oracle vs reference golden")


@needs_golden
def test_rm_golden_is_nontrivial():
    total, minus, windows = 0, 0, set()
    for case, _ in H.RM_CASES:
        calls, _ = H.load_rm_golden(case)
        total += sum(c[6].size - 1 for c in calls)
        minus += sum(c[6].size - 1 for c in calls if c[0] == 1)
        windows |= {(c[4], c[5]) for c in calls}
    assert total > 2000 and minus > 100 and len(windows) > 6
    calls, _ = H.load_rm_golden(H.RM_CASES_BY_NAME["rm_multi_iter"][0])
    assert max(int(c[6][0]["ref_start"]) for c in calls) > 3 * H.RM_CASES_BY_NAME["rm_multi_iter"][0].max_hits_override


def test_rm_interval_windows_follow_main_cpp():
    """repeat_masker_src/main.cpp:323-420 on a sequence of 10 intervals, neighbour proportion 0.3 ->
    3 neighbouring intervals: one on the left, one on the right of the query interval."""
    iv = H.rm_intervals(100_000, 19, 10_000, 0.3)
    assert len(iv) == 10 and iv[0] == (0, 10_000, 0, 30_000) and iv[1] == (10_000, 20_000, 0, 30_000)
    assert iv[4] == (40_000, 50_000, 30_000, 60_000)
    assert iv[9] == (90_000, 99_981, 70_000, 100_000)


def test_rm_sort_dedupe_chain():
    """repeat_masker_src/seed_filter.cu:819-835 on hand-made records."""
    from oracle import sa_oracle_py as sao
    S = H.SEGMENT_DTYPE
    rec = np.array([(100, 50, 30, 4000), (100, 50, 30, 4000),      # exact copies: one survives :821
                    (105, 55, 10, 3100),                            # same diagonal, contained in the first: dropped :827
                    (100, 50, 20, 3500),                            # same start, shorter: contained, dropped
                    (300, 50, 30, 4200), (200, 50, 30, 4200),      # same query_start + score: ref_start DESC in the final order
                    (10, 5, 40, 9000)], dtype=S)
    out = sao.rm_sort_dedupe(rec)
    assert [tuple(r) for r in out] == [(10, 5, 40, 9000), (300, 50, 30, 4200), (200, 50, 30, 4200), (100, 50, 30, 4000)]
    # order of the input does not matter
    rng = np.random.default_rng(3)
    for _ in range(5):
        assert np.array_equal(sao.rm_sort_dedupe(rng.permutation(rec)), out)


def test_rm_revcomp_codes():
    from oracle import sa_oracle_py as sao
    enc = sao.encode(np.frombuffer(b"ACGTacgtNn&RA", dtype=np.uint8))
    assert list(sao.rm_revcomp_codes(enc)) == [3, 6, 7, 5, 5, 4, 4, 4, 4, 0, 1, 2, 3]


# ------------------------------------------------------------------------------ GPU
@needs_golden
@pytest.mark.gpu
@pytest.mark.parametrize("case,prop", H.RM_CASES, ids=RM_IDS)
@pytest.mark.parametrize("device_seeding", [False, True], ids=["vector", "range"])
def test_rm_backend_matches_reference_golden(backend, case, prop, device_seeding):
    want, digest = H.load_rm_golden(case)
    seq, _ = case.inputs()
    assert H.inputs_digest(seq, seq[:0]) == digest
    got = H.run_rm_backend(backend, case, prop, seq, device_seeding=device_seeding)
    H.assert_rm_calls_equal(got, want, "rm backend vs reference golden")


@needs_golden
@pytest.mark.gpu
@pytest.mark.parametrize("case,prop", H.RM_CASES, ids=RM_IDS)
@pytest.mark.parametrize("knob", ["SEGALIGN_B200_FUSED=0", "SEGALIGN_B200_FILTER=0", "SEGALIGN_B200_FILTER_KERNEL=2",
                                  "SEGALIGN_B200_MERGE_MIN=8", "SEGALIGN_B200_MERGE_MIN=0", "SEGALIGN_B200_DEDUP=0",
                                  "SEGALIGN_B200_FINALIZE_CAP=8"])
def test_rm_backend_code_paths_match_reference_golden(backend, case, prop, knob, monkeypatch):
    """General (materialised hit list) path, exact stage alone, tile-walk-only filter, merge pass forced
    / disabled, no duplicate table, device-wide radix sorts instead of the one-block finalisation: all must
    give the reference's bytes."""
    k, v = knob.split("=")
    monkeypatch.setenv(k, v)
    want, _ = H.load_rm_golden(case)
    got = H.run_rm_backend(backend, case, prop, device_seeding=True)
    H.assert_rm_calls_equal(got, want, f"rm backend with {knob} vs reference golden")


@pytest.mark.gpu
@pytest.mark.parametrize("seed", [5, 6])
def test_rm_backend_matches_cpu_oracle_fresh_inputs(backend, seed):
    case = H.Case(f"rm_fresh_{seed}", "repeats", dict(n=22_000, copies=20, low_complexity=8), rng_seed=900 + seed,
                  lastz_interval=7_000, wga_chunk=3_000)
    seq, _ = case.inputs()
    want = H.run_rm_cpu_oracle(case, 0.5, seq, max_hits_device=748058112)
    got = H.run_rm_backend(backend, case, 0.5, seq, device_seeding=bool(seed & 1))
    H.assert_rm_calls_equal(got, want, "rm backend vs cpu oracle")


def gen_rm_n_runs(rng, n=40_000):
    """A repeat family with runs of N (70 bases to 6 kb) right behind some of its copies and elsewhere: off-diagonal
    walks cross a run at 0 per cell under --ambiguous=iupac (N opposite a base of another copy), on the main diagonal N
    meets N.  What stage B's zero-run planes must skip without changing a record."""
    seq, _ = H.gen_repeats(rng, n=n, copies=40, elem=400, elem_div=0.06, low_complexity=6)
    for s0, l in [(3_000, 70), (7_000, 150), (12_000, 1_100), (20_000, 6_000), (31_000, 2_500), (n - 1_500, 1_500)]:
        seq[s0:s0 + l] = ord("N")
    return seq, seq.copy()


H.GENERATORS.update(rm_n_runs=gen_rm_n_runs)


@pytest.mark.gpu
@pytest.mark.parametrize("knob", ["", "SEGALIGN_B200_WIDE=0", "SEGALIGN_B200_ZERO_RUNS=0"], ids=["default", "lane_pair", "no_skip"])
def test_rm_backend_matches_cpu_oracle_across_n_runs(built, knob, monkeypatch):
    from segalign_b200.backend import Backend
    if knob:
        k, v = knob.split("=")
        monkeypatch.setenv(k, v)
    case = H.Case("rm_n_runs", "rm_n_runs", rng_seed=77, ambiguous="iupac", lastz_interval=9_000, wga_chunk=4_000)
    seq, _ = case.inputs()
    want = H.run_rm_cpu_oracle(case, 0.6, seq, max_hits_device=748058112)
    assert sum(w[6].size - 1 for w in want) > 200
    for device_seeding in (False, True):
        be = Backend()
        be.InitializeInterface(1)
        got = H.run_rm_backend(be, case, 0.6, seq, device_seeding=device_seeding)
        H.assert_rm_calls_equal(got, want, f"rm backend [{knob or 'default'}] vs cpu oracle across runs of N")


@needs_golden
@pytest.mark.gpu
@pytest.mark.skipif(not H.RM_NEW_RUNNER.exists(), reason="oracle/_ref/rm_new_runner not built (needs /root/reference at build time)")
@pytest.mark.parametrize("name", ["rm_repeats", "rm_masked_multichrom", "rm_multi_iter"])
def test_rm_dropin_runner_with_reference_host_objects(name, tmp_path, built):
    """The reference's own ntcoding.o / DRAM.o and the repeat masker's g_* symbols (shim.cpp
    -DSA_SHIM_REPEAT_MASKER) on libsegalign_b200.so, driven like repeat_masker_src/main.cpp + seeder.cpp."""
    from tests.golden.make_golden_rm import dump_as_calls
    case, prop = H.RM_CASES_BY_NAME[name]
    want, _ = H.load_rm_golden(case)
    dump = H.run_rm_runner(H.RM_NEW_RUNNER, case, prop, tmp_path)
    H.assert_rm_calls_equal(dump_as_calls(dump), want, "rm drop-in runner vs reference golden")


@pytest.mark.gpu
def test_rm_state_errors(backend):
    from segalign_b200.backend import BackendError
    case, prop = H.RM_CASES[0]
    seq, _ = case.inputs()
    with pytest.raises(BackendError):
        backend.RmSendQueryWriteRequest()          # no block on the device yet
    H.setup_rm_backend(backend, case, seq)
    with pytest.raises(BackendError):
        backend.RmSendQueryWriteRequest()          # slot not cleared
    with pytest.raises(BackendError):
        backend.RmSeedAndFilterRange(0, 100, True, False, 500, 10)   # ref_end < ref_start
    backend.RmClearQuery()
    with pytest.raises(BackendError):
        backend.RmSeedAndFilterRange(0, 100, True, False, 0, 1000)   # no reverse complement resident
