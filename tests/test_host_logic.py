"""Host-side logic: blocking / interval / chunk rules, host seed words, sharding."""
import numpy as np
import pytest

from oracle import sa_oracle_py as sao
from segalign_b200 import genome, sharding
from segalign_b200.backend import shape_pattern

A = lambda b: np.frombuffer(b, dtype=np.uint8).copy()  # noqa: E731


def test_make_blocks_rules():
    # src/main.cpp:359-415: '&' between chromosomes, block closes after the chromosome that
    # pushes it past the limit (no trailing '&'), last block drops its trailing '&'
    chroms = [A(b"A" * 6), A(b"C" * 6), A(b"G" * 3), A(b"T" * 2)]
    blocks = genome.make_blocks(chroms, block_size=10)
    assert [bytes(b) for b in blocks] == [b"AAAAAA&CCCCCC", b"GGG&TT"]
    assert [bytes(b) for b in genome.make_blocks(chroms[:1], block_size=10)] == [b"AAAAAA"]


def test_interval_and_chunk_lists():
    # main.cpp:380-393: [0, len - seed) exclusive; seeder.cpp:33-34 minus-strand mirror
    iv = genome.interval_list(1019, 19, interval=400)
    assert iv == [(0, 400), (400, 800), (800, 1000)]
    calls = genome.chunk_list(1019, 19, "both", interval=400, chunk=250)
    assert calls[:2] == [(0, 0, 250), (0, 250, 400)]
    assert calls[2:4] == [(1, 600, 850), (1, 850, 1000)]  # rc of (0,400) with q_block_len 1000
    assert sum(j1 - j0 for r, j0, j1 in calls if r == 0) == 1000
    assert sum(j1 - j0 for r, j0, j1 in calls if r == 1) == 1000
    assert all(r == 1 for r, _, _ in genome.chunk_list(1019, 19, "minus", 400, 250))


@pytest.mark.parametrize("shape,transition", [("12of19", True), ("14of22", False), ("1101011", True)])
def test_host_seed_words_match_oracle(shape, transition):
    rng = np.random.default_rng(9)
    seq = genome.soft_mask(genome.random_genome(20000, rng), 0.1, rng)
    seq = genome.insert_runs(seq, b"N", 3, 50, rng)
    seq[777] = ord("&")
    sh = sao.Shape(shape)
    j1 = seq.size - sh.span
    a = sh.chunk_seeds(seq, 100, j1, transition)
    b = genome.chunk_seeds(seq, 100, j1, shape_pattern(shape), transition)
    assert np.array_equal(a, b) and a.size > 0


@pytest.mark.parametrize("shape,transition", [("12of19", True), ("12of19", False), ("14of22", True), ("1101011", True),
                                              ("1" * 14 + "0" * 17 + "1", True)])
def test_native_host_seeding_matches_oracle(built, shape, transition):
    """sa_host_chunk_seeds (the C ABI's host seeding loop, BMI2 rolling-window path where the CPU has it)
    against the oracle's restatement of src/seeder.cpp:57-74 + ntcoding.cpp:43-61: same words, same order,
    for ranges that start inside masked stretches, cross N runs and a separator, and tiny ranges."""
    from segalign_b200.backend import Backend
    rng = np.random.default_rng(11)
    seq = genome.soft_mask(genome.random_genome(30000, rng), 0.12, rng, mean_run=40)
    seq = genome.insert_runs(seq, b"N", 4, 30, rng)
    seq[12345] = ord("&")
    seq[:3] = A(b"acg")
    sh = sao.Shape(shape)
    be = Backend()
    assert be.GenerateShapePos(shape) == sh.weight
    last = seq.size - sh.span
    for j0, j1 in [(0, last), (1, 2), (5000, 5001), (777, 12400), (12300, 12400), (last - 1, last), (29000, last)]:
        want = sh.chunk_seeds(seq, j0, j1, transition)
        got = be.host_chunk_seeds(seq, j0, j1, transition)
        assert np.array_equal(got, want), (shape, transition, j0, j1)


def test_revcomp_matches_reference_alphabet():
    s = A(b"ACGTacgtNn&")
    assert bytes(genome.revcomp_ascii(s)) == b"&nNacgtACGT"
    assert np.array_equal(genome.revcomp_ascii(s), sao.revcomp_ascii(s))


def test_mutate_rate_and_determinism():
    rng = np.random.default_rng(3)
    g = genome.random_genome(200000, rng)
    m1 = genome.mutate(g, 0.4, np.random.default_rng(4))
    m2 = genome.mutate(g, 0.4, np.random.default_rng(4))
    assert np.array_equal(m1, m2)
    assert abs((m1 != g).mean() - 0.4) < 0.01
    assert set(np.unique(m1)) <= set(b"ACGT")


def test_shard_units_partition():
    for n in (0, 1, 7, 800, 801):
        for w in (1, 2, 3, 8):
            parts = [sharding.shard_units(n, r, w) for r in range(w)]
            flat = [i for p in parts for i in p]
            assert flat == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    with pytest.raises(ValueError):
        sharding.shard_units(10, 2, 2)


def test_shard_intervals_round_robin():
    iv = genome.interval_list(95_000_019, 19)
    parts = [sharding.shard_intervals(iv, r, 4) for r in range(4)]
    assert sorted(x for p in parts for x in p) == iv
    assert [len(p) for p in parts] == [3, 3, 2, 2]


def test_bench_matrices_match_the_reference_construction():
    """bench.py restates src/main.cpp:187-268 for its own runs (it may not import the oracle on the product
    legs); both matrices must equal the checker's."""
    import bench
    from oracle import sa_oracle_py as sao
    assert np.array_equal(np.asarray(sao.build_matrix("", bench.XDROP)).reshape(64), bench.default_matrix())
    assert np.array_equal(np.asarray(sao.build_matrix("iupac", bench.XDROP)).reshape(64), bench.iupac_matrix())


def test_strong_sharding_covers_every_call_once():
    """bench.py --strong: the ranks' static shares of one query block's SeedAndFilter calls."""
    from segalign_b200 import sharding
    units = genome.chunk_list(100_000_000, 19, "both")
    for world in (2, 3, 8):
        got = [u for r in range(world) for u in sharding.shard_units(len(units), r, world)]
        assert got == list(range(len(units)))
