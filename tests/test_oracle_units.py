"""T1: the CPU restatement (oracle/sa_oracle.c) on hand-checkable micro-cases."""
import numpy as np
import pytest

from oracle import sa_oracle_py as sao

A = lambda s: np.frombuffer(s.encode() if isinstance(s, str) else s, dtype=np.uint8).copy()  # noqa: E731


def default_params(**kw):
    args = dict(sub_mat=sao.build_matrix("", 910), xdrop=910, hspthresh=3000, noentropy=False,
                seed_size=19, max_hits=1 << 30)
    args.update(kw)
    return sao.make_params(**args)


def test_encode_alphabet():
    # common/seed_filter_interface.cu:28-45
    got = sao.encode(A("ACGTacgtnN&RYx-"))
    assert got.tolist() == [0, 1, 2, 3, 4, 4, 4, 4, 5, 5, 7, 6, 6, 6, 6]


def test_encode_rc():
    # src/seed_filter.cu:120-154: A<->T C<->G, L/N/E keep, others X; rc[len-1-i]
    fwd, rc = sao.encode_rc(A("AACGTan&R"))
    assert fwd.tolist() == [0, 0, 1, 2, 3, 4, 5, 7, 6]
    assert rc.tolist() == [6, 7, 5, 4, 0, 1, 2, 3, 3]


def test_matrix_default_and_iupac():
    # SURVEY App. C / src/main.cpp:187-268
    m = sao.build_matrix("", 910).reshape(8, 8)
    assert m[:4, :4].tolist() == [[91, -114, -31, -123], [-114, 100, -125, -31],
                                  [-31, -125, 100, -114], [-123, -31, -114, 91]]
    assert (m[4, :7] == -1000).all() and (m[:7, 4] == -1000).all()
    assert (m[5, :7] == -1000).all() and m[6, :4].tolist() == [-100] * 4 and m[6, 6] == -100
    assert m[6, 4] == -1000 and m[6, 5] == -1000
    assert (m[7, :] == -9100).all() and (m[:, 7] == -9100).all()
    mi = sao.build_matrix("iupac", 910).reshape(8, 8)
    assert (mi[5, :7] == 0).all() and (mi[:7, 5] == 0).all()
    assert (mi[6, :7] == 0).all() and (mi[:7, 6] == 0).all()
    assert mi[4, 4] == -1000 and mi[4, 0] == -1000 and (mi[7, :] == -9100).all()
    mn = sao.build_matrix("n", 500).reshape(8, 8)
    assert (mn[5, :6] == 0).all() and mn[6, 0] == -100 and mn[6, 5] == -1000 and mn[7, 7] == -5000
    mx = sao.build_matrix("iupac,10,20", 910).reshape(8, 8)
    # three fields: field[0] is "iupac", reward 10, penalty -20
    assert mx[5, 5] == 10 and mx[5, 0] == -20 and mx[6, 6] == 10 and mx[0, 6] == -20


def test_shape_and_kmer():
    sh = sao.Shape("12of19")
    assert (sh.weight, sh.span) == (12, 19)
    assert list(sh.c.shape_pos[:12]) == [0, 1, 2, 4, 7, 8, 11, 13, 15, 16, 17, 18]
    seq = A("ACGTACGTACGTACGTACGTAC")
    codes = {"A": 0, "C": 1, "G": 2, "T": 3}
    want = 0
    for p in sh.c.shape_pos[:12]:
        want = (want << 2) + codes[chr(seq[p])]
    assert sh.kmer_at(seq, 0) == want
    # a lowercase / N / & anywhere in the span (don't-care positions included) invalidates
    for bad in (b"a", b"N", b"&", b"R"):
        s2 = seq.copy(); s2[3] = bad[0]  # position 3 is a don't-care position
        assert sh.kmer_at(s2, 0) == 1 << 31
    assert sao.Shape("14of22").weight == 14 and sao.Shape("14of22").span == 22
    assert sao.Shape("1101").weight == 3 and sao.Shape("1x01").weight == 2


def test_chunk_seeds_order_and_transitions():
    # src/seeder.cpp:57-74: position ascending; exact word then t = 0..w-1 variants kmer^(2<<2t)
    sh = sao.Shape("12of19")
    rng = np.random.default_rng(1)
    seq = A(bytes(rng.choice(list(b"ACGT"), size=200).astype(np.uint8)))
    seq[50] = ord("N")
    seeds = sh.chunk_seeds(seq, 0, 150, True)
    pos = (seeds & 0xFFFFFFFF).astype(np.int64)
    assert (np.diff(pos) >= 0).all()
    valid = [j for j in range(150) if not (j <= 50 < j + 19)]
    assert seeds.size == len(valid) * 13
    k0 = int(seeds[0] >> 32)
    assert [int(s >> 32) for s in seeds[1:13]] == [k0 ^ (2 << (2 * t)) for t in range(12)]
    assert sh.chunk_seeds(seq, 0, 150, False).size == len(valid)


def test_table_offsets_step1_and_step3():
    # common/seed_pos_table.cu:58-64: step 1 indexes 1..len-span; position 0 is never indexed
    sh = sao.Shape("12of19")
    rng = np.random.default_rng(2)
    ref = A(bytes(rng.choice(list(b"ACGT"), size=5000).astype(np.uint8)))
    t = sao.Table(sh, ref, ref.size, 1)
    assert t.index.size == 4 ** 12 and t.pos.size == 5000 - 19
    assert t.pos.min() == 1 and t.pos.max() == 5000 - 19
    # bucket k holds exactly the positions whose word is k, ascending
    for p in (1, 77, 4981):
        k = sh.kmer_at(ref, p)
        s, e = (t.index[k - 1] if k else 0), t.index[k]
        assert p in t.pos[s:e].tolist()
    t3 = sao.Table(sh, ref, ref.size, 3)
    # offset = 20 % 3 = 2, start_offset = 1, num_steps = (5000-19+2)/3
    assert t3.pos.size == (5000 - 19 + 2) // 3 and sorted(t3.pos.tolist())[:3] == [1, 4, 7]


def _seqs(ref_s, qry_s):
    return sao.encode(A(ref_s)), sao.encode(A(qry_s))


def test_extend_perfect_match_bounds():
    # identical 200-mers, anchor mid-way: left runs to cell 0 (k = q0), right to the end
    rng = np.random.default_rng(3)
    s = bytes(rng.choice(list(b"ACGT"), size=200).astype(np.uint8))
    ref, qry = _seqs(s, s)
    p = default_params()
    ok, seg = sao.extend_hit(p, ref, qry, 100, 100)
    m = sao.build_matrix("", 910).reshape(8, 8)
    total = sum(int(m[c, c]) for c in ref)
    assert ok and seg["ref_start"] == 0 and seg["query_start"] == 0
    assert seg["len"] == 199  # bases - 1 (A.5)
    assert seg["score"] == total


def test_extend_stops_at_masked_and_separator():
    rng = np.random.default_rng(4)
    s = bytearray(rng.choice(list(b"ACGT"), size=300).astype(np.uint8).tobytes())
    r = bytearray(s); q = bytearray(s)
    r[250] = ord("&")   # chromosome separator right of the anchor: E row = -9100 < -xdrop
    q[40] = ord("a")    # soft-masked base left of the anchor: L row = -1000 < -xdrop
    ref, qry = _seqs(bytes(r), bytes(q))
    ok, seg = sao.extend_hit(default_params(), ref, qry, 150, 150)
    assert ok and seg["ref_start"] == 41 and seg["ref_start"] + seg["len"] == 249


def test_extend_ties_keep_earliest_max_and_xdrop_excludes():
    # strict '>' update: a later equal maximum does not move the end (A.5)
    m = np.zeros(64, dtype=np.int32).reshape(8, 8)
    m[:] = -50
    for c in range(4):
        m[c, c] = 50
    p = default_params(sub_mat=m, xdrop=120, hspthresh=100, noentropy=True)
    # right of anchor: 4 matches (+200), 1 mismatch (150), 1 match (200, tie, not taken), then mismatches
    ref_s = "AAAAAAAAAA" + "ACGT" + "A" + "C" + "AAAAAAAAAAAAAAAA"
    qry_s = "AAAAAAAAAA" + "ACGT" + "C" + "C" + "CCCCCCCCCCCCCCCC"
    ref, qry = _seqs(ref_s, qry_s)
    ok, seg = sao.extend_hit(p, ref, qry, 10, 10)
    assert ok
    # left: 10 matches (cells k=1..10) -> left_extent 10, score 500; right best at k=3 (score 200)
    assert seg["ref_start"] == 0 and seg["len"] == 3 + 10 and seg["score"] == 700


def test_extend_fails_below_threshold():
    rng = np.random.default_rng(5)
    ref = sao.encode(A(bytes(rng.choice(list(b"ACGT"), size=400).astype(np.uint8))))
    qry = sao.encode(A(bytes(rng.choice(list(b"ACGT"), size=400).astype(np.uint8))))
    ok, _ = sao.extend_hit(default_params(), ref, qry, 200, 200)
    assert not ok


def test_entropy_scales_low_complexity():
    # 40 x 'A' on both: score 3640 in [3000, 9000], counts = (40,0,0,0) -> entropy 0 -> fails;
    # with noentropy it passes (src/seed_filter.cu:608-649)
    s = "A" * 40
    ref, qry = _seqs(s, s)
    ok, _ = sao.extend_hit(default_params(), ref, qry, 20, 20)
    assert not ok
    ok, seg = sao.extend_hit(default_params(noentropy=True), ref, qry, 20, 20)
    assert ok and seg["score"] == 40 * 91 and seg["len"] == 39
    # two-letter alphabet: entropy = 0.5 exactly up to rounding -> score halves
    s2 = "AC" * 30
    ref, qry = _seqs(s2, s2)
    ok, seg = sao.extend_hit(default_params(hspthresh=2000), ref, qry, 30, 30)
    raw = 30 * 91 + 30 * 100
    assert ok and abs(seg["score"] - raw * 0.5) <= 1.0 and seg["len"] == 59


def test_sort_dedupe_semantics():
    # A.8: drop cur iff same diagonal as its SORTED predecessor and (same start or contained)
    S = lambda r, q, l, s: (r, q, l, s)  # noqa: E731
    segs = np.array([S(100, 50, 30, 4000),   # diag 50
                     S(100, 50, 10, 3500),   # same start, shorter: sorts first and is KEPT
                     S(105, 55, 20, 3600),   # contained in (100,30) -> dropped
                     S(140, 90, 10, 3100),   # same diag, not contained -> kept
                     S(10, 200, 5, 3000),    # diag wraps (u32): sorts last
                     S(300, 10, 7, 3300)], dtype=sao.SEGMENT_DTYPE)
    out = sao.sort_dedupe(segs)
    got = [tuple(int(x) for x in r) for r in out]
    # final order: (query_start, ref_start, len, score desc)
    assert got == [S(300, 10, 7, 3300), S(100, 50, 10, 3500), S(140, 90, 10, 3100), S(10, 200, 5, 3000)]


def test_dedupe_uses_input_predecessor_not_last_kept():
    # non-transitive predicate: c is compared with b (dropped), not with a
    a = (100, 100, 50, 5000)      # [100,150]
    b = (110, 110, 40, 4000)      # [110,150] contained in a -> dropped
    c = (120, 120, 35, 3500)      # [120,155] not contained in b, not containing b -> kept
    out = sao.sort_dedupe(np.array([c, b, a], dtype=sao.SEGMENT_DTYPE))
    assert [tuple(int(x) for x in r) for r in out] == [a, c]


def test_iteration_plan():
    # A.7: even when everything fits the call is split in two
    prefix = np.array([0, 3, 3, 7, 7, 10, 10, 10], dtype=np.uint32)
    assert sao.iteration_plan(prefix, 1 << 20).tolist() == [4, 7]
    # MAX_HITS = 4: num_iter = 10/4+2 = 4; lower_bound(4)=3 -> pos 2; limit = 3+4 = 7 -> lb=3 -> pos 2; ...
    plan = sao.iteration_plan(prefix, 4)
    assert plan[-1] == 7 and (np.diff(plan.astype(np.int64)) >= 0).all()
    assert sao.iteration_plan(np.zeros(5, dtype=np.uint32), 100).size == 0


def test_seed_and_filter_header_and_scope():
    rng = np.random.default_rng(6)
    base = rng.choice(list(b"ACGT"), size=3000).astype(np.uint8)
    ref = base.copy(); qry = base.copy()
    sh = sao.Shape("12of19")
    tab = sao.Table(sh, ref, ref.size, 1)
    p = default_params()
    seeds = sh.chunk_seeds(qry, 0, qry.size - 19, True)
    out = sao.seed_and_filter(p, tab, sao.encode(ref), sao.encode(qry), seeds)
    # every position except 0 hits the main diagonal; all collapse to one HSP per iteration
    assert out[0]["score"] >= qry.size - 19 - 1
    assert out[0]["len"] == out.size - 1 and 1 <= out.size - 1 <= 2
    assert all(int(s["ref_start"]) == int(s["query_start"]) == 0 and s["len"] == 2999 for s in out[1:])
