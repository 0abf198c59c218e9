"""the Boost-free whole-genome driver (sa_pipeline_run).

The expectation is built independently in Python: blocks, intervals, chunk order and chromosome
tables from segalign_b200/genome.py (restating src/main.cpp:336-415, :380-393 and src/seeder.cpp),
the HSPs of every (reference block, query block, chunk, strand) from the CPU oracle
(oracle/sa_oracle.c), the record text from the format contract of src/segment_printer.cpp:72-94 /
:125-149.  The driver must produce exactly those files -- same names, same bytes -- for a run with
3 reference blocks x 3 query blocks (tiny seq_block_size), several intervals per block and both
strands, i.e. the double-buffered schedule of src/main.cpp:600-741."""
import numpy as np
import pytest

from segalign_b200 import genome
from tests import harness as H


def _fasta(path, chroms, prefix):
    with open(path, "wb") as f:
        for i, c in enumerate(chroms):
            f.write(b">%s%d some description\n" % (prefix.encode(), i))
            for k in range(0, c.size, 70):
                f.write(c[k:k + 70].tobytes() + b"\n")


def _tables(blocks, names_by_block):
    """global chromosome tables (buffer offsets) over all blocks, forward and reverse-complement"""
    fwd, rc, starts, off, chrom = ([], [], []), ([], [], []), [], 0, 0
    for b, blk in enumerate(blocks):
        starts.append(off)
        (n, s, l), (rn, rs, rl) = genome.block_tables(blk, "x", off)
        names = names_by_block[b]
        fwd[0].extend(names); fwd[1].extend(s); fwd[2].extend(l)
        rc[0].extend(names[::-1]); rc[1].extend(rs); rc[2].extend(rl)
        off += blk.size
    return fwd, rc, starts


def _locate(starts, pos):
    return int(np.searchsorted(np.asarray(starts), pos, side="right") - 1)


def test_matrix_builder_matches_oracle(built):
    from oracle import sa_oracle_py as sao
    from segalign_b200.backend import Backend
    be = Backend()
    for amb in ("", "n", "iupac", "x,50,60"):
        for xdrop in (910, 300):
            assert np.array_equal(be.build_matrix(amb, xdrop), sao.build_matrix(amb, xdrop)), (amb, xdrop)


@pytest.mark.parametrize("block_size,interval", [(35_000, 16_000), (10_000, 5_000), (1_000_000, 30_000), (25_999, 100_000)])
def test_driver_plan_blocks_and_intervals_match_python_restatement(built, tmp_path, block_size, interval):
    """sa_pipeline_plan (host only, no GPU): FASTA parsing (wrapped lines, descriptions after the name,
    CRLF, blank lines), the block closure rule and the interval count against segalign_b200/genome.py
    (src/main.cpp:336-415, :380-393)."""
    from segalign_b200.backend import Backend
    rng = np.random.default_rng(block_size)
    ref_chroms = [genome.random_genome(n, rng) for n in (30_000, 9_000, 26_000, 41_000, 12_000, 7_000, 1)]
    query_chroms = [genome.random_genome(n, rng) for n in (20_000, 31_000, 26_000, 21_000, 8_000)]
    _fasta(tmp_path / "ref.fa", ref_chroms, "chrR")
    with open(tmp_path / "query.fa", "wb") as f:  # CRLF + blank lines + tab-separated description
        for i, c in enumerate(query_chroms):
            f.write(b">chrQ%d\tdesc\r\n\r\n" % i)
            for k in range(0, c.size, 61):
                f.write(c[k:k + 61].tobytes() + b"\r\n")
    out = tmp_path / "plan"
    out.mkdir()
    rep = Backend().pipeline_run(tmp_path / "ref.fa", tmp_path / "query.fa", out, plan_only=True,
                                 seq_block_size=block_size, lastz_interval=interval)
    r_blocks = genome.make_blocks(ref_chroms, block_size)
    q_blocks = genome.make_blocks(query_chroms, block_size)
    assert rep["ref_blocks"] == len(r_blocks) and rep["query_blocks"] == len(q_blocks)
    assert rep["intervals"] == len(r_blocks) * sum(len(genome.interval_list(b.size, 19, interval)) for b in q_blocks)
    k = 0
    for b, blk in enumerate(r_blocks):
        n = int((blk == ord("&")).sum()) + 1
        assert (out / f"ref_block{b}.name").read_text().split() == [f"chrR{i}" for i in range(k, k + n)]
        k += n
    k = 0
    for b, blk in enumerate(q_blocks):
        n = int((blk == ord("&")).sum()) + 1
        assert (out / f"query_block{b}.name").read_text().split() == [f"chrQ{i}" for i in range(k, k + n)]
        k += n
    assert not (out / "lastz_commands.txt").exists()


def test_driver_fails_loudly_without_a_gpu(built, tmp_path):
    """No CPU fallback anywhere: on a machine without a CUDA device the driver stops at
    InitializeInterface with the reference's "no GPU" error (exit code 1, seed_filter_interface.cu:54-57)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from segalign_b200.backend import Backend, BackendError
    rng = np.random.default_rng(1)
    _fasta(tmp_path / "ref.fa", [genome.random_genome(5000, rng)], "chrR")
    _fasta(tmp_path / "query.fa", [genome.random_genome(4000, rng)], "chrQ")
    with pytest.raises(BackendError) as ei:
        Backend().pipeline_run(tmp_path / "ref.fa", tmp_path / "query.fa", tmp_path, xdrop=910, hspthresh=3000, transition=1)
    assert ei.value.code == -1


@pytest.mark.gpu
def test_driver_reproduces_blocks_schedule_and_segment_files(backend, tmp_path):
    from oracle import sa_oracle_py as sao
    rng = np.random.default_rng(4242)
    ref_chroms = [genome.soft_mask(genome.random_genome(n, rng), 0.1, rng) for n in (30_000, 9_000, 26_000, 41_000, 12_000, 7_000)]
    # query: diverged copies of reference pieces (plus one inverted), so that both strands have HSPs
    q0 = genome.mutate(ref_chroms[0][:20_000], 0.2, rng)
    q1 = genome.revcomp_ascii(genome.mutate(ref_chroms[3][5_000:36_000], 0.18, rng))
    q2 = genome.mutate(ref_chroms[2], 0.25, rng)
    q3 = genome.mutate(np.concatenate([ref_chroms[4], ref_chroms[1]]), 0.15, rng)
    q4 = genome.random_genome(8_000, rng)
    query_chroms = [q0, q1, q2, q3, q4]
    _fasta(tmp_path / "ref.fa", ref_chroms, "chrR")
    _fasta(tmp_path / "query.fa", query_chroms, "chrQ")
    out = tmp_path / "out"
    out.mkdir()
    block_size, interval, chunk = 35_000, 16_000, 7_000
    rep = backend.pipeline_run(tmp_path / "ref.fa", tmp_path / "query.fa", out, seq_block_size=block_size,
                               lastz_interval=interval, wga_chunk=chunk, transition=1, xdrop=910, ydrop=9430,
                               hspthresh=3000, num_threads=4, data_folder="/data/", gapped=0, notrivial=1)
    # ---- independent expectation
    r_blocks = genome.make_blocks(ref_chroms, block_size)
    q_blocks = genome.make_blocks(query_chroms, block_size)
    assert len(r_blocks) == 3 and len(q_blocks) == 3
    assert rep["ref_blocks"] == 3 and rep["query_blocks"] == 3

    def names_by_block(blocks, prefix):
        res, k = [], 0
        for blk in blocks:
            n = int((blk == ord("&")).sum()) + 1
            res.append([f"{prefix}{i}" for i in range(k, k + n)])
            k += n
        return res
    rn, qn = names_by_block(r_blocks, "chrR"), names_by_block(q_blocks, "chrQ")
    (r_names, r_starts, _), _, r_bstart = _tables(r_blocks, rn)
    (q_names, q_starts, _), (rc_names, rc_starts, _), q_bstart = _tables(q_blocks, qn)
    for b in range(3):
        assert (out / f"ref_block{b}.name").read_text().split() == rn[b]
        assert (out / f"query_block{b}.name").read_text().split() == qn[b]
    shape = sao.Shape("12of19")
    sub = sao.build_matrix("", 910)
    params = sao.make_params(sub, 910, 3000, False, shape.span, 748058112)
    expected, n_intervals = {}, 0
    printer_files, printer_cmds = {}, []   # the same HSPs through the reference's own segment_printer.cpp
    use_printer = H.SEGPRINT_RUNNER.exists()
    (_, _, r_lens), _, _ = _tables(r_blocks, rn)
    (_, _, q_lens), (_, _, rc_lens), _ = _tables(q_blocks, qn)
    for rb, rblk in enumerate(r_blocks):
        table = sao.Table(shape, rblk, rblk.size, 1)
        ref_enc = sao.encode(rblk)
        for qb, qblk in enumerate(q_blocks):
            q_fwd, q_rc = sao.encode_rc(qblk)
            q_rc_ascii = sao.revcomp_ascii(qblk)
            q_block_len = qblk.size - shape.span
            for idx, (s, e) in enumerate(genome.interval_list(qblk.size, shape.span, interval), start=1):
                n_intervals += 1
                by_strand = []
                for rev, (lo, hi) in enumerate(((s, e), (q_block_len - e, q_block_len - s))):
                    hsps = []
                    for j0 in range(lo, hi, chunk):
                        seeds = shape.chunk_seeds(q_rc_ascii if rev else qblk, j0, min(j0 + chunk, hi), True)
                        if seeds.size:
                            hsps.append(sao.seed_and_filter(params, table, ref_enc, q_rc if rev else q_fwd, seeds)[1:])
                    hsps = np.concatenate(hsps) if hsps else np.empty(0, dtype=H.SEGMENT_DTYPE)
                    by_strand.append(hsps)
                    if rev == 1 and use_printer and (by_strand[0].size or by_strand[1].size):
                        files, cmds = H.run_reference_printer(
                            tmp_path / "printer", (r_names, r_starts, r_lens), (q_names, q_starts, q_lens),
                            (rc_names, rc_starts, rc_lens), (rb + 1, qb, r_bstart[rb], q_bstart[qb], rblk.size, q_block_len),
                            (s, e, idx), by_strand[0], by_strand[1], data_folder="/data/", notrivial=True)
                        printer_files.update(files)
                        printer_cmds.extend(cmds)
                    if hsps.size == 0:
                        continue
                    lines = []
                    for h in (hsps[::-1] if rev else hsps):
                        gr, gq = int(h["ref_start"]) + r_bstart[rb], int(h["query_start"]) + q_bstart[qb]
                        ri = _locate(r_starts, gr)
                        qi = _locate(rc_starts if rev else q_starts, gq)
                        qs = (rc_starts if rev else q_starts)[qi]
                        lines.append("\t".join(map(str, [r_names[ri], gr + 1 - r_starts[ri], gr + int(h["len"]) + 1 - r_starts[ri],
                                                         (rc_names if rev else q_names)[qi], gq + 1 - qs, gq + int(h["len"]) + 1 - qs,
                                                         "-" if rev else "+", int(h["score"])])) + "\n")
                    expected[f"tmp{idx}.block{qb}.r{r_bstart[rb]}.{'minus' if rev else 'plus'}.segments"] = "".join(lines)
    got = {p.name: p.read_text() for p in out.glob("*.segments")}
    assert sorted(got) == sorted(expected)
    assert any(".minus." in n for n in got) and any(".plus." in n for n in got) and len(got) >= 6
    for name in expected:
        assert got[name] == expected[name], name
    assert rep["intervals"] == n_intervals and rep["segment_files"] == len(expected)
    assert rep["hsps"] == sum(t.count("\n") for t in expected.values())
    # one LASTZ command per segments file, in the reference's form (segment_printer.cpp:101-112)
    cmds = (out / "lastz_commands.txt").read_text().splitlines()
    assert len(cmds) == len(expected)
    if use_printer:   # ... and both, byte for byte, what the reference's unmodified segment_printer.cpp emits
        assert printer_files == got
        assert sorted(printer_cmds) == sorted(cmds)
    for c in cmds:
        assert c.startswith("lastz /data/ref.2bit[nameparse=darkspace][multiple][subset=ref_block")
        assert " --format=maf- --ydrop=9430 --gappedthresh=3000 --strand=" in c and " --segments=tmp" in c
    # the command line front end (segalign_main.cpp) produces the same files
    import subprocess
    from segalign_b200.build import CLI
    out2 = tmp_path / "out_cli"
    out2.mkdir()
    backend.ShutdownProcessor()
    p = subprocess.run([str(CLI), str(tmp_path / "ref.fa"), str(tmp_path / "query.fa"), "/data", f"--out_dir={out2}",
                        f"--seq_block_size={block_size}", f"--lastz_interval={interval}", f"--wga_chunk={chunk}",
                        "--num_threads=3", "--nogapped"], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    assert {q.name: q.read_text() for q in out2.glob("*.segments")} == got
    assert "segment files %d" % len(got) in p.stderr
