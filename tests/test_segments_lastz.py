"""the segments wire format and the LASTZ hand-off ("still feeds LASTZ").

The HSPs come from the golden dump of the UNMODIFIED reference for the multi-chromosome case;
sa_write_segments (host-only C ABI) turns them into *.segments files exactly as
src/segment_printer.cpp does; LASTZ 1.04.17 (built from the reference's submodule by
oracle/Makefile) must accept them with --segments= and produce gapped alignments."""
import subprocess
from pathlib import Path

import numpy as np
import pytest

from segalign_b200 import genome
from segalign_b200.backend import Backend, SaChromTable
from tests import harness as H

LASTZ = H.ROOT / "oracle" / "_ref" / "lastz"


def _write_fasta(path, block, names_starts_lens, block_start=0):
    names, starts, lens = names_starts_lens
    with open(path, "wb") as f:
        for n, s, l in zip(names, starts, lens):
            f.write(b">" + n.encode() + b"\n" + block[s - block_start:s - block_start + l].tobytes() + b"\n")


@pytest.fixture(scope="module")
def case_data(built):
    case = H.CASES_BY_NAME["masked_multichrom"]
    ref, query = case.inputs()
    calls, digest = H.load_golden(case)
    assert digest == H.inputs_digest(ref, query)
    fw = np.concatenate([c[6] for c in calls if c[0] == 0])   # seeder.cpp:79-83: chunk order
    rc = np.concatenate([c[6] for c in calls if c[0] == 1])
    return case, ref, query, fw, rc


def test_segments_coordinates_and_format(case_data, tmp_path):
    case, ref, query, fw, rc = case_data
    be = Backend()
    (r_fwd, _), (q_fwd, q_rc) = genome.block_tables(ref, "chrR"), genome.block_tables(query, "chrQ")
    assert len(r_fwd[0]) == 4 and len(q_fwd[0]) == 4
    rt = SaChromTable.build(*r_fwd)
    sub = H.matrix_for(case).reshape(8, 8)
    from oracle import sa_oracle_py as sao
    ref_enc = sao.encode(ref)
    for minus, hsps, qt_raw, qblock in ((False, fw, q_fwd, query), (True, rc, q_rc, genome.revcomp_ascii(query))):
        path = tmp_path / ("minus.segments" if minus else "plus.segments")
        be.write_segments(path, hsps, minus, 0, 0, rt, SaChromTable.build(*qt_raw))
        lines = path.read_text().splitlines()
        assert len(lines) == hsps.size > 20
        q_enc = sao.encode(qblock)
        order = hsps[::-1] if minus else hsps
        for line, e in list(zip(lines, order)):
            n1, s1, e1, n2, s2, e2, strand, score = line.split("\t")
            assert strand == ("-" if minus else "+") and int(score) == e["score"]
            ri, qi = r_fwd[0].index(n1), qt_raw[0].index(n2)
            # origin-one closed intervals inside one chromosome on both sides
            assert 1 <= int(s1) <= int(e1) <= r_fwd[2][ri] and 1 <= int(s2) <= int(e2) <= qt_raw[2][qi]
            assert int(e1) - int(s1) == int(e2) - int(s2) == e["len"]
            r0 = r_fwd[1][ri] + int(s1) - 1
            q0 = qt_raw[1][qi] + int(s2) - 1
            assert (r0, q0) == (e["ref_start"], e["query_start"])
            n = e["len"] + 1
            raw = int(sub[ref_enc[r0:r0 + n], q_enc[q0:q0 + n]].sum())
            assert raw >= case.hspthresh and int(score) <= raw
    # minus-strand query names appear in query-FILE order (LASTZ requirement, segment.c:335-365)
    names = [l.split("\t")[3] for l in (tmp_path / "minus.segments").read_text().splitlines()]
    idx = [q_fwd[0].index(n) for n in names]
    assert idx == sorted(idx)


@pytest.mark.skipif(not LASTZ.exists(), reason="oracle/_ref/lastz not built (needs /root/reference at build time)")
def test_lastz_accepts_segments_and_extends_them(case_data, tmp_path):
    case, ref, query, fw, rc = case_data
    be = Backend()
    (r_fwd, _), (q_fwd, q_rc) = genome.block_tables(ref, "chrR"), genome.block_tables(query, "chrQ")
    _write_fasta(tmp_path / "ref.fa", ref, r_fwd)
    _write_fasta(tmp_path / "query.fa", query, q_fwd)
    rt = SaChromTable.build(*r_fwd)
    total = 0
    for strand, minus, hsps, qt in (("plus", False, fw, q_fwd), ("minus", True, rc, q_rc)):
        seg = tmp_path / f"{strand}.segments"
        be.write_segments(seg, hsps, minus, 0, 0, rt, SaChromTable.build(*qt))
        # the reference's command line (segment_printer.cpp:101-112) with FASTA instead of 2bit inputs
        out = tmp_path / f"{strand}.maf"
        p = subprocess.run([str(LASTZ), f"{tmp_path / 'ref.fa'}[multiple]", str(tmp_path / "query.fa"),
                            "--format=maf", "--ydrop=9430", "--gappedthresh=3000", f"--strand={strand}",
                            f"--segments={seg}", f"--output={out}"], capture_output=True, text=True)
        assert p.returncode == 0, p.stderr
        assert "FAILURE" not in p.stderr
        blocks = out.read_text().count("\na score=")
        assert blocks > 0
        total += blocks
    assert total >= 10  # gapped extension merges neighbouring HSPs into fewer blocks
