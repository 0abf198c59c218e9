"""SURVEY 8 f2, pinned against the reference itself: sa_write_segments must write, byte for byte, what
the UNMODIFIED src/segment_printer.cpp writes for the same HSP lists, chromosome tables and block /
interval coordinates (oracle/_ref/segprint_runner = that file compiled where it lies + a driver that
feeds it one printer_input through a real tbb::flow printer_node).  HSPs: golden dumps of the
reference kernels for the multi-chromosome and repeat cases.  CPU only."""
import numpy as np
import pytest

from segalign_b200 import genome
from segalign_b200.backend import Backend, SaChromTable
from tests import harness as H

pytestmark = pytest.mark.skipif(not H.SEGPRINT_RUNNER.exists(),
                                reason="oracle/_ref/segprint_runner not built (needs /root/reference at build time)")


def _split(block, pieces):
    """Pretend the block is made of `pieces` chromosomes of equal size (tables only; no '&' needed:
    the printer works on coordinates)."""
    n = block.size
    cuts = [n * i // pieces for i in range(pieces)]
    return cuts, [(cuts[i + 1] if i + 1 < pieces else n) - cuts[i] for i in range(pieces)]


@pytest.mark.parametrize("name,r_off,q_off", [("masked_multichrom", 0, 0), ("masked_multichrom", 1_000_003, 777),
                                              ("repeats_entropy", 0, 123_456), ("diverged_default", 3_000_000_000, 0)])
def test_write_segments_matches_reference_printer(built, tmp_path, name, r_off, q_off):
    case = H.CASES_BY_NAME[name]
    ref, query = case.inputs()
    calls, digest = H.load_golden(case)
    assert digest == H.inputs_digest(ref, query)
    fw = np.concatenate([c[6] for c in calls if c[0] == 0])   # seeder.cpp:79-83: chunk order
    rc = np.concatenate([c[6] for c in calls if c[0] == 1])
    assert fw.size > 0
    # chromosome tables: the block's real '&' pieces where it has them, else an artificial split; a leading
    # chromosome of another block in front when the block does not start at buffer offset 0
    def tables(block, prefix, off):
        if (block == ord("&")).any():
            (n, s, l), (rn, rs, rl) = genome.block_tables(block, prefix, off)
        else:
            cuts, lens = _split(block, 3)
            n, s, l = [f"{prefix}{i}" for i in range(3)], [c + off for c in cuts], lens
            order = range(2, -1, -1)
            rn = [n[i] for i in order]
            rs = [2 * off + block.size - s[i] - l[i] for i in order]
            rl = [l[i] for i in order]
        if off:
            n, s, l = [prefix + "_prev"] + list(n), [0] + list(s), [off] + list(l)
            rn, rs, rl = [prefix + "_prev"] + list(rn), [0] + list(rs), [off] + list(rl)
        return (list(n), list(s), list(l)), (list(rn), list(rs), list(rl))
    r_fwd, _ = tables(ref, "chrR", r_off)
    q_fwd, q_rc = tables(query, "chrQ", q_off)
    span = 19
    q_len = query.size - span
    files, cmds = H.run_reference_printer(tmp_path, r_fwd, q_fwd, q_rc, (2, 1, r_off, q_off, ref.size, q_len),
                                          (0, q_len, 7), fw, rc, data_folder="/d/", ambiguous="iupac", notrivial=True)
    strands = [(False, fw, q_fwd)] + ([(True, rc, q_rc)] if rc.size else [])
    assert sorted(files) == sorted(f"tmp7.block1.r{r_off}.{'minus' if m else 'plus'}.segments" for m, _, _ in strands)
    be = Backend()
    rt = SaChromTable.build(*r_fwd)
    for minus, hsps, qt in strands:
        path = tmp_path / ("ours.minus" if minus else "ours.plus")
        be.write_segments(path, hsps, minus, r_off, q_off, rt, SaChromTable.build(*qt))
        want = files[f"tmp7.block1.r{r_off}.{'minus' if minus else 'plus'}.segments"]
        assert path.read_text() == want
        assert want.count("\n") == hsps.size
    # the command lines the reference prints for the two files (segment_printer.cpp:96-115, :151-170)
    assert len(cmds) == len(strands) and all(c.startswith("lastz /d/ref.2bit[nameparse=darkspace][multiple][subset=ref_block1.name] ") for c in cmds)
