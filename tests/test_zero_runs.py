"""Host-side property test of stage B's zero-run skipping (segalign_b200/csrc/zero_runs.h).

zero_tile / zero_jump are host+device code: k_extend_wide / k_extend_hits and this test compile the very same
functions.  tests/native/zero_runs_check.cpp builds the bit planes for random code sequences with runs of N / IUPAC codes
of 1 base to hundreds of kilobases (on and off the 32- and 1024-base grids, at both ends of a block, opposite lower case,
separators and other runs) and checks, for random positions and both directions, that a recognised tile is exactly a tile of
zero-scoring pairs, that a jump never skips a cell that is not one (cell-by-cell count), and that long stretches are
crossed piecewise rather than tile by tile.  Parity of the kernels that use them: tests/test_live_reference_gpu.py
(test_n_runs_match_live_reference_kernels) and tests/test_repeat_masker.py (across_n_runs)."""
import subprocess
import tempfile
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def checker():
    out = Path(tempfile.mkdtemp(prefix="zero_runs_check_")) / "zero_runs_check"
    subprocess.run(["g++", "-O2", "-std=c++17", "-o", str(out), str(ROOT / "tests/native/zero_runs_check.cpp")], check=True)
    return out


@pytest.mark.parametrize("seed", [1, 2, 3, 4])   # odd seeds: --ambiguous=iupac code sets, even: --ambiguous=n
def test_jumps_are_sound_and_useful(checker, seed):
    r = subprocess.run([str(checker), str(seed), "120000"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "violations=0" in r.stdout
    fields = dict(kv.split("=") for kv in r.stdout.split())
    # the inputs must exercise what is tested, or the test is vacuous
    assert int(fields["tiles_true"]) > 10000 and int(fields["jumps"]) > 10000 and int(fields["deep"]) > 1000, r.stdout
