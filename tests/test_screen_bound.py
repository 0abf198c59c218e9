"""Host-side property test of the popcount screen (segalign_b200/csrc/screen_bound.h).

The screen is host+device code: the CUDA filter kernel (kernels_screen.cuh) and this test compile
the very same functions.  tests/native/screen_check.cpp checks on ~2 M anchors (random, planted
homologies at 0-45 % divergence, soft-masked / N runs, separators, IUPAC letters, block edges) that
a rejected anchor never reaches hspthresh under the oracle's exact extension
(oracle/sa_oracle.c:sao_extend_hit, restating src/seed_filter.cu:232-652) and that every "decided"
bound is >= the exact score.
"""
import subprocess
import tempfile
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def checker():
    from segalign_b200.build import build_oracle
    build_oracle()
    out = Path(tempfile.mkdtemp(prefix="screen_check_")) / "screen_check"
    subprocess.run(["g++", "-O2", "-std=c++17", "-o", str(out), str(ROOT / "tests/native/screen_check.cpp"),
                    f"-L{ROOT / 'oracle'}", "-lsa_oracle", f"-Wl,-rpath,{ROOT / 'oracle'}", "-lm"], check=True)
    return out


@pytest.mark.parametrize("seed,amb,xdrop,thresh", [
    (1, "", 910, 3000), (2, "", 910, 3000), (3, "iupac", 910, 3000), (4, "n", 910, 3000),
    (5, "", 300, 1000), (6, "", 2000, 5000), (7, "", 910, 2200),
])
def test_screen_never_rejects_an_hsp(checker, seed, amb, xdrop, thresh):
    r = subprocess.run([str(checker), str(seed), "300000", amb, str(xdrop), str(thresh)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "violations=0" in r.stdout
    # the screen must actually decide something on these inputs, or the test is vacuous
    rejected = int(r.stdout.split("rejected=")[1].split()[0])
    assert rejected > 10000, r.stdout
