"""The matrix-derived code classes of the filter stage and of stage B (segalign_b200/csrc/screen_bound.h):
terminator / soft codes (screen_terminator_codes) and flat / partner codes (zero_run_codes), checked against their
definitions on the reference's three stock matrices (src/main.cpp:187-268) and on hand-made ones."""
import ctypes
import subprocess
import tempfile
from pathlib import Path

import numpy as np
import pytest

from oracle import sa_oracle_py as sao

ROOT = Path(__file__).resolve().parent.parent
A, C, G, T, L, N, X, E = range(8)


@pytest.fixture(scope="module")
def classes():
    out = Path(tempfile.mkdtemp(prefix="code_classes_")) / "code_classes.so"
    subprocess.run(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-o", str(out),
                    str(ROOT / "tests/native/code_classes.cpp")], check=True)
    lib = ctypes.CDLL(str(out))

    def f(mat, xdrop):
        m = np.ascontiguousarray(np.asarray(mat, dtype=np.int32).reshape(64))
        res = (ctypes.c_uint32 * 4)()
        lib.sa_test_code_classes(m.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), int(xdrop), res)
        return dict(term=res[0], strict=res[1], flat=res[2], partners=res[3])
    return f


def bits(*codes):
    return sum(1 << c for c in codes)


def check_definitions(mat, xdrop, r):
    """The properties the kernels rely on, straight from the definitions."""
    m = np.asarray(mat).reshape(8, 8)
    term = [c for c in range(4, 8) if (r["term"] >> c) & 1]
    soft = [c for c in range(4, 8) if not (r["term"] >> c) & 1]
    hard = [0, 1, 2, 3] + term
    for c in term:      # a terminator stops the walk against every code that is not soft
        for d in hard:
            assert m[c, d] < -xdrop and m[d, c] < -xdrop, (c, d)
    for c in range(4, 8):   # strict: against everything
        every = all(m[c, d] < -xdrop and m[d, c] < -xdrop for d in range(8))
        assert bool((r["strict"] >> c) & 1) == every
    assert r["strict"] & ~r["term"] == 0
    flat = [c for c in range(8) if (r["flat"] >> c) & 1]
    partners = [c for c in range(8) if (r["partners"] >> c) & 1]
    assert all(c >= 4 for c in flat)
    for c in flat:          # (flat, partner) pairs score 0 in both orientations
        for d in partners:
            assert m[c, d] == 0 and m[d, c] == 0, (c, d)
    if flat:
        assert {0, 1, 2, 3} <= set(partners)
        assert all(m[d, d] != 0 for d in range(4))
    return term, soft, flat, partners


@pytest.mark.parametrize("amb,want", [
    ("", dict(term=bits(L, N, E), strict=bits(L, N, E), flat=0, partners=0)),
    ("n", dict(term=bits(L, E), strict=bits(E), flat=bits(N), partners=bits(A, C, G, T, L, N))),
    ("iupac", dict(term=bits(L, E), strict=bits(E), flat=bits(N, X), partners=bits(A, C, G, T, L, N, X))),
])
def test_stock_matrices(classes, amb, want):
    mat = np.asarray(sao.build_matrix(amb, 910)).reshape(64)
    r = classes(mat, 910)
    assert r == want
    check_definitions(mat, 910, r)


def test_hand_made_matrices(classes):
    rng = np.random.default_rng(3)
    base = np.asarray(sao.build_matrix("iupac", 910)).reshape(8, 8)
    for trial in range(300):
        m = base.copy()
        # perturb the non-ACGT rows / columns: zeros, mild penalties, hard penalties
        for _ in range(rng.integers(1, 12)):
            c, d = rng.integers(4, 8), rng.integers(0, 8)
            v = int(rng.choice([0, 0, -50, -100, -909, -910, -911, -1000, -9100, 5]))
            m[c, d] = v
            if rng.random() < 0.7:
                m[d, c] = v
        if trial % 50 == 0:
            m[0, 0] = 0   # an ACGT match that scores 0: no flat codes at all
        r = classes(m, 910)
        check_definitions(m, 910, r)
        if trial % 50 == 0:
            assert r["flat"] == 0 and r["partners"] == 0


def test_two_candidates_that_do_not_stop_each_other_are_both_soft(classes):
    m = np.asarray(sao.build_matrix("", 910)).reshape(8, 8).copy()
    m[L, N] = m[N, L] = -5          # lower case x N is mild although both are -1000 against ACGT
    r = classes(m, 910)
    assert not (r["term"] >> L) & 1 and not (r["term"] >> N) & 1 and (r["term"] >> E) & 1
    check_definitions(m, 910, r)
