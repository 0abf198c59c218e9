"""The tile-composition rule behind k_extend_wide (segalign_b200/csrc/kernels_extend_wide.cuh).

The reference walks an ungapped extension cell by cell (src/seed_filter.cu:278-652):
running sum s, running maximum M (first position on ties, strict >), stop at the first cell with
M - s > xdrop.  The kernel instead summarises 32-cell tiles -- sum, maxpre (+ first position),
minpre, drop -- and combines 32 of them with two scans; a walk entering a tile with (s, M) stops
inside it iff max(M - s - minpre, drop) > xdrop.  This test restates both in numpy and checks, on
random score streams with the reference's score alphabet, that the composed walk returns the same
(score, position) as the cell-by-cell walk -- including with the kernel's per-4-cell-group
OVER-estimate of `drop`, which may only add sequential tiles, never change the result."""
import numpy as np
import pytest


def walk_cells(sc, xdrop, s=0, M=0, mp=-1, base=0):
    for j, v in enumerate(sc):
        s += int(v)
        if s > M:
            M, mp = s, base + j
        if M - s > xdrop:
            return s, M, mp, True
    return s, M, mp, False


def tile_summary(sc, group_bound):
    ls = np.cumsum(sc)
    lrm = np.maximum.accumulate(ls)
    if group_bound:  # the kernel: max(L_before_group, max of group) - min of group, per 4 cells
        drop, L = 0, -(1 << 29)
        for g in range(0, len(sc), 4):
            L = max(L, int(ls[g:g + 4].max()))
            drop = max(drop, L - int(ls[g:g + 4].min()))
    else:
        drop = int((lrm - ls).max())
    return int(ls[-1]), int(ls.max()), int(np.argmax(ls)), int(ls.min()), drop


def walk_tiles(sc, xdrop, group_bound):
    n_tiles = len(sc) // 32
    s, M, mp, t, seq_tiles = 0, 0, -1, 0, 0
    while t < n_tiles:
        batch = range(t, min(t + 32, n_tiles))
        summ = [tile_summary(sc[32 * i:32 * i + 32], group_bound) for i in batch]
        s_in, m_in, T = s, M, None
        entry = []
        for k, (sm, mx, ap, mn, dr) in enumerate(summ):  # what the two warp scans compute
            entry.append((s_in, m_in))
            if max(m_in - s_in - mn, dr) > xdrop:
                T = k
                break
            if s_in + mx > m_in:
                m_in, mp = s_in + mx, 32 * (t + k) + ap
            s_in += sm
        if T is None:
            s, M, t = s_in, m_in, t + len(summ)
            continue
        s, M = entry[T]
        seq_tiles += 1
        s, M, mp, stop = walk_cells(sc[32 * (t + T):32 * (t + T) + 32], xdrop, s, M, mp, 32 * (t + T))
        if stop:
            return M, mp, seq_tiles
        t += T + 1
    return M, mp, seq_tiles


@pytest.mark.parametrize("p_match,xdrop", [(0.25, 910), (0.55, 910), (0.75, 910), (0.9, 300), (0.6, 2000)])
def test_composed_walk_equals_cell_walk(p_match, xdrop):
    rng = np.random.default_rng(int(p_match * 100) + xdrop)
    false_alarms = 0
    for trial in range(300):
        n = 32 * int(rng.integers(1, 120))
        match = rng.random(n) < p_match
        sc = np.where(match, rng.choice([91, 100], n), rng.choice([-31, -114, -123, -125], n, p=[0.34, 0.33, 0.16, 0.17]))
        if trial % 7 == 0:  # long flat stretches (N runs under --ambiguous score 0) and a terminator
            a = int(rng.integers(0, n))
            sc[a:a + int(rng.integers(1, 400))] = 0
        if trial % 5 == 0:
            sc[int(rng.integers(0, n))] = -1000
        _, M, mp, _ = walk_cells(sc, xdrop)
        exact = walk_tiles(sc, xdrop, group_bound=False)
        kernel = walk_tiles(sc, xdrop, group_bound=True)
        assert exact[:2] == (M, mp), (trial, exact, M, mp)
        assert kernel[:2] == (M, mp), (trial, kernel, M, mp)
        assert kernel[2] >= exact[2]
        false_alarms += kernel[2] - exact[2]
    assert false_alarms >= 0
