// zero_runs.h -- stage B's skipping of zero-scoring runs (host + device code; code sets: screen_bound.h zero_run_codes).
//
// Under --ambiguous=n|iupac N scores 0 against every code but a separator.  A tile of 32 such cell pairs leaves a walk
// of the reference's ungapped extension (src/seed_filter.cu:232-652) as it found it: the running sum and maximum
// stay, the position of the maximum stays (it moves on strict > only), the X-drop test cannot fire, and the entropy
// counters are not touched -- the tile lies behind the maximum, where :444-451 counts equal ACGT codes only, and those
// never score 0 (zero_run_codes checks it).  Per block, four bit planes say where such cells are:
//   f1 / g1   1 bit per base: the cell's code is flat (F) / a partner (G); 0 for padding and past the end
//   F1k / G1k 1 bit per aligned 1024 bases: all of them exist and are flat / partners
// zero_tile  : all 32 cell pairs of a tile are (F, G) pairs in one orientation or the other;
// zero_jump  : how many cells from the next one on lie in 1024-base pieces that are entirely flat on one block and
//              entirely partners on the other (up to 32 pieces = 32 768 cells per call).
// tests/native/zero_runs_check.cpp compiles this file for the host and checks both against a cell-by-cell count.
#pragma once
#include "screen_bound.h"

namespace sa {

struct ZeroPlanes {
    const uint32_t *f1, *g1, *F1k, *G1k;
};

SA_HD uint32_t zr_ld(const uint32_t *p) {
#if defined(__CUDA_ARCH__)
    return __ldg(p);
#else
    return *p;
#endif
}
SA_HD uint32_t zr_trailing_ones(uint32_t w) { // 32 for all-ones
#if defined(__CUDA_ARCH__)
    return w == 0xFFFFFFFFu ? 32u : (uint32_t)__ffs((int)~w) - 1u;
#else
    return w == 0xFFFFFFFFu ? 32u : (uint32_t)__builtin_ctz(~w);
#endif
}
SA_HD uint32_t zr_leading_ones(uint32_t w) { // 32 for all-ones
#if defined(__CUDA_ARCH__)
    return (uint32_t)__clz((int)~w);
#else
    return w == 0xFFFFFFFFu ? 32u : (uint32_t)__builtin_clz(~w);
#endif
}
// 32 bits of a plane starting at bit c (bit c in bit 0); reads words c >> 5 and (c >> 5) + 1
SA_HD uint32_t zr_window(const uint32_t *plane, uint32_t c) {
    const uint32_t w = c >> 5;
    return scr_funnel_r(zr_ld(plane + w), zr_ld(plane + w + 1), c & 31u);
}

// reference cells rc0 .. rc0+31 against query cells qc0 .. qc0+31 (the tile lies inside both blocks)
SA_HD bool zero_tile(const ZeroPlanes &R, const ZeroPlanes &Q, uint32_t rc0, uint32_t qc0) {
    const uint32_t fr = zr_window(R.f1, rc0), fq = zr_window(Q.f1, qc0);
    if ((fr | fq) != 0xFFFFFFFFu) return false; // the common case: two window loads
    const uint32_t gr = zr_window(R.g1, rc0), gq = zr_window(Q.g1, qc0);
    return ((fr & gq) | (fq & gr)) == 0xFFFFFFFFu;
}

// cells c, c+1, .. that lie in consecutive marked pieces of the coarse plane k1
SA_HD uint32_t coarse_span_up(const uint32_t *k1, uint32_t c) {
    const uint32_t n = zr_trailing_ones(zr_window(k1, c >> 10)); // pieces c >> 10 .. + 31
    return n ? n * 1024u - (c & 1023u) : 0u;
}
// cells top-1, top-2, .. that lie in consecutive marked pieces
SA_HD uint32_t coarse_span_down(const uint32_t *k1, uint32_t top) {
    if (top == 0) return 0u;
    const uint32_t b = (top - 1u) >> 10; // piece of the first cell to visit -> bit 31 of w
    const uint32_t w = b >= 31u ? zr_window(k1, b - 31u) : zr_ld(k1) << (31u - b);
    const uint32_t n = zr_leading_ones(w);
    return n ? (n - 1u) * 1024u + ((top - 1u) & 1023u) + 1u : 0u;
}

// Cells (a multiple of 32) a walk may skip.  right: the next cells are r, r+1, .. / q, q+1, ..; left: the next cells are
// r-1, r-2, .. / q-1, q-2, ...  All of them lie in pieces that are entirely flat on one block and entirely partners on
// the other -- hence inside both blocks, and every pair scores 0.
SA_HD uint32_t zero_jump(const ZeroPlanes &R, const ZeroPlanes &Q, uint32_t r, uint32_t q, bool left) {
    uint32_t k, k2;
    if (!left) {
        k = coarse_span_up(R.F1k, r); k2 = coarse_span_up(Q.G1k, q);
        k = k < k2 ? k : k2;
        if (k < 32u) {
            k = coarse_span_up(Q.F1k, q); k2 = coarse_span_up(R.G1k, r);
            k = k < k2 ? k : k2;
        }
    } else {
        k = coarse_span_down(R.F1k, r); k2 = coarse_span_down(Q.G1k, q);
        k = k < k2 ? k : k2;
        if (k < 32u) {
            k = coarse_span_down(Q.F1k, q); k2 = coarse_span_down(R.G1k, r);
            k = k < k2 ? k : k2;
        }
    }
    return k & ~31u;
}

} // namespace sa
