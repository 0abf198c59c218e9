// screen_bound.h -- the bit-parallel score screen of the filter stage (host + device code).
//
// Stage A of the extension (kernels_filter.cuh) needs, per seed hit, an UPPER BOUND of the
// reference's ungapped X-drop score (src/seed_filter.cu:278-652).  The tile walk there
// computes a tight bound at ~14 instructions per cell-pair group; almost every hit is a random
// match whose walk dies within ~20 cells to the right and ~40 to the left.  The screen below
// decides those hits from popcounts alone:
//
//   * the window [anchor-96, anchor+64) of both sequences is cut into 16-cell blocks that start at
//     the anchor (one 32-bit word of 2-bit codes per block);
//   * per block, XOR of the two words classifies the 16 cells: match / transition / transversion;
//     two popcounts give the class counts (m, s, 16-m-s);
//   * every class has an upper and a lower score, taken from the ACGT x ACGT block of the matrix:
//     block sums Bhi >= true block sum >= Blo, prefix sums Phi_j / Plo_j at block ends;
//   * bound of the running maximum: Mhat = max(0, max_j (Phi_{j-1} + amax * m_j)), amax = largest
//     diagonal entry: a true prefix inside block j cannot exceed the prefix before the block
//     plus its matches (transitions and transversions must score <= 0, checked on the host);
//   * termination: if at some block end  max(0, max_{i<=j} Plo_i) - Phi_j > xdrop  the reference's
//     walk has stopped by that cell (its running max is >= every earlier true prefix >= Plo_i,
//     its running sum <= Phi_j); a terminator cell in the window (a code whose every matrix
//     entry is < -xdrop, or a cell past the end of a block) stops it as well.
//
// Mhat is taken over ALL blocks of the window, also those behind the proven stop: a superset of
// the prefixes the reference visits, hence still an upper bound.  A hit is rejected only if both
// directions are proven to stop inside the window, no soft cell (non-ACGT, non-terminator) lies in
// it, and MhatR + MhatL < hspthresh.  Everything else goes on to the tile walk.  Decisions are
// conservative by construction; tests/test_screen_bound.py checks them against the oracle's exact
// extension on random, homologous, masked and block-edge inputs using THIS code compiled for the
// host.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define SA_HD __host__ __device__ __forceinline__
#else
#define SA_HD static inline
#endif

namespace sa {

constexpr int SCREEN_RB = 4;  // right blocks: cells [0, 64) from the anchor
constexpr int SCREEN_LB = 6;  // left blocks: cells [-96, 0)
constexpr int SCREEN_RECS = 6; // records wr-3 .. wr+2 cover both windows for any anchor offset
constexpr int SCREEN_ROW_WORDS = 12; // aligned window of one sequence: 4 + 6 words of codes, flags, spare

enum { SCREEN_F_TR = 1, SCREEN_F_TL = 2, SCREEN_F_SOFT = 4 };

struct ScreenConsts {
    int enabled;   // matrix admits the screen (see screen_consts_from_matrix)
    int amax;      // max(0, largest diagonal entry)
    uint32_t whi;  // int8 x3: upper class scores {match, transition, transversion}
    uint32_t wlo;  // int8 x3: lower class scores
    int xdrop;
    int hspthresh;
    int kc1;       // 1 - 65536: packs the match count into the class-count word.  A kernel parameter (constant bank /
                   // uniform register operand): as a literal ptxas re-materialises it in a register for every block
};

struct ScreenRec { // = uint4 {p2 lo, p2 hi, terminator bits, soft bits} of 32 cells
    uint32_t x, y, z, w;
};

SA_HD uint32_t scr_funnel_r(uint32_t lo, uint32_t hi, uint32_t sh) {
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(lo, hi, sh);
#else
    sh &= 31u;
    return sh ? (lo >> sh) | (hi << (32u - sh)) : lo;
#endif
}
SA_HD int scr_popc(uint32_t v) {
#if defined(__CUDA_ARCH__)
    return __popc(v);
#else
    return __builtin_popcount(v);
#endif
}
SA_HD int scr_dp4a(uint32_t a, uint32_t b, int c) {
#if defined(__CUDA_ARCH__)
    return __dp4a((int)a, (int)b, c);
#else
    for (int i = 0; i < 4; i++) c += (int)(int8_t)(a >> (8 * i)) * (int)(int8_t)(b >> (8 * i));
    return c;
#endif
}

// Align the window of one sequence to the anchor.  a[0..5] = records w-3 .. w+2 where w = anchor>>5,
// sh = anchor & 31.  row[0..3] = right blocks (block j = cells anchor+16j ..), row[4..9] = left
// blocks (block j = cells anchor-16(j+1) .. anchor-16j-1), row[10] = flags.
SA_HD void screen_align(const ScreenRec (&a)[SCREEN_RECS], uint32_t sh, uint32_t (&row)[SCREEN_ROW_WORDS]) {
    const bool o = sh >= 16u;
    const uint32_t k = (2u * sh) & 31u;
    const uint32_t s[12] = {a[0].x, a[0].y, a[1].x, a[1].y, a[2].x, a[2].y, a[3].x, a[3].y, a[4].x, a[4].y, a[5].x, a[5].y};
    uint32_t t[11];
#pragma unroll
    for (int i = 0; i < 11; i++) t[i] = o ? s[i + 1] : s[i];
#pragma unroll
    for (int j = 0; j < SCREEN_RB; j++) row[j] = scr_funnel_r(t[6 + j], t[7 + j], k);
#pragma unroll
    for (int j = 0; j < SCREEN_LB; j++) row[SCREEN_RB + j] = scr_funnel_r(t[5 - j], t[6 - j], k);
    const uint32_t low = (1u << sh) - 1u; // cells of the anchor record below the anchor
    const uint32_t tr = (a[3].z >> sh) | a[4].z | (a[5].z & low);          // cells [0, 64)
    const uint32_t tl = (a[0].z >> sh) | a[1].z | a[2].z | (a[3].z & low); // cells [-96, 0)
    const uint32_t so = a[0].w | a[1].w | a[2].w | a[3].w | a[4].w | a[5].w; // superset of the window
    row[10] = (tr ? (uint32_t)SCREEN_F_TR : 0u) | (tl ? (uint32_t)SCREEN_F_TL : 0u) | (so ? (uint32_t)SCREEN_F_SOFT : 0u);
    row[11] = 0u;
}

// The same for the REFERENCE side of a hit from the bare 2-bit plane: a[0..5] = words w-3 .. w+2 of
// the p2 plane (32 cells per 64-bit word, non-ACGT cells stored as 0).  No terminator or soft bits
// travel with the window (48 bytes per hit instead of 96):
//   * terminator cells (and cells past the end of the block) may be ignored altogether.  The
//     reference's walk stops at the first such cell p (its score is < -xdrop, or the walk runs out of
//     the block: cells past the end score 0 and end the walk with their tile), so every cell it
//     scores lies before p, where the stored codes are the true ones: Mhat still bounds its running
//     maximum, and "stopped by the end of block j" is true for every block that reaches p whatever
//     the popcounts say.  Ignoring them only loses the shortcut "a terminator in the window proves the
//     stop"; the X-drop proof fires on the cells behind p as it does on random sequence;
//   * soft cells (non-ACGT, not a terminator: a real score the stored code 0 would misstate) must
//     send the hit to the tile walk: `soft` = any soft cell in records w-3 .. w+2, from the block's
//     one-bit-per-record map (a superset of the window).
SA_HD void screen_align_p2(const uint64_t (&a)[SCREEN_RECS], uint32_t sh, bool soft, uint32_t (&row)[SCREEN_ROW_WORDS]) {
    const bool o = sh >= 16u;
    const uint32_t k = (2u * sh) & 31u;
    const uint32_t s[12] = {(uint32_t)a[0], (uint32_t)(a[0] >> 32), (uint32_t)a[1], (uint32_t)(a[1] >> 32),
                            (uint32_t)a[2], (uint32_t)(a[2] >> 32), (uint32_t)a[3], (uint32_t)(a[3] >> 32),
                            (uint32_t)a[4], (uint32_t)(a[4] >> 32), (uint32_t)a[5], (uint32_t)(a[5] >> 32)};
    uint32_t t[11];
#pragma unroll
    for (int i = 0; i < 11; i++) t[i] = o ? s[i + 1] : s[i];
#pragma unroll
    for (int j = 0; j < SCREEN_RB; j++) row[j] = scr_funnel_r(t[6 + j], t[7 + j], k);
#pragma unroll
    for (int j = 0; j < SCREEN_LB; j++) row[SCREEN_RB + j] = scr_funnel_r(t[5 - j], t[6 - j], k);
    row[10] = soft ? (uint32_t)SCREEN_F_SOFT : 0u;
    row[11] = 0u;
}

// One direction: NB blocks in walk order.  Returns the bound of the running maximum; proven = the
// reference's walk stops inside the window by the X-drop rule.
template <int NB>
SA_HD int screen_walk(const uint32_t *r, const uint32_t *q, const ScreenConsts &C, bool &proven) {
    int phi = 0, plo = 0, lm = 0, mhat = 0;
    bool stop = false;
#pragma unroll
    for (int j = 0; j < NB; j++) {
        const uint32_t x = r[j] ^ q[j];
        const uint32_t t = x >> 1;
        const uint32_t mm = ~(x | t) & 0x55555555u; // cells with equal codes
        const uint32_t tt = (t & ~x) & 0x55555555u; // codes differ by 2: A<->G, C<->T
        const int m = scr_popc(mm), s = scr_popc(tt);
        // class counts as bytes {m, s, 16-m-s, 0}
        const uint32_t cnt = (uint32_t)(m * C.kc1 + s * (256 - 65536) + (16 << 16));
        const int cand = phi + C.amax * m;
        mhat = cand > mhat ? cand : mhat;
        phi = scr_dp4a(cnt, C.whi, phi);
        plo = scr_dp4a(cnt, C.wlo, plo);
        lm = plo > lm ? plo : lm;
        stop |= lm - phi > C.xdrop;
    }
    proven = stop;
    return mhat;
}

// The screen of one hit: rr / qr = aligned rows of the reference and the query.  true = the hit
// cannot be an HSP.
SA_HD bool screen_reject(const uint32_t (&rr)[SCREEN_ROW_WORDS], const uint32_t (&qr)[SCREEN_ROW_WORDS],
                         const ScreenConsts &C, int &bound, bool &decided) {
    bool pr, pl;
    const int mr = screen_walk<SCREEN_RB>(rr, qr, C, pr);
    const int ml = screen_walk<SCREEN_LB>(rr + SCREEN_RB, qr + SCREEN_RB, C, pl);
    const uint32_t f = rr[10] | qr[10];
    pr |= (f & SCREEN_F_TR) != 0;
    pl |= (f & SCREEN_F_TL) != 0;
    bound = mr + ml;
    decided = pr && pl && !(f & SCREEN_F_SOFT);
    return decided && bound < C.hspthresh;
}

// Class scores from the ACGT x ACGT block of sub_mat (row = reference code, 8 columns per row).
static inline ScreenConsts screen_consts_from_matrix(const int *sub_mat, int xdrop, int hspthresh) {
    int hi[3] = {-1000000, -1000000, -1000000}, lo[3] = {1000000, 1000000, 1000000};
    for (int a = 0; a < 4; a++)
        for (int b = 0; b < 4; b++) {
            const int c = (a ^ b) == 0 ? 0 : ((a ^ b) == 2 ? 1 : 2);
            const int v = sub_mat[a * 8 + b];
            if (v > hi[c]) hi[c] = v;
            if (v < lo[c]) lo[c] = v;
        }
    ScreenConsts C;
    C.enabled = 1;
    for (int c = 0; c < 3; c++)
        if (hi[c] > 127 || lo[c] < -128) C.enabled = 0;
    if (hi[1] > 0 || hi[2] > 0) C.enabled = 0; // only matches may raise a prefix sum
    if (xdrop < 0) C.enabled = 0;
    C.amax = hi[0] > 0 ? hi[0] : 0;
    C.whi = C.wlo = 0;
    if (C.enabled)
        for (int c = 0; c < 3; c++) {
            C.whi |= (uint32_t)(uint8_t)(int8_t)hi[c] << (8 * c);
            C.wlo |= (uint32_t)(uint8_t)(int8_t)lo[c] << (8 * c);
        }
    C.xdrop = xdrop;
    C.hspthresh = hspthresh;
    C.kc1 = 1 - 65536;
    return C;
}

// Which non-ACGT codes (4..7) the filter stage may treat as X-drop terminators; every other non-ACGT code is
// "soft".  The two classes are defined together:
//   * a hit with a soft cell anywhere in its examined range is never decided by the filter (screen: soft flag of
//     either window; tile walk: a soft cell at or before the first terminator cell makes the hit a survivor);
//   * so a terminator code only has to stop the reference's walk when the cell it is paired with is NOT soft:
//     code c is a terminator iff sub_mat[c][d] < -xdrop and sub_mat[d][c] < -xdrop for every d that is A/C/G/T or
//     itself a terminator (the running sum then falls more than xdrop below the running maximum at that cell).
// The strict rule "every entry of the code's row and column < -xdrop" (round 1) is the special case without soft
// codes.  It made lower case a soft code under --ambiguous=n|iupac (lower case x N scores 0 there although lower case
// x ACGT is -1000), and on a half soft-masked block nearly every hit then went to the tile walk.
// `strict` (optional): the codes that are terminators against everything.
static inline uint32_t screen_terminator_codes(const int *sub_mat, int xdrop, uint32_t *strict = nullptr) {
    uint32_t soft = 0, all = 0;
    for (int c = 4; c < 8; c++) {
        bool every = true;
        for (int d = 0; d < 8; d++) {
            const bool high = sub_mat[c * 8 + d] >= -xdrop || sub_mat[d * 8 + c] >= -xdrop;
            if (high) every = false;
            if (high && d < 4) soft |= 1u << c;
        }
        if (every) all |= 1u << c;
    }
    for (bool changed = true; changed;) { // two remaining candidates that do not stop each other: both soft
        changed = false;
        for (int c = 4; c < 8; c++)
            for (int d = 4; d < 8; d++) {
                if (((soft >> c) | (soft >> d)) & 1u) continue;
                if (sub_mat[c * 8 + d] >= -xdrop || sub_mat[d * 8 + c] >= -xdrop) { soft |= (1u << c) | (1u << d); changed = true; }
            }
    }
    if (strict) *strict = all;
    return ~soft & 0xF0u;
}

// Zero runs (stage B).  Under --ambiguous=n|iupac N scores 0 against everything but a block separator, so a walk that
// reaches a run of N (an 18 Mb centromere) passes through it cell by cell without its state changing: the running sum
// stays, the running maximum and its position stay (the maximum moves on strict > only), the X-drop test cannot fire,
// and the entropy counters are not touched -- every cell of such a tile lies behind the maximum, where the reference
// counts equal ACGT codes only (into count_del, src/seed_filter.cu:444-451), and two equal ACGT codes never score 0
// (checked below).  Two code sets make such cells recognisable from bit planes:
//   F ("flat"):     non-ACGT codes c with sub_mat[c][d] == 0 == sub_mat[d][c] for every d in A/C/G/T;
//   G ("partners"): codes d with sub_mat[c][d] == 0 == sub_mat[d][c] for every c in F (N itself under --ambiguous).
// A cell pair (F, G) in either orientation scores 0.  F is empty under the default matrix.
static inline void zero_run_codes(const int *sub_mat, uint32_t *flat, uint32_t *partners) {
    uint32_t F = 0, G = 0;
    bool diag_nonzero = true;
    for (int d = 0; d < 4; d++)
        if (sub_mat[d * 9] == 0) diag_nonzero = false;
    for (int c = 4; c < 8 && diag_nonzero; c++) {
        bool ok = true;
        for (int d = 0; d < 4; d++)
            if (sub_mat[c * 8 + d] != 0 || sub_mat[d * 8 + c] != 0) ok = false;
        if (ok) F |= 1u << c;
    }
    if (F)
        for (int d = 0; d < 8; d++) {
            bool ok = true;
            for (int c = 4; c < 8; c++)
                if (((F >> c) & 1u) && (sub_mat[c * 8 + d] != 0 || sub_mat[d * 8 + c] != 0)) ok = false;
            if (ok) G |= 1u << d;
        }
    *flat = F;
    *partners = G;
}

} // namespace sa
