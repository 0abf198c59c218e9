// kernels_extend_wide.cuh -- stage B, first pass: exact extension of the filter's survivors with one
// WARP per hit and one 32-cell tile per LANE.
//
// Same semantics as find_hsps (src/seed_filter.cu:232-652) and as k_extend_hits
// (kernels_extend.cuh), which walks a hit tile after tile with one lane per direction.  That walk
// is a chain of dependent loads: ~3 us per tile.  Survivors of closely related genomes sit inside
// HSPs that are kilobases long (every seed hit of a conserved run extends over the whole run), so
// the chain -- not the work -- set the duration of stage B (~0.35 ms per call, resident next to
// the filter kernel of the following call and displacing one of its three blocks per SM).
//
// The X-drop rule composes over tiles.  For a tile with local prefix sums ls_j (j = 0..31):
//     sum = ls_31, maxpre = max_j ls_j (first position argpos), minpre = min_j ls_j,
//     drop = max_j (max_{i<=j} ls_i - ls_j)
// and for a walk entering the tile with running sum s and running maximum M:
//     it stops inside the tile  <=>  max(M - s - minpre, drop) > xdrop
//     otherwise it leaves with  s + sum,  max(M, s + maxpre)   (position updated on strict >).
// So a warp takes 32 consecutive tiles of a direction at once: every lane summarises its tile,
// two warp scans give every tile its entry state, the first tile that stops the walk (or needs
// the cell-by-cell code: a block end) is finished by its own lane with the exact
// sequential tile code, and the walk either ends there or continues behind it.  `drop` is
// over-estimated per 4-cell group (never under-estimated): a false alarm only costs a sequential
// tile, never a different result.  The first-maximum tie rule (strict >) is kept by taking the
// FIRST tile / first cell that attains the final maximum.
//
// Hits that need the entropy factor (hspthresh <= score <= 3*hspthresh, :608) are handed to
// k_extend_hits, which reproduces the reference's counter arithmetic; everything else is decided
// here with the same finish_hit / duplicate table / append code.
#pragma once
#include "kernels_extend.cuh"

namespace sa {

constexpr int WIDE_THREADS = 128;
constexpr int WIDE_LUT_COLS = 4;
constexpr int WIDE_LUT_WORDS = 256 * WIDE_LUT_COLS;
constexpr uint32_t WIDE_K_MUL = 1u | (16u << 8);

struct WalkState {
    int s, M, mp;
};

// Summary of one clean tile (all 32 cells upper-case ACGT, inside both blocks).  rw / qw hold the
// cells in processing order (cell j at bits 2j).
__device__ __forceinline__ void wide_tile_summary(uint32_t lut_lane, uint64_t rw, uint64_t qw, int &sum,
                                                  int &maxpre, int &argpos, int &minpre, int &dropub) {
    const uint32_t rl = (uint32_t)rw, rh = (uint32_t)(rw >> 32), ql = (uint32_t)qw, qh = (uint32_t)(qw >> 32);
    int s = 0, L = -(1 << 29), mn_all = 1 << 29, dub = 0, ap = 0;
#pragma unroll
    for (int g = 0; g < 8; g++) {
        const uint32_t y = g < 4 ? __byte_perm(rl, ql, 0x0040 | (g & 3) << 8 | (4 + (g & 3)))
                                 : __byte_perm(rh, qh, 0x0040 | (g & 3) << 8 | (4 + (g & 3)));
        const uint32_t v = group_scores(lut_lane, WIDE_K_MUL, y);
        const int p1 = __dp4a((int)v, 0x00000001, s), p2 = __dp4a((int)v, 0x00000101, s);
        const int p3 = __dp4a((int)v, 0x00010101, s), p4 = __dp4a((int)v, 0x01010101, s);
        const int mx = max(__vimax3_s32(p1, p2, p3), p4);
        const int mn = min(__vimin3_s32(p1, p2, p3), p4);
        if (mx > L) ap = 4 * g + (p1 == mx ? 0 : (p2 == mx ? 1 : (p3 == mx ? 2 : 3)));
        L = max(L, mx);
        dub = max(dub, L - mn); // >= every (running local max - prefix) of this group
        mn_all = min(mn_all, mn);
        s = p4;
    }
    sum = s; maxpre = L; argpos = ap; minpre = mn_all; dropub = dub;
}

// One tile, cell by cell, exactly as the reference scores it (no entropy counters).  tt = cells of
// this direction in front of the tile.  Returns true if the walk stops in this tile.
__device__ __forceinline__ bool wide_exact_tile(const ExtendParams &P, const int *sub, const int *lut16,
                                                uint32_t r0, uint32_t q0, bool left, uint32_t tt, WalkState &W) {
    const int X = P.xdrop;
    const int base = left ? (int)tt + 1 : (int)tt;
    bool inside;
    uint32_t rc0, qc0;
    if (!left) {
        rc0 = r0 + tt; qc0 = q0 + tt;
        inside = ((unsigned long long)rc0 + 32ull <= P.ref_len) && ((unsigned long long)qc0 + 32ull <= P.query_len);
    } else {
        inside = (r0 >= tt + 32u) && (q0 >= tt + 32u) && (r0 - tt <= P.ref_len) && (q0 - tt <= P.query_len);
        rc0 = r0 - tt - 32u; qc0 = q0 - tt - 32u;
    }
    uint32_t m = 0xFFFFFFFFu;
    if (inside) m = load_m1_window(P.rm1, rc0) | load_m1_window(P.qm1, qc0);
    int s = W.s, M = W.M, mp = W.mp;
    bool stop = false;
    if (m == 0) {
        uint64_t rw = load_p2_window(P.rp2, rc0), qw = load_p2_window(P.qp2, qc0);
        if (left) { rw = reverse_fields32(rw); qw = reverse_fields32(qw); }
#pragma unroll 8
        for (int j = 0; j < 32; j++) {
            const int idx = (int)(((rw >> (2 * j)) & 3u) << 2 | ((qw >> (2 * j)) & 3u));
            s += lut16[idx];
            if (s > M) { M = s; mp = base + j; }
            if (M - s > X) { stop = true; break; }
        }
    } else {
        // non-ACGT cell or block end: 1 B/base codes, bounds per cell (:328-336 / :482); stop on
        // x-drop or when the tile's last cell is out of bounds (:420)
        bool xd = false, last_in = true;
#pragma unroll 4
        for (int j = 0; j < 32; j++) {
            const uint32_t k = (uint32_t)(base + j);
            const bool in = left ? (r0 >= k && q0 >= k)
                                 : ((unsigned long long)r0 + k < P.ref_len && (unsigned long long)q0 + k < P.query_len);
            const uint32_t rp = left ? r0 - k : r0 + k, qp = left ? q0 - k : q0 + k;
            int v = 0;
            if (in) v = sub[__ldg(P.rb8 + rp) * 8 + __ldg(P.qb8 + qp)];
            if (!xd) {
                s += v;
                if (s > M) { M = s; mp = base + j; }
                if (M - s > X) xd = true;
            }
            if (j == 31) last_in = in;
        }
        stop = xd || !last_in;
    }
    W.s = s; W.M = M; W.mp = mp;
    return stop;
}

// One direction of one hit, all lanes of the warp together.  Every lane returns the same result.
__device__ __forceinline__ DirResult wide_extend_dir(const ExtendParams &P, const int *sub, const int *lut16,
                                                     uint32_t lut_lane, uint32_t r0, uint32_t q0, bool left) {
    const uint32_t lane = threadIdx.x & 31u;
    const int X = P.xdrop;
    WalkState W;
    W.s = 0; W.M = 0; W.mp = left ? 0 : -1;
    uint32_t t = 0; // cells of this direction already walked
    for (;;) {
        if (P.zskip && (!left || (r0 >= t && q0 >= t))) {
            // the next cells lie in 1024-base pieces that are flat on one block and partners on the other (zero_runs.h):
            // nothing changes over them.  Warp-uniform.
            const uint32_t k = zero_jump(P.rz, P.qz, left ? r0 - t : r0 + t, left ? q0 - t : q0 + t, left);
            if (k >= 1024u) { t += k; continue; }
        }
        // ---- every lane summarises tile `lane` of the next 32
        const uint32_t tt = t + 32u * lane;
        bool inside;
        uint32_t rc0, qc0;
        if (!left) {
            rc0 = r0 + tt; qc0 = q0 + tt;
            inside = ((unsigned long long)r0 + tt + 32ull <= P.ref_len) && ((unsigned long long)q0 + tt + 32ull <= P.query_len);
        } else {
            inside = (r0 >= tt + 32u) && (q0 >= tt + 32u) && (r0 - tt <= P.ref_len) && (q0 - tt <= P.query_len);
            rc0 = r0 - tt - 32u; qc0 = q0 - tt - 32u;
        }
        uint32_t m = 0xFFFFFFFFu;
        if (inside) m = load_m1_window(P.rm1, rc0) | load_m1_window(P.qm1, qc0);
        const bool clean = m == 0;
        // tiles that touch a block end always go to the cell-by-cell code (:420 rule); tiles with non-ACGT
        // cells do so unless such cells can be walked through (N / lower case not being terminators)
        const bool zero = P.zskip && inside && !clean && zero_tile(P.rz, P.qz, rc0, qc0); // 32 scores of 0: the empty summary
        const bool summarised = clean || zero || (inside && P.soft_runs);
        int sum = 0, maxpre = zero ? 0 : -(1 << 29), argpos = 0, minpre = 0, dropub = 0;
        if (zero) {
        } else if (clean) {
            uint64_t rw = load_p2_window(P.rp2, rc0), qw = load_p2_window(P.qp2, qc0);
            if (left) { rw = reverse_fields32(rw); qw = reverse_fields32(qw); }
            wide_tile_summary(lut_lane, rw, qw, sum, maxpre, argpos, minpre, dropub);
        } else if (summarised) {
            // non-ACGT cells under a matrix that lets a walk pass through them (N / IUPAC runs score 0
            // with --ambiguous=n|iupac): the same summary from the 1 B/base codes, exact drop.  A run of
            // thousands of N is then 32 tiles per step like any other stretch; a terminator cell shows
            // up as a drop > xdrop (the rest of such a tile is irrelevant: it is walked cell by cell).
            const int base = left ? (int)tt + 1 : (int)tt;
            int s = 0, L = -(1 << 29), mn = 1 << 29, dub = 0, ap = 0;
#pragma unroll 4
            for (int j = 0; j < 32; j++) {
                const uint32_t k = (uint32_t)(base + j);
                const uint32_t rp = left ? r0 - k : r0 + k, qp = left ? q0 - k : q0 + k;
                s += sub[__ldg(P.rb8 + rp) * 8 + __ldg(P.qb8 + qp)];
                if (s > L) { L = s; ap = j; }
                mn = min(mn, s);
                dub = max(dub, L - s);
                if (dub > X) break;
            }
            sum = s; maxpre = L; argpos = ap; minpre = mn; dropub = dub;
        }
        // ---- entry state of every tile: exclusive scans over the lanes in front of it
        int s_in = sum;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int up = __shfl_up_sync(0xFFFFFFFFu, s_in, off);
            if (lane >= (uint32_t)off) s_in += up;
        }
        s_in = s_in - sum + W.s;              // running sum in front of this tile
        const int cand = s_in + maxpre;       // the running maximum this tile proposes
        int m_in = cand;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int up = __shfl_up_sync(0xFFFFFFFFu, m_in, off);
            if (lane >= (uint32_t)off) m_in = max(m_in, up);
        }
        const int m_out = max(m_in, W.M);     // inclusive
        m_in = __shfl_up_sync(0xFFFFFFFFu, m_out, 1);
        if (lane == 0) m_in = W.M;            // running maximum in front of this tile
        const bool flag = !summarised || max(m_in - s_in - minpre, dropub) > X;
        const unsigned fm = __ballot_sync(0xFFFFFFFFu, flag);
        const uint32_t T = fm ? (uint32_t)__ffs((int)fm) - 1u : 32u; // first tile that needs the cell-by-cell code
        // ---- state in front of tile T (behind tile 31 if no tile is flagged)
        const uint32_t src = T < 32u ? T : 31u;
        int s_T = __shfl_sync(0xFFFFFFFFu, T < 32u ? s_in : s_in + sum, src);
        int M_T = __shfl_sync(0xFFFFFFFFu, T < 32u ? m_in : m_out, src);
        if (M_T > W.M) { // the maximum moved: first tile in front of T that attains it, first cell inside
            const unsigned am = __ballot_sync(0xFFFFFFFFu, cand == M_T) & (T < 32u ? ((1u << T) - 1u) : 0xFFFFFFFFu);
            const uint32_t ta = (uint32_t)__ffs((int)am) - 1u;
            const int ap = __shfl_sync(0xFFFFFFFFu, argpos, ta);
            W.mp = (left ? (int)t + 1 : (int)t) + 32 * (int)ta + ap;
        }
        W.s = s_T; W.M = M_T;
        if (T == 32u) { t += 1024u; continue; }
        // ---- tile T, cell by cell, by its own lane
        bool stop = false;
        WalkState V = W;
        if (lane == T) stop = wide_exact_tile(P, sub, lut16, r0, q0, left, tt, V);
        stop = __shfl_sync(0xFFFFFFFFu, (int)stop, T) != 0;
        W.s = __shfl_sync(0xFFFFFFFFu, V.s, T);
        W.M = __shfl_sync(0xFFFFFFFFu, V.M, T);
        W.mp = __shfl_sync(0xFFFFFFFFu, V.mp, T);
        if (stop) break;
        t += 32u * (T + 1u);
    }
    return DirResult{W.M, W.mp};
}

// surv[0 .. counters[surv_ctr]) -> passing HSPs appended to `anchors` (through the duplicate
// table); hits that need the entropy factor -> surv2 (count in counters[CTR_SURV2]).
// merge_min > 0: a call with more filter survivors than that is left to the merge pass
// (kernels_merge.cuh) -- the kernel returns at once and the host replays stage B on the representatives.
__global__ void __launch_bounds__(WIDE_THREADS)
k_extend_wide(ExtendParams P, const int *__restrict__ sub_mat, const SurvRec *__restrict__ surv, uint32_t surv_cap, int surv_ctr,
              uint32_t merge_min, SurvRec *__restrict__ surv2, int fused, const uint32_t *__restrict__ hit_bound,
              const uint32_t *__restrict__ plan, Anchor *__restrict__ anchors, uint32_t anchor_cap,
              uint32_t *__restrict__ counters, DedupTable dedup) {
    __shared__ uint32_t lut[WIDE_LUT_WORDS];
    __shared__ int sub[64];
    __shared__ int lut16[16];
    for (int i = threadIdx.x; i < WIDE_LUT_WORDS; i += blockDim.x) {
        const int idx = i / WIDE_LUT_COLS, rn = idx >> 4, qn = idx & 15;
        const int s0 = sub_mat[(rn & 3) * 8 + (qn & 3)], s1 = sub_mat[(rn >> 2) * 8 + (qn >> 2)];
        lut[i] = (uint32_t)(uint8_t)(int8_t)s0 | ((uint32_t)(uint8_t)(int8_t)s1 << 8);
    }
    for (int i = threadIdx.x; i < 64; i += blockDim.x) sub[i] = sub_mat[i];
    if (threadIdx.x < 16) lut16[threadIdx.x] = sub_mat[(threadIdx.x >> 2) * 8 + (threadIdx.x & 3)];
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t lut_lane = (uint32_t)__cvta_generic_to_shared(lut) + (lane & (uint32_t)(WIDE_LUT_COLS - 1)) * 4u;
    if (merge_min && counters[CTR_SURV] > merge_min) return;
    const uint32_t n = min(counters[surv_ctr], surv_cap);
    const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n; i += warps) {
        const SurvRec rec = surv[i];
        const DirResult R = wide_extend_dir(P, sub, lut16, lut_lane, rec.r0, rec.q0, false);
        const DirResult L = wide_extend_dir(P, sub, lut16, lut_lane, rec.r0, rec.q0, true);
        if (lane != 0) continue;
        const int total = R.score + L.score;
        if (total >= P.hspthresh && total <= 3 * P.hspthresh && !P.noentropy) { // :608 entropy factor needed
            surv2[atomicAdd(counters + CTR_SURV2, 1u)] = rec; // at most n records: surv2 holds surv_cap
            continue;
        }
        const int cnt[4] = {0, 0, 0, 0};
        sa_segment seg;
        uint32_t tag = 0;
        if (finish_hit(P, rec.r0, rec.q0, R, L, cnt, &seg) &&
            dedup_is_new(dedup, seg, tag = fused ? (rec.key >= counters[CTR_LASTKEY] ? 1u : 0u)
                                                  : iteration_of(hit_bound, plan[0], rec.key))) {
            const uint32_t slot = atomicAdd(counters, 1u);
            if (slot < anchor_cap) {
                Anchor a;
                a.tag = tag;
                a.ref_start = seg.ref_start;
                a.query_start = seg.query_start;
                a.len = seg.len;
                a.score = seg.score;
                anchors[slot] = a;
            }
        }
    }
}

} // namespace sa
