// kernels_extend.cuh -- stage B: exact ungapped X-drop extension of the filter's survivors.
//
// Replaces find_hsps (src/seed_filter.cu:232-652) + the flag scan and compress_output
// (:654-680, :769-774).  Semantics: SURVEY.md Appendix A.5/A.6, restated for the CPU in
// oracle/sa_oracle.c.  Differences in HOW (not what):
//   * two lanes per hit -- one walks right, one walks left -- over 32-base tiles of the 2-bit
//     planes (0.25 B/base), instead of one warp per hit gathering 1 B/base with four shuffle
//     scans per tile; the pair meets through one warp shuffle;
//   * a tile whose 32 ref and query cells are all upper-case ACGT and identical is scored
//     with four popcounts (self-alignments, long exact repeats);
//   * the per-nucleotide match counters of the entropy factor are kept per tile with popcounts
//     on the 2-bit planes, following the reference's count[]/count_del[] merge rule tile by
//     tile (including its aliasing for codes >= 4, which only tiles with non-ACGT cells can
//     trigger; those take the byte path);
//   * passing HSPs are appended with an atomic cursor (order is irrelevant, A.8) instead of
//     flag-scan + compaction over all hits.
#pragma once
#include "kernels_filter.cuh"
#include "sa_common.cuh"

namespace sa {

struct DirResult {
    int score;
    int pos;
};

__device__ __forceinline__ int diag_sum32(uint64_t win, const int *diag) {
    const uint64_t M5 = 0x5555555555555555ull;
    uint64_t lo = win & M5, hi = (win >> 1) & M5;
    int nT = __popcll(lo & hi), nC = __popcll(lo & ~hi), nG = __popcll(hi & ~lo);
    int nA = 32 - nT - nC - nG;
    return nA * diag[0] + nC * diag[1] + nG * diag[2] + nT * diag[3];
}

// Entropy counters of one direction: cnt[c] = count[c], del[c] = count_del[c]
// (src/seed_filter.cu:314-321).  The reference indexes both short[4] arrays with codes up to 7:
// count[4+i] aliases count_del[i]; count_del[4+i] lies outside its 16-byte frame.
struct Counters {
    int cnt[4];
    int del[4];
};

// reverse the order of the 32 two-bit fields of a window
__device__ __forceinline__ uint64_t reverse_fields32(uint64_t x) {
    x = __brevll(x);
    return ((x >> 1) & 0x5555555555555555ull) | ((x & 0x5555555555555555ull) << 1);
}

// Counter update of a tile without non-ACGT cells (:436-451 / :587-602).  rw/qw hold the cells
// in processing order; nle = number of leading cells whose position is <= the running max
// position.
__device__ __forceinline__ void count_fast_tile(uint64_t rw, uint64_t qw, int nle, Counters &C) {
    const uint64_t M5 = 0x5555555555555555ull;
    const uint64_t x = rw ^ qw;
    const uint64_t eq = ~(x | (x >> 1)) & M5;           // bit 2i set: cell i matches
    const uint64_t le = nle <= 0 ? 0ull : (nle >= 32 ? ~0ull : ((1ull << (2 * nle)) - 1ull));
    const uint64_t lo = rw & M5, hi = (rw >> 1) & M5;
    const uint64_t m[4] = {eq & ~lo & ~hi, eq & lo & ~hi, eq & hi & ~lo, eq & lo & hi};
#pragma unroll
    for (int c = 0; c < 4; c++) {
        C.cnt[c] += __popcll(m[c] & le);
        C.del[c] += __popcll(m[c] & ~le);
    }
}

// One direction of one hit (:300-453 right, :457-604 left), exact.
//   right: cells k = 0,1,..  at (r0+k, q0+k); best starts at (0,-1)
//   left : cells k = 1,2,..  at (r0-k, q0-k); best starts at (0, 0)
// `left` is a run-time flag on purpose: the two lanes of a hit sit in one warp and must execute
// the same instructions to run concurrently.
__device__ __forceinline__ DirResult extend_dir(const ExtendParams &P, const int *sub, const int *lut16,
                                             const int *diag, uint32_t r0, uint32_t q0, bool left,
                                             Counters &C, unsigned long long *cells) {
    int s = 0, M = 0, mp = left ? 0 : -1;
    const int X = P.xdrop;
    uint32_t t = 0;
#pragma unroll
    for (int c = 0; c < 4; c++) C.del[c] = 0; // :318-321 / :471-474
    for (;;) {
        const int prev_mp = mp;
        const int base = left ? (int)t + 1 : (int)t; // position of the tile's first cell in processing order
        bool inside;
        uint32_t rc0, qc0; // window start (lowest address)
        if (!left) {
            rc0 = r0 + t; qc0 = q0 + t;
            inside = (rc0 + 32u <= P.ref_len) && (qc0 + 32u <= P.query_len);
        } else {
            inside = (r0 >= t + 32u) && (q0 >= t + 32u) && (r0 - t <= P.ref_len) && (q0 - t <= P.query_len);
            rc0 = r0 - t - 32u; qc0 = q0 - t - 32u;
        }
        uint32_t m = 0xFFFFFFFFu;
        uint64_t rw = 0, qw = 0;
        if (inside) {
            m = load_m1_window(P.rm1, rc0) | load_m1_window(P.qm1, qc0);
            if (m == 0) {
                rw = load_p2_window(P.rp2, rc0);
                qw = load_p2_window(P.qp2, qc0);
                if (left) { rw = reverse_fields32(rw); qw = reverse_fields32(qw); } // cell k = t+1+j first
            }
        }
        bool stop = false;
        if (m == 0) {
            if (rw == qw && P.diag_all_positive) {
                // strictly increasing prefix: no x-drop possible, the best moves to the last cell
                s += diag_sum32(rw, diag);
                if (s > M) { M = s; mp = base + 31; }
            } else if (P.scores_fit_int8) {
                // 4-cell groups: all table lookups first (they do not depend on the running score),
                // then per group four dp4a prefix sums; the group is walked cell by cell only when
                // an x-drop inside it cannot be excluded.  Same result as the sequential rule.
                uint32_t sc[8];
#pragma unroll
                for (int g = 0; g < 8; g++) {
                    const uint32_t rb = (uint32_t)(rw >> (8 * g)) & 0xFFu, qb = (uint32_t)(qw >> (8 * g)) & 0xFFu;
                    uint32_t v = 0;
#pragma unroll
                    for (int c = 0; c < 4; c++) {
                        const uint32_t idx = ((rb >> (2 * c)) & 3u) << 2 | ((qb >> (2 * c)) & 3u);
                        v |= ((uint32_t)lut16[idx] & 0xFFu) << (8 * c);
                    }
                    sc[g] = v;
                }
#pragma unroll
                for (int g = 0; g < 8; g++) {
                    if (stop) break;
                    const int p1 = __dp4a((int)sc[g], 0x00000001, s), p2 = __dp4a((int)sc[g], 0x00000101, s);
                    const int p3 = __dp4a((int)sc[g], 0x00010101, s), p4 = __dp4a((int)sc[g], 0x01010101, s);
                    const int mx = max(__vimax3_s32(p1, p2, p3), p4);
                    const int mn = min(__vimin3_s32(p1, p2, p3), p4);
                    if (max(M, mx) - mn <= X) { // no cell of the group can be > xdrop below the running max
                        if (mx > M) {
                            M = mx;
                            mp = base + 4 * g + (p1 == mx ? 0 : (p2 == mx ? 1 : (p3 == mx ? 2 : 3)));
                        }
                        s = p4;
                    } else {
                        const int pj[4] = {p1, p2, p3, p4};
#pragma unroll
                        for (int c = 0; c < 4; c++) {
                            if (!stop) {
                                s = pj[c];
                                if (s > M) { M = s; mp = base + 4 * g + c; }
                                if (M - s > X) stop = true;
                            }
                        }
                    }
                }
            } else {
#pragma unroll 8
                for (int j = 0; j < 32; j++) {
                    const int idx = (int)(((rw >> (2 * j)) & 3u) << 2 | ((qw >> (2 * j)) & 3u));
                    s += lut16[idx];
                    if (s > M) { M = s; mp = base + j; }
                    if (M - s > X) { stop = true; break; }
                }
            }
            if (mp > prev_mp) { // :408-411, :436-441
#pragma unroll
                for (int c = 0; c < 4; c++) { C.cnt[c] += C.del[c]; C.del[c] = 0; }
            }
            count_fast_tile(rw, qw, mp - base + 1, C);
        } else if (P.zskip && inside && zero_tile(P.rz, P.qz, rc0, qc0)) {
            // a tile of zero-scoring pairs without an ACGT match changes nothing; neither do the pieces behind it that
            // are flat on one block and partners on the other (an N run of megabases: 32 768 cells per trip); zero_runs.h
            const uint32_t k = max(32u, zero_jump(P.rz, P.qz, left ? r0 - t : rc0, left ? q0 - t : qc0, left));
            t += k;
            *cells += k;
            continue;
        } else {
            // tile with a non-ACGT cell or touching a block end: 1 B/base codes, bounds per cell
            // (:328-336 / :482); stop on x-drop or when the tile's last cell is out of bounds (:420)
            bool xd = false, last_in = true;
#pragma unroll 4
            for (int j = 0; j < 32; j++) {
                const uint32_t k = (uint32_t)(base + j);
                const bool in = left ? (r0 >= k && q0 >= k) : (r0 + k < P.ref_len && q0 + k < P.query_len);
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
            if (mp > prev_mp) {
#pragma unroll
                for (int c = 0; c < 4; c++) { C.cnt[c] += C.del[c]; C.del[c] = 0; }
            }
#pragma unroll 4
            for (int j = 0; j < 32; j++) { // :444-451 with the reference's out-of-range indexing
                const uint32_t k = (uint32_t)(base + j);
                const bool in = left ? (r0 >= k && q0 >= k) : (r0 + k < P.ref_len && q0 + k < P.query_len);
                if (!in) continue;
                const uint32_t rp = left ? r0 - k : r0 + k, qp = left ? q0 - k : q0 + k;
                const int c = __ldg(P.rb8 + rp);
                if (c != (int)__ldg(P.qb8 + qp)) continue;
                const bool le = base + j <= mp;
                if (c < 4) { if (le) C.cnt[c]++; else C.del[c]++; }
                else if (le) C.del[c - 4]++;
            }
        }
        if (stop) break;
        t += 32;
        if (t >= 64) *cells += 32;
    }
    return DirResult{M, mp};
}

// Threshold + entropy (:608-649).  c[] = count[] after both directions.
__device__ __forceinline__ bool finish_hit(const ExtendParams &P, uint32_t r0, uint32_t q0, DirResult R,
                                           DirResult L, const int *cnt, sa_segment *out) {
    const int total = R.score + L.score;
    const uint32_t left_extent = (uint32_t)L.pos;
    const int extent = R.pos + (int)left_extent;
    double entropy = 1.0;
    if (total >= P.hspthresh && total <= 3 * P.hspthresh && !P.noentropy) { // :608
        // the reference keeps the counters in `short` and sums them across lanes in `short`
        int c[4];
#pragma unroll
        for (int i = 0; i < 4; i++) c[i] = (int)(short)cnt[i];
        if (c[0] + c[1] + c[2] + c[3] >= 20) { // :617
            const double Ld = (double)(extent + 1);
            double e = 0.0;
#pragma unroll
            for (int i = 0; i < 4; i++) { // :620-622, one fma.rn.f64 per term
                double pr = __ddiv_rn((double)c[i], Ld);
                double lg = (c[i] != 0) ? log(pr) : 0.0;
                e = __fma_rn(pr, lg, e);
            }
            // :623 -entropy/log(4.0f): the float overload, constant-folded by nvcc
            entropy = __ddiv_rn(-e, (double)1.38629436111989061883f);
        }
    }
    // :633 pass test in (float)score * entropy, :638 stored score in (double)score * entropy
    if ((int)__dmul_rn((double)(float)total, entropy) >= P.hspthresh) {
        out->ref_start = r0 - left_extent;
        out->query_start = q0 - left_extent;
        out->len = (uint32_t)extent;
        out->score = entropy > 0 ? (int)__dmul_rn((double)total, entropy) : 0;
        return true;
    }
    return false;
}

__device__ __forceinline__ uint32_t iteration_of(const uint32_t *__restrict__ hit_bound,
                                                 uint32_t num_iter, uint32_t h) {
    // number of iteration ends <= h
    if (num_iter == 0xFFFFFFFFu) return 0; // the plan overflowed its arrays: the host discards this attempt and replays
    uint32_t lo = 0, hi = num_iter;
    while (lo < hi) {
        uint32_t mid = (lo + hi) >> 1;
        if (hit_bound[mid] <= h) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// Exact-duplicate suppression at append time.  Every seed hit inside one homologous run extends
// to the same HSP, so ~99 % of the passing records of a call are byte-identical copies.  Removing
// identical records (same iteration tag) before the diagonal sort cannot change the reference's
// sort -> unique_copy -> sort result: copies sort next to each other and unique_copy keeps the
// first of them, compared against the same predecessor.  Open addressing; a full
// or contended table only means "not suppressed" -- never a dropped record.
struct DedupTable {
    unsigned long long *k0; // ref_start | query_start << 32   (all-ones = empty)
    unsigned long long *k1; // len | score << 32               (all-ones = empty)
    uint32_t *tagbits;      // bit t: the record was already appended under iteration tag t
    uint32_t mask;          // slots - 1 (power of two); 0 disables the table
};

__device__ __forceinline__ bool dedup_is_new(const DedupTable &T, const sa_segment &seg, uint32_t tag) {
    if (T.mask == 0 || tag >= 32u) return true;
    const unsigned long long my0 = (unsigned long long)seg.ref_start | ((unsigned long long)seg.query_start << 32);
    const unsigned long long my1 = (unsigned long long)seg.len | ((unsigned long long)(uint32_t)seg.score << 32);
    if (my0 == ~0ull || my1 == ~0ull) return true;
    uint32_t i = (uint32_t)((my0 * 0x9E3779B97F4A7C15ull) >> 40) ^ (uint32_t)(my1 * 0x85EBCA6Bu);
    for (int probe = 0; probe < 64; probe++, i++) {
        i &= T.mask;
        const unsigned long long o0 = atomicCAS(T.k0 + i, ~0ull, my0);
        if (o0 != ~0ull && o0 != my0) continue;
        const unsigned long long o1 = atomicCAS(T.k1 + i, ~0ull, my1);
        if (o1 != ~0ull && o1 != my1) continue; // same start pair, other length/score: next slot
        return ((atomicOr(T.tagbits + i, 1u << tag) >> tag) & 1u) == 0;
    }
    return true;
}

constexpr int EXTEND_THREADS = 32; // one warp per block: fits beside the resident filter blocks

// counters: see the CTR_* enum (kernels_filter.cuh).
// With surv != nullptr the kernel walks survivor records (those of the filter kernel, or the ones
// k_extend_wide handed on; their count is read from counters[surv_ctr] on the device so the host
// need not synchronise in between);
// otherwise it walks hits [0, min(plan[1], h_end)) directly.  Lanes 2i / 2i+1 of a warp extend work
// item i to the right / to the left.
// Iteration tag of a hit (dedupe scope): general path = position of the hit index in
// the plan's hit bounds; fused path (one iteration pair, num_hits < MAX_HITS) = 0 for hits of seed
// words before the last hit-bearing seed word, 1 from that seed word on.
__global__ void __launch_bounds__(EXTEND_THREADS)
k_extend_hits(ExtendParams P, const int *__restrict__ sub_mat, const uint2 *__restrict__ hits,
              uint32_t h_end, const SurvRec *__restrict__ surv, uint32_t surv_cap, int surv_ctr, uint32_t merge_min, int fused,
              const uint32_t *__restrict__ hit_bound,
              const uint32_t *__restrict__ plan, Anchor *__restrict__ anchors,
              uint32_t anchor_cap, uint32_t *__restrict__ counters, DedupTable dedup) {
    __shared__ int sub[64];
    __shared__ int lut16[16];
    __shared__ int diag[4];
    for (int i = threadIdx.x; i < 64; i += blockDim.x) sub[i] = sub_mat[i];
    if (threadIdx.x < 16) lut16[threadIdx.x] = sub_mat[(threadIdx.x >> 2) * 8 + (threadIdx.x & 3)];
    if (threadIdx.x < 4) diag[threadIdx.x] = sub_mat[threadIdx.x * 9];
    __syncthreads();
    if (surv && merge_min && counters[CTR_SURV] > merge_min) return; // left to the merge pass (kernels_merge.cuh)
    h_end = surv ? min(counters[surv_ctr], surv_cap) : min(plan[1], h_end); // counts known only on the device
    const uint32_t items_per_pass = (gridDim.x * blockDim.x) >> 1;
    const bool left = threadIdx.x & 1u;
    unsigned long long cells = 0;
    // whole warps iterate together (the pair exchange below is a full-warp shuffle)
    const uint32_t n_pass = (h_end + items_per_pass - 1) / items_per_pass;
    for (uint32_t pass = 0; pass < n_pass; pass++) {
        const uint32_t i = pass * items_per_pass + ((blockIdx.x * blockDim.x + threadIdx.x) >> 1);
        bool valid = i < h_end;
        uint32_t h = 0;
        uint2 hit = make_uint2(0, 0);
        DirResult D = {0, 0};
        Counters C;
#pragma unroll
        for (int c = 0; c < 4; c++) { C.cnt[c] = 0; C.del[c] = 0; }
        if (valid) {
            if (surv) { const SurvRec rec = surv[i]; h = rec.key; hit = make_uint2(rec.r0, rec.q0); }
            else { h = i; hit = hits[i]; }
            valid = hit.x - P.win_lo <= P.win_hi - P.win_lo; // repeat-masker variant: outside the reference window
            if (valid) D = extend_dir(P, sub, lut16, diag, hit.x, hit.y, left, C, &cells);
        }
        // the right lane (even) receives the left lane's result
        DirResult L;
        L.score = __shfl_down_sync(0xFFFFFFFFu, D.score, 1);
        L.pos = __shfl_down_sync(0xFFFFFFFFu, D.pos, 1);
        int cnt[4];
#pragma unroll
        for (int c = 0; c < 4; c++) cnt[c] = C.cnt[c] + __shfl_down_sync(0xFFFFFFFFu, C.cnt[c], 1);
        if (valid && !left) {
            sa_segment seg;
            uint32_t tag = 0;
            if (finish_hit(P, hit.x, hit.y, D, L, cnt, &seg) &&
                dedup_is_new(dedup, seg, tag = fused ? (h >= counters[CTR_LASTKEY] ? 1u : 0u)
                                                  : iteration_of(hit_bound, plan[0], h))) {
                uint32_t slot = atomicAdd(counters, 1u);
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
    if (cells && !surv) atomicAdd(reinterpret_cast<unsigned long long *>(counters + 2), cells);
}

} // namespace sa
