// kernels_extend.cuh -- exact ungapped X-drop extension, one thread per seed hit.
//
// Replaces find_hsps (src/seed_filter.cu:232-652) + the flag scan and compress_output
// (:654-680, :769-774).  Semantics: SURVEY.md Appendix A.5/A.6, restated for the CPU in
// oracle/sa_oracle.c.  Differences in HOW (not what):
//   * one thread per hit walking 32-base tiles of the 2-bit planes (0.25 B/base), instead of
//     one warp per hit gathering 1 B/base with four shuffle scans per tile;
//   * a tile whose 32 ref and query cells are all upper-case ACGT and identical is scored
//     with four popcounts (self-alignments, long exact repeats);
//   * the per-nucleotide match counters are not carried during the walk: the entropy factor
//     is only needed when hspthresh <= score <= 3*hspthresh, and then a second, tile-faithful
//     pass recounts (including the reference's count[]/count_del[] aliasing for codes >= 4);
//   * passing HSPs are appended with an atomic cursor (order is irrelevant, A.8) instead of
//     flag-scan + compaction over all hits.
#pragma once
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

// Right extension: cells k = 0,1,.. at (r0+k, q0+k); best starts at (0,-1)  (:300-453)
__device__ __forceinline__ DirResult extend_right(const ExtendParams &P, const int *sub,
                                                  const int *lut16, const int *diag, uint32_t r0,
                                                  uint32_t q0, unsigned long long *cells) {
    int s = 0, M = 0, mp = -1;
    const int X = P.xdrop;
    uint32_t t = 0;
    for (;;) {
        const uint32_t rc0 = r0 + t, qc0 = q0 + t;
        // cells at or past the end of either block are masked in the planes' padding, but a
        // window may start beyond the padding: clamp to the slow path there
        bool inside = (rc0 + 32u <= P.ref_len) && (qc0 + 32u <= P.query_len);
        uint32_t m = 0xFFFFFFFFu;
        uint64_t rw = 0, qw = 0;
        if (inside) {
            m = load_m1_window(P.rm1, rc0) | load_m1_window(P.qm1, qc0);
            if (m == 0) {
                rw = load_p2_window(P.rp2, rc0);
                qw = load_p2_window(P.qp2, qc0);
            }
        }
        if (m == 0) {
            if (rw == qw && P.diag_all_positive) {
                // strictly increasing prefix: no x-drop possible, best can only move to the
                // last cell of the tile
                s += diag_sum32(rw, diag);
                if (s > M) { M = s; mp = (int)(t + 31u); }
            } else {
                bool stop = false;
#pragma unroll
                for (int j = 0; j < 32; j++) {
                    int idx = (int)(((rw >> (2 * j)) & 3u) << 2 | ((qw >> (2 * j)) & 3u));
                    s += lut16[idx];
                    if (s > M) { M = s; mp = (int)(t + j); }
                    if (M - s > X) { stop = true; break; }
                }
                if (stop) break;
            }
            t += 32;
            if (t >= 64) *cells += 32;
            continue;
        }
        // slow tile: bounds + 1 B/base codes (:332-336); stop on x-drop or when the tile's
        // last cell is out of bounds (:420)
        bool stop = false, last_in = true;
        for (int j = 0; j < 32; j++) {
            uint32_t rp = rc0 + j, qp = qc0 + j;
            bool in = rp < P.ref_len && qp < P.query_len;
            if (in) s += sub[__ldg(P.rb8 + rp) * 8 + __ldg(P.qb8 + qp)];
            if (s > M) { M = s; mp = (int)(t + j); }
            if (M - s > X) { stop = true; break; }
            if (j == 31) last_in = in;
        }
        if (stop || !last_in) break;
        t += 32;
        if (t >= 64) *cells += 32;
    }
    return DirResult{M, mp};
}

// Left extension: cells k = 1,2,.. at (r0-k, q0-k); best starts at (0,0)  (:457-604)
__device__ __forceinline__ DirResult extend_left(const ExtendParams &P, const int *sub,
                                                 const int *lut16, const int *diag, uint32_t r0,
                                                 uint32_t q0, unsigned long long *cells) {
    int s = 0, M = 0, mp = 0;
    const int X = P.xdrop;
    uint32_t t = 0;
    for (;;) {
        // tile covers k = t+1 .. t+32, i.e. positions r0-t-32 .. r0-t-1
        bool inside = (r0 >= t + 32u) && (q0 >= t + 32u) && (r0 - t <= P.ref_len) &&
                      (q0 - t <= P.query_len);
        uint32_t m = 0xFFFFFFFFu;
        uint64_t rw = 0, qw = 0;
        if (inside) {
            const uint32_t rc0 = r0 - t - 32u, qc0 = q0 - t - 32u;
            m = load_m1_window(P.rm1, rc0) | load_m1_window(P.qm1, qc0);
            if (m == 0) {
                rw = load_p2_window(P.rp2, rc0);
                qw = load_p2_window(P.qp2, qc0);
            }
        }
        if (m == 0) {
            if (rw == qw && P.diag_all_positive) {
                s += diag_sum32(rw, diag);
                if (s > M) { M = s; mp = (int)(t + 32u); }
            } else {
                bool stop = false;
#pragma unroll
                for (int j = 0; j < 32; j++) {
                    // cell k = t+1+j sits at window cell 31-j
                    int sh = 2 * (31 - j);
                    int idx = (int)(((rw >> sh) & 3u) << 2 | ((qw >> sh) & 3u));
                    s += lut16[idx];
                    if (s > M) { M = s; mp = (int)(t + 1u + j); }
                    if (M - s > X) { stop = true; break; }
                }
                if (stop) break;
            }
            t += 32;
            if (t >= 64) *cells += 32;
            continue;
        }
        bool stop = false, last_in = true;
        for (int j = 0; j < 32; j++) {
            uint32_t k = t + 1u + j;
            bool in = r0 >= k && q0 >= k; // :482
            if (in) s += sub[__ldg(P.rb8 + (r0 - k)) * 8 + __ldg(P.qb8 + (q0 - k))];
            if (s > M) { M = s; mp = (int)k; }
            if (M - s > X) { stop = true; break; }
            if (j == 31) last_in = in;
        }
        if (stop || !last_in) break;
        t += 32;
        if (t >= 64) *cells += 32;
    }
    return DirResult{M, mp};
}

// Tile-faithful recount of the entropy counters for one direction (:436-451 / :587-602).
// frame[0..3] = count, frame[4..7] = count_del; count[c] for c>=4 aliases count_del[c-4] and
// count_del[c] for c>=4 lies outside the reference's 16-byte frame.
__device__ __noinline__ void recount_direction(const ExtendParams &P, const int *sub, uint32_t r0,
                                               uint32_t q0, bool left, int *frame) {
    int prev_score = 0, prev_max = 0, prev_pos = left ? 0 : -1;
    uint32_t tile = 0;
    frame[4] = frame[5] = frame[6] = frame[7] = 0;
    const int X = P.xdrop;
    for (;;) {
        int s = prev_score, M = prev_max, mp = prev_pos;
        bool xd = false, last_in = true;
        for (int lane = 0; lane < 32; lane++) {
            uint32_t off = left ? tile + 1u + lane : tile + lane;
            bool in;
            int v = 0;
            if (!left) {
                uint32_t rp = r0 + off, qp = q0 + off;
                in = rp < P.ref_len && qp < P.query_len;
                if (in) v = sub[__ldg(P.rb8 + rp) * 8 + __ldg(P.qb8 + qp)];
            } else {
                in = r0 >= off && q0 >= off;
                if (in) v = sub[__ldg(P.rb8 + (r0 - off)) * 8 + __ldg(P.qb8 + (q0 - off))];
            }
            if (lane == 31) last_in = in;
            if (!xd) {
                s += v;
                if (s > M) { M = s; mp = (int)off; }
                if (M - s > X) xd = true;
            }
        }
        bool new_max = mp > prev_pos;
        bool stop = xd || !last_in;
        if (!stop) { prev_score = s; prev_max = M; tile += 32; }
        prev_pos = mp;
        if (new_max) {
#pragma unroll
            for (int i = 0; i < 4; i++) { frame[i] += frame[4 + i]; frame[4 + i] = 0; }
        }
        uint32_t base = stop ? tile : tile - 32u;
        for (int lane = 0; lane < 32; lane++) {
            uint32_t off = left ? base + 1u + lane : base + lane;
            uint8_t rc, qc;
            if (!left) {
                uint32_t rp = r0 + off, qp = q0 + off;
                if (!(rp < P.ref_len && qp < P.query_len)) continue;
                rc = __ldg(P.rb8 + rp); qc = __ldg(P.qb8 + qp);
            } else {
                if (!(r0 >= off && q0 >= off)) continue;
                rc = __ldg(P.rb8 + (r0 - off)); qc = __ldg(P.qb8 + (q0 - off));
            }
            if (rc == qc) {
                int idx = ((int)off <= prev_pos) ? rc : 4 + rc;
                if (idx < 8) frame[idx] += 1;
            }
        }
        if (stop) return;
    }
}

// One hit -> HSP or nothing.  Returns true if the hit passes (d_done = 1 in the reference).
__device__ __forceinline__ bool extend_hit(const ExtendParams &P, const int *sub, const int *lut16,
                                           const int *diag, uint32_t r0, uint32_t q0,
                                           sa_segment *out, unsigned long long *cells) {
    DirResult R = extend_right(P, sub, lut16, diag, r0, q0, cells);
    DirResult L = extend_left(P, sub, lut16, diag, r0, q0, cells);
    const int total = R.score + L.score;
    const uint32_t left_extent = (uint32_t)L.pos;
    const int extent = R.pos + (int)left_extent;
    double entropy = 1.0;
    if (total >= P.hspthresh && total <= 3 * P.hspthresh && !P.noentropy) { // :608
        int frame[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        recount_direction(P, sub, r0, q0, false, frame);
        recount_direction(P, sub, r0, q0, true, frame);
        // the reference keeps the counters in `short` and sums them across lanes in `short`
        int c[4];
#pragma unroll
        for (int i = 0; i < 4; i++) c[i] = (int)(short)frame[i];
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
    uint32_t lo = 0, hi = num_iter;
    while (lo < hi) {
        uint32_t mid = (lo + hi) >> 1;
        if (hit_bound[mid] <= h) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// counters[0] = anchor cursor, counters[2..3] = ext_cells (64-bit), counters[4] = survivor count.
// With surv != nullptr the kernel walks the survivor list of the filter kernel (hit indices,
// count read from counters[4] on the device so the host need not synchronise in between);
// otherwise it walks hits [h_begin, h_end) directly.
__global__ void __launch_bounds__(128)
k_extend_hits(ExtendParams P, const int *__restrict__ sub_mat, const uint2 *__restrict__ hits,
              uint32_t h_begin, uint32_t h_end, const uint32_t *__restrict__ surv,
              const uint32_t *__restrict__ hit_bound,
              const uint32_t *__restrict__ plan, Anchor *__restrict__ anchors,
              uint32_t anchor_cap, uint32_t *__restrict__ counters) {
    __shared__ int sub[64];
    __shared__ int lut16[16];
    __shared__ int diag[4];
    if (threadIdx.x < 64) sub[threadIdx.x] = sub_mat[threadIdx.x];
    if (threadIdx.x < 16) lut16[threadIdx.x] = sub_mat[(threadIdx.x >> 2) * 8 + (threadIdx.x & 3)];
    if (threadIdx.x < 4) diag[threadIdx.x] = sub_mat[threadIdx.x * 9];
    __syncthreads();
    const uint32_t stride = gridDim.x * blockDim.x;
    unsigned long long cells = 0;
    if (surv) { h_begin = 0; h_end = counters[4]; }
    for (uint32_t i = h_begin + blockIdx.x * blockDim.x + threadIdx.x; i < h_end; i += stride) {
        const uint32_t h = surv ? surv[i] : i;
        uint2 hit = hits[h];
        sa_segment seg;
        if (extend_hit(P, sub, lut16, diag, hit.x, hit.y, &seg, &cells)) {
            uint32_t slot = atomicAdd(counters, 1u);
            if (slot < anchor_cap) {
                Anchor a;
                a.tag = iteration_of(hit_bound, plan[0], h);
                a.ref_start = seg.ref_start;
                a.query_start = seg.query_start;
                a.len = seg.len;
                a.score = seg.score;
                anchors[slot] = a;
            }
        }
    }
    if (cells && !surv) atomicAdd(reinterpret_cast<unsigned long long *>(counters + 2), cells);
}

} // namespace sa
