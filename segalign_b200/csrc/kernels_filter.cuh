// kernels_filter.cuh -- stage A of the extension: a conservative score filter over ALL seed hits.
//
// Replaces the bulk of find_hsps (src/seed_filter.cu:232-652).  Almost every seed hit is a random
// match that scores far below hspthresh; the reference nevertheless runs the full warp-per-hit
// machinery on each.  Here every hit first goes through this kernel, which computes an UPPER
// BOUND U >= right_score + left_score of the reference's X-drop extension and keeps
// the hit only if U >= hspthresh.  Since the reference emits a hit only if
// (int)(score * entropy) >= hspthresh with entropy <= 1 (:608-633), a dropped hit can never be
// an HSP; survivors are re-extended by the exact kernel (kernels_extend.cuh), which alone decides
// what is emitted.  Output parity therefore rests on two properties, both checked by tests:
// the bound never under-estimates, and stage B is exact.
//
// Why the bound holds (per direction, cells in processing order, running sum s, running max M):
//   * cells are scored with the true sub_mat values of the ACGT x ACGT block (int8 LUT);
//   * the walk stops only (a) at a cell that is > xdrop below a maximum seen BEFORE its 4-cell
//     group started -- the reference's own rule would have stopped there or earlier -- or (b) at a
//     "terminator" cell: a non-ACGT code whose every matrix entry is < -xdrop, where the
//     reference's rule always fires, or a cell past the end of a block (those score 0 and can
//     never raise the maximum, :332-336,:420);
//   * M is raised by every cell of every visited group, including cells the reference would no
//     longer visit (over-estimate only);
//   * a non-ACGT cell that is not a terminator (code X under the default matrix, N/X under
//     --ambiguous) inside the walked range makes the hit a survivor outright.
//
// Execution model: "persistent lanes".  Each lane owns one hit at a time and advances it by one
// 32-cell tile per loop trip; a lane whose hit is finished takes the next hit from a warp-level
// cursor.  All lanes therefore execute the same tile body every trip, whatever the length of
// their extensions (the reference's one-warp-per-hit and a naive one-thread-per-hit loop both
// idle most lanes: ncu showed 10.6 of 32 threads active per instruction for the latter).
// Per 4-cell group: two conflict-free shared-memory lookups (pair LUT replicated per lane/bank)
// give four int8 scores, four dp4a produce the prefix sums, vimin3/vimax3 the group min/max.
#pragma once
#include "sa_common.cuh"

namespace sa {

constexpr int FILTER_THREADS = 256;
constexpr int FILTER_LUT_WORDS = 256 * 16;          // 16 KB: entry idx for lane l at [idx*16 + (l & 15)]: at most 2-way bank conflicts
constexpr uint32_t FILTER_CHUNK = 128;              // hits per warp-level work grab (staged in shared memory)

struct FilterParams {
    const uint4 *rrec;   // reference records, index 0 = first 32 bases (front/back padded)
    const uint4 *qrec;   // query records (forward or reverse-complement block)
    int xdrop;
    int hspthresh;
    int diag_all_positive;
    // loop constants handed over as kernel parameters so that they are read from the constant
    // bank as instruction operands (as immediates ptxas re-materialises them in every group)
    uint32_t k_mul; // 4 | 64 << 8 : dp2a multipliers (16-bit fields -> LUT byte offsets)
    uint32_t k_m4;  // 0x01010101   : dp4a selector of the full group
};

// counters layout shared with the host (uint32 words)
enum { CTR_ANCHORS = 0, CTR_DEDUPE = 1, CTR_EXT_LO = 2, CTR_EXT_HI = 3, CTR_SURV = 4, CTR_CHUNK = 5, CTR_OUT = 6 };

// 32 cells starting at cell c (may be negative / past the end: the pads are terminators).
// R: 2-bit codes, cell i at bits 2i.  T: terminator bits.  S: soft (non-ACGT, non-terminator) bits.
__device__ __forceinline__ void load_window(const uint4 *__restrict__ rec, int c, uint64_t &R,
                                            uint32_t &T, uint32_t &S) {
    const int w = c >> 5;
    const uint32_t sh = (uint32_t)c & 31u;
    const uint4 a = __ldg(rec + w), b = __ldg(rec + w + 1);
    const bool lo = sh < 16u;
    const uint32_t w0 = lo ? a.x : a.y, w1 = lo ? a.y : b.x, w2 = lo ? b.x : b.y;
    const uint32_t k = (2u * sh) & 31u;
    R = ((uint64_t)__funnelshift_r(w1, w2, k) << 32) | __funnelshift_r(w0, w1, k);
    T = __funnelshift_r(a.z, b.z, sh);
    S = __funnelshift_r(a.w, b.w, sh);
}

// reverse the order of the 32 two-bit fields
__device__ __forceinline__ uint64_t reverse_fields(uint64_t x) {
    x = __brevll(x);
    return ((x >> 1) & 0x5555555555555555ull) | ((x & 0x5555555555555555ull) << 1);
}

__device__ __forceinline__ int diag_sum32_f(uint64_t win, const int *diag) {
    const uint64_t M5 = 0x5555555555555555ull;
    uint64_t lo = win & M5, hi = (win >> 1) & M5;
    int nT = __popcll(lo & hi), nC = __popcll(lo & ~hi), nG = __popcll(hi & ~lo);
    int nA = 32 - nT - nC - nG;
    return nA * diag[0] + nC * diag[1] + nG * diag[2] + nT * diag[3];
}

// One 4-cell group.  y = ref byte << 16 | query byte (bytes 1 and 3 are junk): the group's 8 bits
// of ref codes and 8 bits of query codes.  m1..m4 = dp4a selectors of the four prefixes in
// processing order (ascending cells for the right walk, descending for the left walk).
// Index arithmetic runs on the FMA pipe (dp2a: 16-bit fields x 8-bit multipliers) because the
// integer ALU pipe is this kernel's bottleneck.  Returns true if the walk must stop (a cell fell
// more than xdrop below the pre-group maximum).
__device__ __forceinline__ uint32_t lds_u32(uint32_t saddr) {
    uint32_t v;
    asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(saddr));
    return v;
}

// Scores of one 4-cell group as four int8.  lut_lane = shared-space byte address of this lane's
// LUT column (lut + (lane & 15)*4); mul = 4 | 64<<8; y = ref byte << 16 | query byte (bytes 1, 3 junk).
// Index arithmetic runs on the FMA pipe (dp2a: 16-bit fields x 8-bit multipliers) because the
// integer ALU pipe is this kernel's busiest.
__device__ __forceinline__ uint32_t group_scores(uint32_t lut_lane, uint32_t mul, uint32_t y) {
    const uint32_t oa = __dp2a_lo(y & 0x000F000Fu, mul, 0u) * 16u + lut_lane;   // (Rlo*16 + Qlo) * 64 + column
    const uint32_t ob = __dp2a_lo(y & 0x00F000F0u, mul, lut_lane);              // (Rhi*16 + Qhi) * 64 + column
    return __byte_perm(lds_u32(oa), lds_u32(ob), 0x5410);
}

__global__ void __launch_bounds__(FILTER_THREADS, 6)
k_filter_hits(FilterParams P, const int *__restrict__ sub_mat, const uint2 *__restrict__ hits,
              const uint32_t *__restrict__ plan, uint32_t hits_cap, uint32_t *__restrict__ surv,
              uint32_t *__restrict__ counters) {
    extern __shared__ uint32_t lut[];
    __shared__ int diag[4];
    for (int i = threadIdx.x; i < FILTER_LUT_WORDS; i += blockDim.x) {
        const int idx = i >> 4, rn = idx >> 4, qn = idx & 15;
        const int s0 = sub_mat[(rn & 3) * 8 + (qn & 3)], s1 = sub_mat[(rn >> 2) * 8 + (qn >> 2)];
        lut[i] = (uint32_t)(uint8_t)(int8_t)s0 | ((uint32_t)(uint8_t)(int8_t)s1 << 8);
    }
    if (threadIdx.x < 4) diag[threadIdx.x] = sub_mat[threadIdx.x * 9];
    __syncthreads();

    const uint32_t num_hits = min(plan[1], hits_cap);
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t lt_mask = (1u << lane) - 1u;
    // shared-space address of this lane's LUT column and the loop constants, pinned in registers
    // (otherwise ptxas re-materialises them through the uniform datapath inside every group)
    const uint32_t lut_lane = (uint32_t)__cvta_generic_to_shared(lut) + (lane & 15u) * 4u;
    const uint32_t mul = P.k_mul, m4 = P.k_m4;
    const int X = P.xdrop;

    // Work distribution: a warp grabs FILTER_CHUNK consecutive hits at a time from a global
    // counter (fine-grained, so that all warps finish together) and stages them in shared memory
    // with one coalesced load; lanes then pick hits from the staged chunk as they become free.
    __shared__ uint2 hitbuf[FILTER_THREADS / 32][FILTER_CHUNK];
    uint2 *mybuf = hitbuf[threadIdx.x >> 5];
    uint32_t chunk_start = 0, cursor = 0, limit = 0; // warp-uniform: current chunk [chunk_start, limit), next unassigned hit
    bool exhausted = false;                          // warp-uniform: the global chunk counter ran past num_hits
    // current hit of this lane
    bool active = false, left = false;
    uint32_t h = 0, r0 = 0, q0 = 0, t = 0;
    int s = 0, M = 0, right_score = 0;
    unsigned long long ext_cells = 0;

    for (;;) {
        const unsigned need = __ballot_sync(0xFFFFFFFFu, !active);
        if (need && !exhausted) {
            if (cursor == limit) { // chunk used up: grab and stage the next one
                uint32_t c = 0;
                if (lane == 0) c = atomicAdd(counters + CTR_CHUNK, 1u);
                c = __shfl_sync(0xFFFFFFFFu, c, 0);
                const unsigned long long start = (unsigned long long)c * FILTER_CHUNK;
                chunk_start = start < num_hits ? (uint32_t)start : num_hits;
                limit = (num_hits - chunk_start < FILTER_CHUNK) ? num_hits : chunk_start + FILTER_CHUNK;
                cursor = chunk_start;
                exhausted = cursor == limit;
                __syncwarp();
#pragma unroll
                for (uint32_t k = 0; k < FILTER_CHUNK / 32; k++) {
                    const uint32_t i = chunk_start + k * 32u + lane;
                    if (i < limit) mybuf[k * 32u + lane] = __ldg(hits + i);
                }
                __syncwarp();
            }
            const uint32_t avail = limit - cursor;
            const uint32_t rank = __popc(need & lt_mask);
            if (!active && rank < avail) {
                h = cursor + rank;
                const uint2 hit = mybuf[h - chunk_start];
                r0 = hit.x; q0 = hit.y;
                active = true; left = false; t = 0; s = 0; M = 0;
            }
            const uint32_t nneed = __popc(need);
            cursor += nneed < avail ? nneed : avail;
        }
        if (!__any_sync(0xFFFFFFFFu, active)) {
            if (exhausted) break;
            continue;
        }
        if (active) {
            // window of this trip: right = cells r0+t .. r0+t+31 ; left = cells r0-t-32 .. r0-t-1
            const int cr = left ? (int)r0 - (int)t - 32 : (int)r0 + (int)t;
            const int cq = left ? (int)q0 - (int)t - 32 : (int)q0 + (int)t;
            uint64_t R, Q;
            uint32_t Tr, Tq, Sr, Sq;
            load_window(P.rrec, cr, R, Tr, Sr);
            load_window(P.qrec, cq, Q, Tq, Sq);
            uint32_t T = Tr | Tq, S = Sr | Sq;
            uint32_t rl = (uint32_t)R, rh = (uint32_t)(R >> 32), ql = (uint32_t)Q, qh = (uint32_t)(Q >> 32);
            // dp4a prefix selectors in processing order
            uint32_t m1 = 0x00000001u, m2 = 0x00000101u, m3 = 0x00010101u;
            if (left) {
                // processing order = descending cells: reverse the BYTES (groups) of the window;
                // inside a group the descending order is taken by the selectors
                const uint32_t a = __byte_perm(rh, 0, 0x0123), b = __byte_perm(rl, 0, 0x0123);
                const uint32_t c = __byte_perm(qh, 0, 0x0123), d = __byte_perm(ql, 0, 0x0123);
                rl = a; rh = b; ql = c; qh = d;
                T = __brev(T); S = __brev(S);
                m1 = 0x01000000u; m2 = 0x01010000u; m3 = 0x01010100u;
            }
            const int n_eff = __clz(__brev(T));          // cells before the first terminator (32 if none)
            const uint32_t valid = n_eff >= 32 ? 0xFFFFFFFFu : ((1u << n_eff) - 1u);
            const bool survive = (S & valid) != 0;        // soft cell in range: let the exact kernel decide
            // ---- the tile body is straight-line code: first all table lookups of the tile (independent
            // of the running score, so their latencies overlap), then the short dependent chain.
            // Groups at or past the one holding the first terminator are neutralised (score 0).
            const int ng = survive ? 0 : (n_eff + 3) >> 2; // groups to visit (the last may run past the terminator)
            bool dropped = false;
            if (n_eff == 32 && !survive && rl == ql && rh == qh && P.diag_all_positive) {
                s += diag_sum32_f(R, diag);               // all-match tile: strictly increasing prefix
                M = max(M, s);
            } else {
                uint32_t sc[8];
#pragma unroll
                for (int g = 0; g < 8; g++) {
                    const uint32_t y = g < 4 ? __byte_perm(rl, ql, 0x0040 | (g & 3) << 8 | (4 + (g & 3)))
                                             : __byte_perm(rh, qh, 0x0040 | (g & 3) << 8 | (4 + (g & 3)));
                    const uint32_t v = group_scores(lut_lane, mul, y);
                    sc[g] = g < ng ? v : 0u;
                }
#pragma unroll
                for (int g = 0; g < 8; g++) {
                    const int p1 = __dp4a((int)sc[g], (int)m1, s);
                    const int p2 = __dp4a((int)sc[g], (int)m2, s);
                    const int p3 = __dp4a((int)sc[g], (int)m3, s);
                    const int p4 = __dp4a((int)sc[g], (int)m4, s);
                    const int mn = min(__vimin3_s32(p1, p2, p3), p4);
                    // a cell more than xdrop below a maximum reached BEFORE this group: the reference's
                    // walk has stopped by here.  Later groups of the tile still raise M (over-estimate).
                    dropped |= mn < M - X;
                    M = __vimax3_s32(__vimax3_s32(p1, p2, p3), p4, M);
                    s = p4;
                }
            }
            const bool done = dropped || n_eff < 32;
            if (t >= 32u) ext_cells += 32;
            if (survive) {
                surv[atomicAdd(counters + CTR_SURV, 1u)] = h;
                active = false;
            } else if (done) {
                if (!left) {
                    right_score = M;
                    left = true; t = 0; s = 0; M = 0;
                } else {
                    if (right_score + M >= P.hspthresh) surv[atomicAdd(counters + CTR_SURV, 1u)] = h;
                    active = false;
                }
            } else {
                t += 32u;
            }
        }
    }
    if (ext_cells) atomicAdd(reinterpret_cast<unsigned long long *>(counters + CTR_EXT_LO), ext_cells);
}

// ASCII-independent record builder: b8 -> {p2 lo, p2 hi, terminator bits, soft bits} per 32 bases.
// term_codes: bit c (c = 4..7) set if code c is a guaranteed X-drop terminator under the current
// matrix.  Record index runs over [-front, words): negative and past-the-end records are pure
// terminators (block boundary).  rec points at record 0.
__global__ void __launch_bounds__(256)
k_pack_records(const uint8_t *__restrict__ b8, uint32_t len, uint4 *__restrict__ rec, int front,
               uint32_t words, uint32_t term_codes) {
    const uint32_t total = words + (uint32_t)front;
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int w = (int)i - front;
        uint64_t bits = 0;
        uint32_t term = 0, soft = 0;
        for (int cell = 0; cell < 32; cell++) {
            const long long pos = (long long)w * 32 + cell;
            if (pos < 0 || pos >= (long long)len) { term |= 1u << cell; continue; }
            const uint32_t c = b8[pos];
            if (c < 4) bits |= (uint64_t)c << (2 * cell);
            else if ((term_codes >> c) & 1u) term |= 1u << cell;
            else soft |= 1u << cell;
        }
        rec[w] = make_uint4((uint32_t)bits, (uint32_t)(bits >> 32), term, soft);
    }
}

} // namespace sa
