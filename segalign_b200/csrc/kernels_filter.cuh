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
//     "terminator" cell: a non-ACGT code whose matrix entries against every code that is not
//     "soft" are < -xdrop (screen_bound.h: screen_terminator_codes), where the reference's rule
//     always fires unless the opposite cell is soft, or a cell past the end of a block (those
//     score 0 and can never raise the maximum, :332-336,:420);
//   * M is raised by every cell of every visited group, including cells the reference would no
//     longer visit (over-estimate only);
//   * a non-ACGT cell that is not a terminator (code X under the default matrix, N/X under
//     --ambiguous) inside the walked range, the stopping cell included, makes the hit a survivor
//     outright.
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
    const uint64_t *rp2;     // reference 2-bit plane (front/back padded like the records): the screen's window
    const uint32_t *rsoft;   // reference soft-record map, bit (w + REC_FRONT); read only if ref_has_soft
    int ref_has_soft;
    int xdrop;
    int hspthresh;
    int diag_all_positive;
    // loop constants handed over as kernel parameters so that they are read from the constant
    // bank as instruction operands (as immediates ptxas re-materialises them in every group)
    uint32_t k_mul; // 4 | 64 << 8 : dp2a multipliers (16-bit fields -> LUT byte offsets)
    uint32_t k_m4;  // 0x01010101   : dp4a selector of the full group
};

// counters layout shared with the host (uint32 words)
enum { CTR_ANCHORS = 0, CTR_DEDUPE = 1, CTR_EXT_LO = 2, CTR_EXT_HI = 3, CTR_SURV = 4, CTR_CHUNK = 5, CTR_OUT = 6,
       CTR_NHITS = 7,    // fused sources: total seed hits of the call
       CTR_LASTKEY = 8,  // fused sources: seed order index of the last seed word with a non-empty bucket
       CTR_NSEEDS = 9,   // fused sources: number of seed words
       CTR_WALKED = 10,  // k_filter_hits3: hits the popcount screen left undecided (tile-walked)
       CTR_SURV2 = 11,   // k_extend_wide: survivors handed on to k_extend_hits (entropy factor needed)
       CTR_MERGED = 12,  // k_merge_mark: survivors left after dropping provable copies (kernels_merge.cuh)
       CTR_NHITS64 = 14, // (two words, 8-byte aligned) 64-bit hit total of the call: the repeat-masker header needs it
       CTR_WORDS = 16 };

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

// Where the hits of a call come from.
//   SRC_HITS : the materialised hit list of k_expand_hits (general path: any number of iterations)
//   SRC_SEEDS: straight from the caller's seed words  -- lookup, expansion and filter fused
//   SRC_RANGE: straight from the resident query block -- seeding, lookup, expansion and filter fused
// The fused sources never write the seed words' hit counts, their prefix sums or the hit list to
// HBM; they are valid when the whole call is one reference "iteration pair" (num_hits < MAX_HITS,
// SURVEY A.7), which the host checks from the hit total this kernel reports.
enum { SRC_HITS = 0, SRC_SEEDS = 1, SRC_RANGE = 2 };

struct HitSource {
    // SRC_HITS
    const uint2 *hits;
    const uint32_t *plan;      // plan[1] = num_hits
    uint32_t hits_cap;
    // SRC_SEEDS / SRC_RANGE
    const uint64_t *seeds;     // seed words (kmer << 32) + query position
    uint32_t num_items;        // seed words, or (positions x words per position) for SRC_RANGE
    const uint32_t *index_table;
    const uint32_t *pos_table;
    uint32_t seed_size;
    // repeat-masker variant: only hits whose reference anchor lies in [win_lo, win_hi] are extended; the others
    // are enumerated and counted like any hit (repeat_masker_src/seed_filter.cu:239-244).  0 .. 0xFFFFFFFF otherwise.
    uint32_t win_lo, win_hi;
    uint32_t index_size;       // 4^weight: seed words with a larger k-mer field are treated as empty buckets
    uint32_t query_len;        // seed words whose span runs past the query block likewise (caller-supplied vectors)
    // SRC_RANGE: seed words of src/seeder.cpp:57-74 generated on the fly
    uint32_t j0;               // first query position of the range
    uint32_t per;              // words per valid position: 1 + transition variants
    ShapeDesc shape;
};

// survivor of the filter: anchor pair + key (hit index for SRC_HITS, seed order index otherwise)
struct SurvRec {
    uint32_t r0, q0, key;
};

// ---------------------------------------------------------------------------------------------
// The tile-walk filter kernel (general path, and the fallback for matrices the popcount screen of
// kernels_screen.cuh does not admit).
//
// Profiling a first, purely persistent-lane version (one lane = one hit, one 32-cell tile per loop
// trip; retired) showed ~240 of its ~420 instructions per trip outside the eight 4-cell groups:
// per-trip window loads and assembly, the per-lane walk state machine and its share of work staging
// -- paid 2.9 times per hit -- plus one exposed L2 latency per trip.  Almost every hit needs
// exactly: right tile 0, left tile 0, and usually left tile 1.
//   phase 1  a warp takes 32 fresh hits, one per lane, loads the 4+4 records that cover
//            [r0-64, r0+32) / [q0-64, q0+32) with eight independent 16-byte loads, and walks
//            right 0, left 0, left 1 as straight-line code: no state machine, no per-tile loads,
//            all lanes in the same tile.  ~70 % of the hits are decided here.
//   phase 2  hits whose walk is still open (right beyond 32 cells, left beyond 64) go to a per-warp
//            continuation queue in shared memory with their walk state; when the queue holds
//            CQ_DRAIN entries (or no fresh hits are left) the warp drains it with a persistent-
//            lane loop: each lane owns one open walk and advances it by one tile per trip; a lane
//            whose walk is finished takes the next queued one.
constexpr int CQ_CAP = 64;   // continuation queue entries per warp (static + dynamic shared memory <= 48 KB)
constexpr int CQ_DRAIN = 32; // drain when at least this many are queued

struct Cont {
    uint32_t r0, q0, key, t;   // t = cells already walked in the open direction
    int s, M, right_score;
    uint32_t left;
};

// 32 cells from two adjacent records at cell offset sh (0..31) of the first
__device__ __forceinline__ void assemble_window(const uint4 a, const uint4 b, uint32_t sh, uint64_t &R,
                                                uint32_t &T, uint32_t &S) {
    const bool lo = sh < 16u;
    const uint32_t w0 = lo ? a.x : a.y, w1 = lo ? a.y : b.x, w2 = lo ? b.x : b.y;
    const uint32_t k = (2u * sh) & 31u;
    R = ((uint64_t)__funnelshift_r(w1, w2, k) << 32) | __funnelshift_r(w0, w1, k);
    T = __funnelshift_r(a.z, b.z, sh);
    S = __funnelshift_r(a.w, b.w, sh);
}

// Walk of one 32-cell window (cells in ascending address order in R/Q/T/S; `left` walks it from the
// top down).  Updates the running sum s and running max M; done = the walk of this direction ends
// in this tile; survive = a soft cell lies in the walked range (the exact kernel must decide).
__device__ __forceinline__ void tile_walk(uint32_t lut_lane, uint32_t mul, uint32_t m4, const int *diag,
                                          const FilterParams &P, uint64_t R, uint64_t Q, uint32_t T, uint32_t S,
                                          bool left, int &s, int &M, bool &done, bool &survive) {
    uint32_t rl = (uint32_t)R, rh = (uint32_t)(R >> 32), ql = (uint32_t)Q, qh = (uint32_t)(Q >> 32);
    uint32_t m1 = 0x00000001u, m2 = 0x00000101u, m3 = 0x00010101u; // dp4a prefix selectors, processing order
    if (left) {
        // descending cells: reverse the BYTES (groups); inside a group the selectors take the order
        const uint32_t a = __byte_perm(rh, 0, 0x0123), b = __byte_perm(rl, 0, 0x0123);
        const uint32_t c = __byte_perm(qh, 0, 0x0123), d = __byte_perm(ql, 0, 0x0123);
        rl = a; rh = b; ql = c; qh = d;
        T = __brev(T); S = __brev(S);
        m1 = 0x01000000u; m2 = 0x01010000u; m3 = 0x01010100u;
    }
    const int n_eff = __clz(__brev(T));          // cells before the first terminator (32 if none)
    // a soft cell in front of the first terminator cell, or AT it: a terminator code only stops the walk against a
    // cell that is not soft (screen_terminator_codes: lower case x N scores 0 under --ambiguous)
    const uint32_t valid = n_eff >= 31 ? 0xFFFFFFFFu : ((2u << n_eff) - 1u);
    survive = (S & valid) != 0;
    const int ng = survive ? 0 : (n_eff + 3) >> 2; // groups to visit (the last may run past the terminator)
    bool dropped = false;
    const int X = P.xdrop;
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
            dropped |= mn < M - X;
            M = __vimax3_s32(__vimax3_s32(p1, p2, p3), p4, M);
            s = p4;
        }
    }
    done = dropped || n_eff < 32;
}

template <int SRC>
__global__ void __launch_bounds__(FILTER_THREADS, 4)
k_filter_hits2(FilterParams P, HitSource H, const int *__restrict__ sub_mat, SurvRec *__restrict__ surv,
               uint32_t surv_cap, uint32_t *__restrict__ counters) {
    extern __shared__ uint32_t lut[];
    __shared__ int diag[4];
    for (int i = threadIdx.x; i < FILTER_LUT_WORDS; i += blockDim.x) {
        const int idx = i >> 4, rn = idx >> 4, qn = idx & 15;
        const int s0 = sub_mat[(rn & 3) * 8 + (qn & 3)], s1 = sub_mat[(rn >> 2) * 8 + (qn >> 2)];
        lut[i] = (uint32_t)(uint8_t)(int8_t)s0 | ((uint32_t)(uint8_t)(int8_t)s1 << 8);
    }
    if (threadIdx.x < 4) diag[threadIdx.x] = sub_mat[threadIdx.x * 9];
    __syncthreads();

    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t lt_mask = (1u << lane) - 1u;
    const uint32_t lut_lane = (uint32_t)__cvta_generic_to_shared(lut) + (lane & 15u) * 4u;
    const uint32_t mul = P.k_mul, m4 = P.k_m4;
    const int thr = P.hspthresh;

    __shared__ uint2 hitbuf[FILTER_THREADS / 32][FILTER_CHUNK];
    __shared__ uint8_t ownbuf[FILTER_THREADS / 32][FILTER_CHUNK];
    __shared__ uint4 contq[FILTER_THREADS / 32][CQ_CAP][2];
    uint2 *mybuf = hitbuf[threadIdx.x >> 5];
    uint8_t *myown = ownbuf[threadIdx.x >> 5];
    uint4(*myq)[2] = contq[threadIdx.x >> 5];
    const uint32_t total_items = SRC == SRC_HITS ? min(H.plan[1], H.hits_cap) : H.num_items;
    // warp-uniform staging state: key_base = hit index of staged slot 0 / seed index of the group's lane 0;
    // staged slots [cursor, limit) are unassigned; (fused) hits of the current seed group / already staged
    uint32_t key_base = 0, cursor = 0, limit = 0, g_total = 0, g_done = 0;
    bool exhausted = false;
    uint32_t acc_hits = 0, acc_seeds = 0, acc_last = 0;
    bool any_hits = false;
    uint32_t qcount = 0; // warp-uniform: queued continuations
    uint32_t ext_tiles = 0;

    auto emit = [&](uint32_t r0, uint32_t q0, uint32_t key) {
        const uint32_t slot = atomicAdd(counters + CTR_SURV, 1u);
        if (slot < surv_cap) { SurvRec rec; rec.r0 = r0; rec.q0 = q0; rec.key = key; surv[slot] = rec; }
    };

    for (;;) {
        // ---------------- stage the next chunk of fresh hits: SRC_HITS = a slice of the hit list (one
        // coalesced load); fused sources = the warp grabs 32 consecutive seed words from a global counter,
        // looks their buckets up and expands them, FILTER_CHUNK hits at a time
        while (cursor == limit && !exhausted) {
            if (SRC == SRC_HITS) {
                uint32_t c = 0;
                if (lane == 0) c = atomicAdd(counters + CTR_CHUNK, 1u);
                c = __shfl_sync(0xFFFFFFFFu, c, 0);
                const unsigned long long start = (unsigned long long)c * FILTER_CHUNK;
                key_base = start < total_items ? (uint32_t)start : total_items;
                const uint32_t cnt = min(total_items - key_base, FILTER_CHUNK);
                exhausted = cnt == 0;
                __syncwarp();
#pragma unroll
                for (uint32_t k = 0; k < FILTER_CHUNK / 32; k++)
                    if (k * 32u + lane < cnt) mybuf[k * 32u + lane] = __ldg(H.hits + key_base + k * 32u + lane);
                __syncwarp();
                cursor = 0; limit = cnt;
            } else {
                if (g_done == g_total) { // next group of 32 seed words
                    uint32_t c = 0;
                    if (lane == 0) c = atomicAdd(counters + CTR_CHUNK, 1u);
                    c = __shfl_sync(0xFFFFFFFFu, c, 0);
                    const unsigned long long start = (unsigned long long)c * 32u;
                    if (start >= total_items) { exhausted = true; break; }
                    key_base = (uint32_t)start;
                    g_done = 0;
                    g_total = 0xFFFFFFFFu;
                }
                const uint32_t k = key_base + lane;
                uint32_t b_start = 0, n = 0, qa = 0;
                bool valid = false;
                if (k < total_items) {
                    uint32_t kmer = 0, qpos = 0;
                    bool in_range = true;
                    if (SRC == SRC_SEEDS) {
                        const uint64_t word = __ldg(H.seeds + k);
                        kmer = (uint32_t)(word >> 32); qpos = (uint32_t)word;
                        valid = true;
                        in_range = kmer < H.index_size && (unsigned long long)qpos + H.seed_size <= H.query_len;
                    } else {
                        const uint32_t pi = k / H.per, v = k - pi * H.per;
                        qpos = H.j0 + pi;
                        uint64_t W; uint32_t Tw, Sw;
                        load_window(P.qrec, (int)qpos, W, Tw, Sw);
                        const uint32_t span_mask = H.shape.span >= 32 ? 0xFFFFFFFFu : ((1u << H.shape.span) - 1u);
                        valid = ((Tw | Sw) & span_mask) == 0;
                        for (int i = 0; i < H.shape.weight; i++) kmer = (kmer << 2) | (uint32_t)((W >> (2 * H.shape.pos[i])) & 3u);
                        if (v > 0) kmer ^= 2u << (2 * H.shape.tvar[v - 1]);
                    }
                    if (valid && in_range) {
                        const uint32_t b_end = __ldg(H.index_table + kmer);
                        b_start = kmer > 0 ? __ldg(H.index_table + kmer - 1) : 0u;
                        n = b_end - b_start;
                        qa = qpos + H.seed_size;
                    }
                }
                uint32_t incl = n;
#pragma unroll
                for (int off = 1; off < 32; off <<= 1) {
                    const uint32_t up = __shfl_up_sync(0xFFFFFFFFu, incl, off);
                    if (lane >= (uint32_t)off) incl += up;
                }
                const uint32_t excl = incl - n;
                if (g_total == 0xFFFFFFFFu) {
                    g_total = __shfl_sync(0xFFFFFFFFu, incl, 31);
                    acc_hits += g_total;
                    acc_seeds += __popc(__ballot_sync(0xFFFFFFFFu, valid));
                    const unsigned with_hits = __ballot_sync(0xFFFFFFFFu, n > 0);
                    if (with_hits) { acc_last = key_base + (31u - __clz(with_hits)); any_hits = true; }
                    if (g_total == 0) continue;
                }
                const uint32_t cnt = min(g_total - g_done, FILTER_CHUNK);
                __syncwarp();
#pragma unroll
                for (uint32_t kk = 0; kk < FILTER_CHUNK / 32; kk++) {
                    const uint32_t f = g_done + kk * 32u + lane;
                    uint32_t lo = 0, hi = 31;
#pragma unroll
                    for (int it = 0; it < 5; it++) {
                        const uint32_t mid = (lo + hi) >> 1;
                        const uint32_t vmid = __shfl_sync(0xFFFFFFFFu, incl, mid);
                        if (vmid > f) hi = mid; else lo = mid + 1;
                    }
                    const uint32_t o_excl = __shfl_sync(0xFFFFFFFFu, excl, lo);
                    const uint32_t o_start = __shfl_sync(0xFFFFFFFFu, b_start, lo);
                    const uint32_t o_q = __shfl_sync(0xFFFFFFFFu, qa, lo);
                    if (kk * 32u + lane < cnt) {
                        const uint32_t r = __ldg(H.pos_table + o_start + (f - o_excl)) + H.seed_size;
                        mybuf[kk * 32u + lane] = make_uint2(r, o_q);
                        myown[kk * 32u + lane] = (uint8_t)lo;
                    }
                }
                __syncwarp();
                g_done += cnt;
                cursor = 0; limit = cnt;
            }
        }
        const bool fresh = cursor < limit;
        if (!fresh && qcount == 0) break; // exhausted and nothing queued

        // ---------------- phase 1: 32 fresh hits, straight-line right 0 / left 0 / left 1
        if (fresh) {
            const uint32_t n1 = min(limit - cursor, 32u);
            bool have = lane < n1;
            uint32_t r0 = 0, q0 = 0, key = 0;
            if (have) {
                const uint32_t slot = cursor + lane;
                const uint2 hit = mybuf[slot];
                r0 = hit.x; q0 = hit.y;
                key = key_base + (SRC == SRC_HITS ? slot : (uint32_t)myown[slot]);
                have = r0 - H.win_lo <= H.win_hi - H.win_lo; // outside the caller's reference window: counted, not extended
            }
            cursor += n1;
            // records w0-2 .. w0+1 cover [r0-64, r0+32); the three windows share the cell offset r0 & 31
            const int wr = (int)(r0 >> 5), wq = (int)(q0 >> 5);
            const uint32_t shr = r0 & 31u, shq = q0 & 31u;
            const uint4 rr0 = __ldg(P.rrec + wr - 2), rr1 = __ldg(P.rrec + wr - 1), rr2 = __ldg(P.rrec + wr), rr3 = __ldg(P.rrec + wr + 1);
            const uint4 qq0 = __ldg(P.qrec + wq - 2), qq1 = __ldg(P.qrec + wq - 1), qq2 = __ldg(P.qrec + wq), qq3 = __ldg(P.qrec + wq + 1);
            uint64_t R, Q;
            uint32_t Tr, Tq, Sr, Sq;
            int s = 0, M = 0, right_score = 0;
            bool done = false, survive = false;
            bool open = have;        // this lane's hit is still undecided inside phase 1
            bool push = false;       // ... and goes to the continuation queue
            uint32_t c_t = 0, c_left = 0;
            // right tile 0: cells r0 .. r0+31
            assemble_window(rr2, rr3, shr, R, Tr, Sr);
            assemble_window(qq2, qq3, shq, Q, Tq, Sq);
            tile_walk(lut_lane, mul, m4, diag, P, R, Q, Tr | Tq, Sr | Sq, false, s, M, done, survive);
            if (open) {
                if (survive || M >= thr) { emit(r0, q0, key); open = false; }
                else if (!done) { push = true; c_t = 32; c_left = 0; open = false; }
                else right_score = M;
            }
            int cs = s, cM = M; // state of a pushed right walk
            // left tile 0: cells r0-32 .. r0-1, walked downwards
            if (__any_sync(0xFFFFFFFFu, open)) {
                s = 0; M = 0;
                assemble_window(rr1, rr2, shr, R, Tr, Sr);
                assemble_window(qq1, qq2, shq, Q, Tq, Sq);
                tile_walk(lut_lane, mul, m4, diag, P, R, Q, Tr | Tq, Sr | Sq, true, s, M, done, survive);
                if (open) {
                    if (survive || right_score + M >= thr) { emit(r0, q0, key); open = false; }
                    else if (done) open = false; // decided: not an HSP
                }
                // left tile 1: cells r0-64 .. r0-33
                if (__any_sync(0xFFFFFFFFu, open)) {
                    assemble_window(rr0, rr1, shr, R, Tr, Sr);
                    assemble_window(qq0, qq1, shq, Q, Tq, Sq);
                    tile_walk(lut_lane, mul, m4, diag, P, R, Q, Tr | Tq, Sr | Sq, true, s, M, done, survive);
                    if (open) {
                        ext_tiles++;
                        if (survive || right_score + M >= thr) emit(r0, q0, key);
                        else if (!done) { push = true; c_t = 64; c_left = 1; cs = s; cM = M; }
                        open = false;
                    }
                }
            }
            // queue the open walks (ballot compaction; CQ_CAP - CQ_DRAIN >= 32 guarantees room)
            const unsigned pm = __ballot_sync(0xFFFFFFFFu, push);
            if (push) {
                const uint32_t idx = qcount + __popc(pm & lt_mask);
                myq[idx][0] = make_uint4(r0, q0, key, c_t);
                myq[idx][1] = make_uint4((uint32_t)cs, (uint32_t)cM, (uint32_t)right_score, c_left);
            }
            qcount += __popc(pm);
            __syncwarp();
        }

        // ---------------- phase 2: drain the continuation queue with persistent lanes
        if (qcount >= (uint32_t)CQ_DRAIN || (qcount > 0 && cursor == limit && exhausted)) {
            uint32_t qhead = 0;
            bool active = false, left = false;
            uint32_t key = 0, r0 = 0, q0 = 0, t = 0;
            int s = 0, M = 0, right_score = 0;
            for (;;) {
                const unsigned need = __ballot_sync(0xFFFFFFFFu, !active);
                if (need && qhead < qcount) {
                    const uint32_t avail = qcount - qhead;
                    const uint32_t rank = __popc(need & lt_mask);
                    if (!active && rank < avail) {
                        const uint4 a = myq[qhead + rank][0], b = myq[qhead + rank][1];
                        r0 = a.x; q0 = a.y; key = a.z; t = a.w;
                        s = (int)b.x; M = (int)b.y; right_score = (int)b.z; left = b.w != 0;
                        active = true;
                    }
                    const uint32_t nneed = __popc(need);
                    qhead += nneed < avail ? nneed : avail;
                }
                if (!__any_sync(0xFFFFFFFFu, active)) break;
                if (active) {
                    const int cr = left ? (int)r0 - (int)t - 32 : (int)r0 + (int)t;
                    const int cq = left ? (int)q0 - (int)t - 32 : (int)q0 + (int)t;
                    uint64_t R, Q;
                    uint32_t Tr, Tq, Sr, Sq;
                    load_window(P.rrec, cr, R, Tr, Sr);
                    load_window(P.qrec, cq, Q, Tq, Sq);
                    bool done, survive;
                    tile_walk(lut_lane, mul, m4, diag, P, R, Q, Tr | Tq, Sr | Sq, left, s, M, done, survive);
                    ext_tiles += t >= 32u ? 1u : 0u;
                    if (survive || (left ? right_score : 0) + M >= thr) {
                        emit(r0, q0, key);
                        active = false;
                    } else if (done) {
                        if (!left) { right_score = M; left = true; t = 0; s = 0; M = 0; }
                        else active = false;
                    } else {
                        t += 32u;
                    }
                }
            }
            qcount = 0;
            __syncwarp();
        }
    }
    if (ext_tiles) atomicAdd(reinterpret_cast<unsigned long long *>(counters + CTR_EXT_LO), 32ull * ext_tiles);
    if (SRC != SRC_HITS && lane == 0) {
        if (acc_hits) {
            atomicAdd(counters + CTR_NHITS, acc_hits);
            atomicAdd(reinterpret_cast<unsigned long long *>(counters + CTR_NHITS64), (unsigned long long)acc_hits);
        }
        if (acc_seeds) atomicAdd(counters + CTR_NSEEDS, acc_seeds);
        if (any_hits) atomicMax(counters + CTR_LASTKEY, acc_last);
    }
}

// ASCII-independent record builder: b8 -> {p2 lo, p2 hi, terminator bits, soft bits} per 32 bases.
// term_codes: bit c (c = 4..7) set if code c is a guaranteed X-drop terminator under the current
// matrix.  Record index runs over [-front, words): negative and past-the-end records are pure
// terminators (block boundary).  rec points at record 0.
__global__ void __launch_bounds__(256)
k_pack_records(const uint8_t *__restrict__ b8, uint32_t len, uint4 *__restrict__ rec, int front,
               uint32_t words, uint32_t term_codes, uint32_t *__restrict__ softmap, uint32_t softmap_words) {
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
        if (soft) { // rare: IUPAC letters, or N under --ambiguous
            atomicOr(softmap + (i >> 5), 1u << (i & 31u));
            atomicAdd(softmap + softmap_words, 1u);
        }
    }
}

// "any soft cell in records w-3 .. w+2" of the record that holds cell `anchor` (six consecutive map bits)
__device__ __forceinline__ bool soft_window(const uint32_t *__restrict__ softmap, uint32_t anchor) {
    const uint32_t b = (anchor >> 5) + (uint32_t)REC_FRONT - 3u;
    const uint32_t lo = __ldg(softmap + (b >> 5)), hi = __ldg(softmap + (b >> 5) + 1);
    return (__funnelshift_r(lo, hi, b & 31u) & 0x3Fu) != 0;
}

} // namespace sa
