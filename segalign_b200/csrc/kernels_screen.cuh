// kernels_screen.cuh -- the default filter kernel: popcount screen first, tile walk for the rest.
//
// Same contract as k_filter_hits / k_filter_hits2 (kernels_filter.cuh): consumes every seed hit of
// a SeedAndFilter call (src/seed_filter.cu:157-230 fused with the bulk of :232-652), writes the
// hits that may reach hspthresh to the survivor list for the exact kernel.  What changed is where
// the time goes.  ncu of the tile-walk kernels (profiles/r1j_*) showed the L1/shared-memory pipe
// at 70 % and ~1400 instructions per hit: every lane gathered its own 16-byte records (one L1 tag
// lookup per lane and record, ~7 per hit) and walked ~3 tiles through the pair LUT.  Here
//
//   1. query windows are aligned once per seed word (not per hit) and kept in shared memory:
//      all hits of a seed word share the query anchor;
//   2. the reference window of a hit (6 consecutive records = 96 bytes) is fetched by six
//      neighbouring lanes with cp.async straight into shared memory: a 32-lane request touches
//      ~9 lines instead of 32, the owner lane then reads its records with conflict-free LDS.128;
//   3. the hit is decided by the popcount screen of screen_bound.h (~15 instructions per 16-cell
//      block, 10 blocks) -- about 98 % of random hits end here;
//   4. undecided hits are queued per warp and walked by the persistent-lane tile loop of
//      kernels_filter.cuh (same code, same bound as before), 64 at a time.
#pragma once
#include "kernels_filter.cuh"
#include "screen_bound.h"

namespace sa {

constexpr int SCR_STAGE_STRIDE = 7;  // uint4 slots per hit in the staging buffer (6 used; 7 = conflict-free LDS.128)
constexpr int SCR_ROW_STRIDE = SCREEN_ROW_WORDS; // 48 bytes: conflict-free for 16-byte reads
constexpr int SCR_RING = 256;        // staged hits per warp (ring, power of two)
constexpr int SCR_ROWS = 64;         // aligned query rows per warp (ring, power of two)
constexpr int SCR_REFILL = 64;       // refill the ring when fewer hits than this are staged (two rounds)
constexpr int SCR_Q_CAP = 96;
constexpr int SCR_Q_DRAIN = 64;
constexpr int SCR_WARPS = FILTER_THREADS / 32;
constexpr int SCR_LUT_COLS = 4;      // pair-LUT replicas (the tile walk is ~6 % of this kernel: conflicts are cheap)
constexpr int SCR_LUT_WORDS = 256 * SCR_LUT_COLS;
constexpr uint32_t SCR_K_MUL = 1u | (16u << 8); // dp2a multipliers of group_scores for a 4-column LUT

// dynamic shared memory layout of k_filter_hits3 (bytes)
constexpr size_t SCR_OFF_LUT = 0;
constexpr size_t SCR_OFF_STAGE = SCR_OFF_LUT + SCR_LUT_WORDS * 4;
constexpr size_t SCR_OFF_ROWS = SCR_OFF_STAGE + (size_t)SCR_WARPS * 32 * SCR_STAGE_STRIDE * 16;
constexpr size_t SCR_OFF_RING = SCR_OFF_ROWS + (size_t)SCR_WARPS * SCR_ROWS * SCR_ROW_STRIDE * 4;
constexpr size_t SCR_OFF_QUEUE = SCR_OFF_RING + (size_t)SCR_WARPS * SCR_RING * 4;
constexpr size_t SCR_OFF_RROW = SCR_OFF_QUEUE + (size_t)SCR_WARPS * SCR_Q_CAP * 8;
constexpr size_t SCR_SMEM_BYTES = SCR_OFF_RROW + (size_t)SCR_WARPS * SCR_RING;

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

__device__ __forceinline__ ScreenRec as_rec(const uint4 v) {
    ScreenRec r; r.x = v.x; r.y = v.y; r.z = v.z; r.w = v.w;
    return r;
}

// Work flow of one warp (all state below is warp-uniform unless it says "lane"):
//   refill   while fewer than SCR_REFILL hits are staged: take the next group of 32 seed words
//            (global counter), look their buckets up, align the query window of every seed word
//            that has hits into a row of the row ring, expand the buckets into the hit ring
//            (reference position + row id per hit).  Hits of different groups queue up behind
//            each other, so every round below has 32 hits until the very end of the call.
//   round    the 32 oldest staged hits, one per lane: their reference records were requested by
//            the previous round (cp.async into the staging buffer); the owner lanes pull them
//            into registers, the requests of the NEXT round go out, then the screen runs -- the
//            gather latency of round n+1 hides behind the arithmetic of round n.
//   drain    undecided hits (reference anchor + seed index) wait in a per-warp queue and are
//            tile-walked SCR_Q_DRAIN at a time.
template <int SRC>
__global__ void __launch_bounds__(FILTER_THREADS, 3)
k_filter_hits3(FilterParams P, ScreenConsts C, HitSource H, const int *__restrict__ sub_mat,
               SurvRec *__restrict__ surv, uint32_t surv_cap, uint32_t *__restrict__ counters) {
    static_assert(SRC == SRC_SEEDS || SRC == SRC_RANGE, "the screen needs the seed word of every hit");
    extern __shared__ __align__(16) unsigned char smem[];
    uint32_t *lut = reinterpret_cast<uint32_t *>(smem + SCR_OFF_LUT);
    __shared__ int diag[4];
    for (int i = threadIdx.x; i < SCR_LUT_WORDS; i += blockDim.x) {
        const int idx = i / SCR_LUT_COLS, rn = idx >> 4, qn = idx & 15;
        const int s0 = sub_mat[(rn & 3) * 8 + (qn & 3)], s1 = sub_mat[(rn >> 2) * 8 + (qn >> 2)];
        lut[i] = (uint32_t)(uint8_t)(int8_t)s0 | ((uint32_t)(uint8_t)(int8_t)s1 << 8);
    }
    if (threadIdx.x < 4) diag[threadIdx.x] = sub_mat[threadIdx.x * 9];
    __syncthreads();

    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;
    const uint32_t lut_lane = (uint32_t)__cvta_generic_to_shared(lut) + (lane & (uint32_t)(SCR_LUT_COLS - 1)) * 4u;
    const uint32_t mul = P.k_mul, m4 = P.k_m4;
    const int thr = P.hspthresh;

    uint4 *stage = reinterpret_cast<uint4 *>(smem + SCR_OFF_STAGE) + warp * 32 * SCR_STAGE_STRIDE;
    uint32_t *rows = reinterpret_cast<uint32_t *>(smem + SCR_OFF_ROWS) + warp * SCR_ROWS * SCR_ROW_STRIDE;
    uint32_t *ring_r = reinterpret_cast<uint32_t *>(smem + SCR_OFF_RING) + warp * SCR_RING;
    uint32_t *myq = reinterpret_cast<uint32_t *>(smem + SCR_OFF_QUEUE) + warp * SCR_Q_CAP * 2;
    uint8_t *ring_row = reinterpret_cast<uint8_t *>(smem + SCR_OFF_RROW) + warp * SCR_RING;
    const uint4 *rrec_m3 = P.rrec - 3; // record w-3 of a window (REC_FRONT >= 3 records of front padding)

    const uint32_t total_items = H.num_items;
    uint32_t key_base = 0, g_total = 0, g_done = 0, g_row_base = 0;
    uint32_t head = 0, tail = 0;   // hit ring: [head, tail) staged and not yet screened (monotonic counters)
    uint32_t row_tail = 0;         // rows handed out so far (monotonic; row id = counter mod 256)
    uint32_t pre_n = 0;            // hits whose records are already requested (the next round)
    bool exhausted = false;
    uint32_t acc_hits = 0, acc_seeds = 0, acc_last = 0;
    bool any_hits = false;
    uint32_t qcount = 0; // queued undecided hits
    uint32_t ext_tiles = 0, acc_walked = 0;

    auto emit = [&](uint32_t r0, uint32_t q0, uint32_t key) {
        const uint32_t slot = atomicAdd(counters + CTR_SURV, 1u);
        if (slot < surv_cap) { SurvRec rec; rec.r0 = r0; rec.q0 = q0; rec.key = key; surv[slot] = rec; }
    };
    // reference records w-3 .. w+2 of n staged hits starting at ring position `from`: six
    // neighbouring lanes per hit, 16 bytes each, straight into the staging buffer
    auto request_records = [&](uint32_t from, uint32_t n) {
#pragma unroll
        for (uint32_t i = 0; i < SCREEN_RECS; i++) {
            const uint32_t f = i * 32u + lane;
            const uint32_t hs = f / SCREEN_RECS, rc = f - hs * SCREEN_RECS;
            if (hs < n) {
                const uint32_t r = ring_r[(from + hs) & (SCR_RING - 1)];
                cp_async16(stage + hs * SCR_STAGE_STRIDE + rc, rrec_m3 + ((r >> 5) + rc));
            }
        }
    };

    for (;;) {
        // ---------------- refill the hit ring
        while (tail - head < (uint32_t)SCR_REFILL && !exhausted) {
            const bool new_group = g_done == g_total;
            if (new_group) {
                // a new group may take up to 32 rows: the rows of the unscreened hits must survive
                if (tail != head) {
                    const uint32_t live = (row_tail - ring_row[head & (SCR_RING - 1)]) & 0xFFu;
                    if (live > (uint32_t)(SCR_ROWS - 32)) break; // then more than 32 hits are staged: screen first
                }
                uint32_t c = 0;
                if (lane == 0) c = atomicAdd(counters + CTR_CHUNK, 1u);
                c = __shfl_sync(0xFFFFFFFFu, c, 0);
                const unsigned long long start = (unsigned long long)c * 32u;
                if (start >= total_items) { exhausted = true; break; }
                key_base = (uint32_t)start;
                g_done = 0;
            }
            // this lane's seed word and bucket (recomputed when a group is staged in several parts)
            const uint32_t k = key_base + lane;
            uint32_t b_start = 0, n = 0, qa = 0;
            bool valid = false;
            if (k < total_items) {
                uint32_t kmer = 0, qpos = 0;
                if (SRC == SRC_SEEDS) {
                    const uint64_t word = __ldg(H.seeds + k);
                    kmer = (uint32_t)(word >> 32); qpos = (uint32_t)word;
                    valid = true;
                } else {
                    const uint32_t pi = k / H.per, v = k - pi * H.per;
                    qpos = H.j0 + pi;
                    uint64_t W; uint32_t Tw, Sw;
                    load_window(P.qrec, (int)qpos, W, Tw, Sw);
                    const uint32_t span_mask = H.shape.span >= 32 ? 0xFFFFFFFFu : ((1u << H.shape.span) - 1u);
                    valid = ((Tw | Sw) & span_mask) == 0; // all span cells upper-case ACGT (ntcoding.cpp:47-52)
                    for (int i = 0; i < H.shape.weight; i++) kmer = (kmer << 2) | (uint32_t)((W >> (2 * H.shape.pos[i])) & 3u);
                    if (v > 0) kmer ^= 2u << (2 * H.shape.tvar[v - 1]); // seeder.cpp:64-71
                }
                if (valid) {
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
            const unsigned with_hits = __ballot_sync(0xFFFFFFFFu, n > 0);
            if (new_group) {
                g_total = __shfl_sync(0xFFFFFFFFu, incl, 31);
                acc_hits += g_total;
                acc_seeds += __popc(__ballot_sync(0xFFFFFFFFu, valid));
                if (with_hits) { acc_last = key_base + (31u - __clz(with_hits)); any_hits = true; } // groups come in ascending order per warp
                if (g_total == 0) continue;
                // aligned query window of every seed word that has hits (shared by all its hits)
                g_row_base = row_tail;
                if (n > 0) {
                    const uint4 *qr = P.qrec + (int)(qa >> 5) - 3;
                    const ScreenRec a[SCREEN_RECS] = {as_rec(__ldg(qr)), as_rec(__ldg(qr + 1)), as_rec(__ldg(qr + 2)),
                                                      as_rec(__ldg(qr + 3)), as_rec(__ldg(qr + 4)), as_rec(__ldg(qr + 5))};
                    uint32_t row[SCREEN_ROW_WORDS];
                    screen_align(a, qa & 31u, row);
                    const uint32_t slot = (row_tail + __popc(with_hits & lt_mask)) & (uint32_t)(SCR_ROWS - 1);
                    uint4 *dst = reinterpret_cast<uint4 *>(rows + slot * SCR_ROW_STRIDE);
                    dst[0] = make_uint4(row[0], row[1], row[2], row[3]);
                    dst[1] = make_uint4(row[4], row[5], row[6], row[7]);
                    dst[2] = make_uint4(row[8], row[9], row[10], k); // word 11 = seed order index of the row
                }
                row_tail += __popc(with_hits);
            }
            const uint32_t cnt = min(g_total - g_done, (uint32_t)SCR_RING - (tail - head));
            __syncwarp();
            for (uint32_t kk = 0; kk * 32u < cnt; kk++) {
                const uint32_t f = g_done + kk * 32u + lane; // flat index inside the group
                uint32_t lo = 0, hi = 31;                    // owner = first lane whose inclusive prefix exceeds f
#pragma unroll
                for (int it = 0; it < 5; it++) {
                    const uint32_t mid = (lo + hi) >> 1;
                    const uint32_t vmid = __shfl_sync(0xFFFFFFFFu, incl, mid);
                    if (vmid > f) hi = mid; else lo = mid + 1;
                }
                const uint32_t o_excl = __shfl_sync(0xFFFFFFFFu, excl, lo);
                const uint32_t o_start = __shfl_sync(0xFFFFFFFFu, b_start, lo);
                if (kk * 32u + lane < cnt) {
                    const uint32_t slot = (tail + kk * 32u + lane) & (uint32_t)(SCR_RING - 1);
                    ring_r[slot] = __ldg(H.pos_table + o_start + (f - o_excl)) + H.seed_size;
                    ring_row[slot] = (uint8_t)(g_row_base + __popc(with_hits & ((1u << lo) - 1u)));
                }
            }
            __syncwarp();
            g_done += cnt;
            tail += cnt;
        }
        if (tail == head && qcount == 0) break; // exhausted, everything screened and walked

        // ---------------- one round of the screen: the 32 oldest staged hits, one per lane
        if (tail != head) {
            const uint32_t n1 = pre_n ? pre_n : min(tail - head, 32u);
            if (!pre_n) request_records(head, n1);
            const bool have = lane < n1;
            uint32_t r0 = 0, rowid = 0;
            if (have) {
                const uint32_t slot = (head + lane) & (uint32_t)(SCR_RING - 1);
                r0 = ring_r[slot];
                rowid = ring_row[slot];
            }
            const uint4 *qrow = reinterpret_cast<const uint4 *>(rows + (rowid & (uint32_t)(SCR_ROWS - 1)) * SCR_ROW_STRIDE);
            const uint4 q_a = qrow[0], q_b = qrow[1], q_c = qrow[2];
            const uint32_t qr[SCREEN_ROW_WORDS] = {q_a.x, q_a.y, q_a.z, q_a.w, q_b.x, q_b.y, q_b.z, q_b.w, q_c.x, q_c.y, q_c.z, 0u};
            const uint32_t key = q_c.w;
            cp_async_wait_all();
            __syncwarp();
            uint32_t rr[SCREEN_ROW_WORDS];
            {
                const uint4 *mine = stage + lane * SCR_STAGE_STRIDE;
                const ScreenRec a[SCREEN_RECS] = {as_rec(mine[0]), as_rec(mine[1]), as_rec(mine[2]),
                                                  as_rec(mine[3]), as_rec(mine[4]), as_rec(mine[5])};
                screen_align(a, r0 & 31u, rr);
            }
            __syncwarp(); // every lane holds its window: the staging buffer is free again
            head += n1;
            pre_n = min(tail - head, 32u);
            if (pre_n) request_records(head, pre_n);
            int bound; bool decided;
            const bool push = have && !screen_reject(rr, qr, C, bound, decided);
            const unsigned pm = __ballot_sync(0xFFFFFFFFu, push);
            if (push) {
                const uint32_t idx = qcount + __popc(pm & lt_mask);
                myq[idx * 2 + 0] = r0; myq[idx * 2 + 1] = key;
            }
            qcount += __popc(pm);
            acc_walked += __popc(pm);
            __syncwarp();
        }

        // ---------------- tile walk of the undecided hits (persistent lanes, kernels_filter.cuh)
        if (qcount >= (uint32_t)SCR_Q_DRAIN || (qcount > 0 && tail == head && exhausted)) {
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
                        const uint32_t idx = qhead + rank;
                        r0 = myq[idx * 2 + 0]; key = myq[idx * 2 + 1];
                        const uint32_t qpos = SRC == SRC_RANGE ? H.j0 + key / H.per : (uint32_t)__ldg(H.seeds + key);
                        q0 = qpos + H.seed_size;
                        t = 0; s = 0; M = 0; right_score = 0; left = false;
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
    if (lane == 0) {
        if (acc_hits) atomicAdd(counters + CTR_NHITS, acc_hits); // uint32 wrap-around like the reference's scan
        if (acc_seeds) atomicAdd(counters + CTR_NSEEDS, acc_seeds);
        if (any_hits) atomicMax(counters + CTR_LASTKEY, acc_last);
        if (acc_walked) atomicAdd(counters + CTR_WALKED, acc_walked);
    }
}

} // namespace sa
