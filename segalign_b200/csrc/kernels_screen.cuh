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
constexpr int SCR_Q_CAP = 96;
constexpr int SCR_Q_DRAIN = 64;
constexpr int SCR_WARPS = FILTER_THREADS / 32;

// dynamic shared memory layout of k_filter_hits3 (bytes)
constexpr size_t SCR_OFF_LUT = 0;
constexpr size_t SCR_OFF_STAGE = SCR_OFF_LUT + FILTER_LUT_WORDS * 4;
constexpr size_t SCR_OFF_ROWS = SCR_OFF_STAGE + (size_t)SCR_WARPS * 32 * SCR_STAGE_STRIDE * 16;
constexpr size_t SCR_OFF_HITS = SCR_OFF_ROWS + (size_t)SCR_WARPS * 32 * SCR_ROW_STRIDE * 4;
constexpr size_t SCR_OFF_QUEUE = SCR_OFF_HITS + (size_t)SCR_WARPS * FILTER_CHUNK * 8;
constexpr size_t SCR_OFF_OWN = SCR_OFF_QUEUE + (size_t)SCR_WARPS * SCR_Q_CAP * 12;
constexpr size_t SCR_SMEM_BYTES = SCR_OFF_OWN + (size_t)SCR_WARPS * FILTER_CHUNK;

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

__device__ __forceinline__ ScreenRec as_rec(const uint4 v) {
    ScreenRec r; r.x = v.x; r.y = v.y; r.z = v.z; r.w = v.w;
    return r;
}

template <int SRC>
__global__ void __launch_bounds__(FILTER_THREADS, 3)
k_filter_hits3(FilterParams P, ScreenConsts C, HitSource H, const int *__restrict__ sub_mat,
               SurvRec *__restrict__ surv, uint32_t surv_cap, uint32_t *__restrict__ counters) {
    static_assert(SRC == SRC_SEEDS || SRC == SRC_RANGE, "the screen needs the seed word of every hit");
    extern __shared__ __align__(16) unsigned char smem[];
    uint32_t *lut = reinterpret_cast<uint32_t *>(smem + SCR_OFF_LUT);
    __shared__ int diag[4];
    for (int i = threadIdx.x; i < FILTER_LUT_WORDS; i += blockDim.x) {
        const int idx = i >> 4, rn = idx >> 4, qn = idx & 15;
        const int s0 = sub_mat[(rn & 3) * 8 + (qn & 3)], s1 = sub_mat[(rn >> 2) * 8 + (qn >> 2)];
        lut[i] = (uint32_t)(uint8_t)(int8_t)s0 | ((uint32_t)(uint8_t)(int8_t)s1 << 8);
    }
    if (threadIdx.x < 4) diag[threadIdx.x] = sub_mat[threadIdx.x * 9];
    __syncthreads();

    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;
    const uint32_t lut_lane = (uint32_t)__cvta_generic_to_shared(lut) + (lane & 15u) * 4u;
    const uint32_t mul = P.k_mul, m4 = P.k_m4;
    const int thr = P.hspthresh;

    uint4 *stage = reinterpret_cast<uint4 *>(smem + SCR_OFF_STAGE) + warp * 32 * SCR_STAGE_STRIDE;
    uint32_t *rows = reinterpret_cast<uint32_t *>(smem + SCR_OFF_ROWS) + warp * 32 * SCR_ROW_STRIDE;
    uint2 *mybuf = reinterpret_cast<uint2 *>(smem + SCR_OFF_HITS) + warp * FILTER_CHUNK;
    uint32_t *myq = reinterpret_cast<uint32_t *>(smem + SCR_OFF_QUEUE) + warp * SCR_Q_CAP * 3;
    uint8_t *myown = reinterpret_cast<uint8_t *>(smem + SCR_OFF_OWN) + warp * FILTER_CHUNK;

    const uint32_t total_items = H.num_items;
    uint32_t key_base = 0, cursor = 0, limit = 0, g_total = 0, g_done = 0; // as in k_filter_hits
    bool exhausted = false;
    uint32_t acc_hits = 0, acc_seeds = 0, acc_last = 0;
    bool any_hits = false;
    uint32_t qcount = 0; // warp-uniform: queued undecided hits
    uint32_t ext_tiles = 0, acc_walked = 0;

    auto emit = [&](uint32_t r0, uint32_t q0, uint32_t key) {
        const uint32_t slot = atomicAdd(counters + CTR_SURV, 1u);
        if (slot < surv_cap) { SurvRec rec; rec.r0 = r0; rec.q0 = q0; rec.key = key; surv[slot] = rec; }
    };

    for (;;) {
        // ---------------- stage the next chunk of fresh hits
        while (cursor == limit && !exhausted) {
            const bool new_group = g_done == g_total;
            if (new_group) { // next group of 32 seed words
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
            if (new_group) {
                g_total = __shfl_sync(0xFFFFFFFFu, incl, 31);
                acc_hits += g_total;
                acc_seeds += __popc(__ballot_sync(0xFFFFFFFFu, valid));
                const unsigned with_hits = __ballot_sync(0xFFFFFFFFu, n > 0);
                if (with_hits) { acc_last = key_base + (31u - __clz(with_hits)); any_hits = true; }
                if (g_total == 0) continue;
                // aligned query window of every seed word that has hits (shared by all its hits)
                if (n > 0) {
                    const uint4 *qr = P.qrec + (int)(qa >> 5) - 3;
                    const ScreenRec a[SCREEN_RECS] = {as_rec(__ldg(qr)), as_rec(__ldg(qr + 1)), as_rec(__ldg(qr + 2)),
                                                      as_rec(__ldg(qr + 3)), as_rec(__ldg(qr + 4)), as_rec(__ldg(qr + 5))};
                    uint32_t row[SCREEN_ROW_WORDS];
                    screen_align(a, qa & 31u, row);
                    uint4 *dst = reinterpret_cast<uint4 *>(rows + lane * SCR_ROW_STRIDE);
                    dst[0] = make_uint4(row[0], row[1], row[2], row[3]);
                    dst[1] = make_uint4(row[4], row[5], row[6], row[7]);
                    dst[2] = make_uint4(row[8], row[9], row[10], row[11]);
                }
            }
            const uint32_t cnt = min(g_total - g_done, FILTER_CHUNK);
            __syncwarp();
#pragma unroll
            for (uint32_t kk = 0; kk < FILTER_CHUNK / 32; kk++) {
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
        const bool fresh = cursor < limit;
        if (!fresh && qcount == 0) break; // exhausted and nothing queued

        // ---------------- screen: 32 fresh hits, one per lane
        if (fresh) {
            const uint32_t n1 = min(limit - cursor, 32u);
            // reference records w-3 .. w+2 of every hit, six neighbouring lanes per hit
#pragma unroll
            for (uint32_t i = 0; i < SCREEN_RECS; i++) {
                const uint32_t f = i * 32u + lane;
                const uint32_t hs = f / SCREEN_RECS, rc = f - hs * SCREEN_RECS;
                if (hs < n1) {
                    const uint32_t r = mybuf[cursor + hs].x;
                    cp_async16(stage + hs * SCR_STAGE_STRIDE + rc, P.rrec + (int)(r >> 5) - 3 + (int)rc);
                }
            }
            const bool have = lane < n1;
            uint32_t r0 = 0, q0 = 0, own = 0;
            if (have) {
                const uint2 hit = mybuf[cursor + lane];
                r0 = hit.x; q0 = hit.y;
                own = myown[cursor + lane];
            }
            cursor += n1;
            const uint4 *qrow = reinterpret_cast<const uint4 *>(rows + own * SCR_ROW_STRIDE);
            const uint4 q_a = qrow[0], q_b = qrow[1], q_c = qrow[2];
            const uint32_t qr[SCREEN_ROW_WORDS] = {q_a.x, q_a.y, q_a.z, q_a.w, q_b.x, q_b.y, q_b.z, q_b.w, q_c.x, q_c.y, q_c.z, q_c.w};
            cp_async_wait_all();
            __syncwarp();
            bool push = false;
            if (have) {
                const uint4 *mine = stage + lane * SCR_STAGE_STRIDE;
                const ScreenRec a[SCREEN_RECS] = {as_rec(mine[0]), as_rec(mine[1]), as_rec(mine[2]),
                                                  as_rec(mine[3]), as_rec(mine[4]), as_rec(mine[5])};
                uint32_t rr[SCREEN_ROW_WORDS];
                screen_align(a, r0 & 31u, rr);
                int bound; bool decided;
                push = !screen_reject(rr, qr, C, bound, decided);
            }
            const unsigned pm = __ballot_sync(0xFFFFFFFFu, push);
            if (push) {
                const uint32_t idx = qcount + __popc(pm & lt_mask);
                myq[idx * 3 + 0] = r0; myq[idx * 3 + 1] = q0; myq[idx * 3 + 2] = key_base + own;
            }
            qcount += __popc(pm);
            acc_walked += __popc(pm);
            __syncwarp(); // staging buffer and queue are reused / read below
        }

        // ---------------- tile walk of the undecided hits (persistent lanes, kernels_filter.cuh)
        if (qcount >= (uint32_t)SCR_Q_DRAIN || (qcount > 0 && cursor == limit && exhausted)) {
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
                        r0 = myq[idx * 3 + 0]; q0 = myq[idx * 3 + 1]; key = myq[idx * 3 + 2];
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
