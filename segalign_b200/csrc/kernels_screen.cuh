// kernels_screen.cuh -- the default filter kernel: popcount screen first, tile walk for the rest.
//
// Same contract as k_filter_hits / k_filter_hits2 (kernels_filter.cuh): consumes every seed hit of
// a SeedAndFilter call (src/seed_filter.cu:157-230 fused with the bulk of :232-652), writes the
// hits that may reach hspthresh to the survivor list for the exact kernel.  What changed is where
// the time goes.  ncu of the tile-walk kernels (profiles/r1j_*) showed the L1/shared-memory pipe
// at 70 % and ~1400 instructions per hit: every lane gathered its own 16-byte records (one L1 tag
// lookup per lane and record, ~7 per hit) and walked ~3 tiles through the pair LUT.  Here
//
//   1. query windows are aligned once per seed word (seed vectors) or once per query position
//      (device seeding: the 13 words of a position share it), not per hit, and kept in shared
//      memory: all hits of a seed word share the query anchor;
//   2. the reference window of a hit (6 consecutive words of the bare 2-bit plane = 48 bytes; no
//      terminator / soft bits travel with it, see screen_align_p2 in screen_bound.h) is fetched by six
//      neighbouring lanes with cp.async straight into shared memory: a 32-lane request touches
//      ~7 lines instead of 32, the owner lane then reads its words with conflict-free LDS.128.
//      (Round 1 fetched 16-byte records, 96 bytes per hit: ncu at a 500 Mb reference block showed
//      15.9 GB of DRAM reads per launch = 163 bytes per hit, 87 % of the measured HBM peak, L2 hit
//      rate 29 % -- the 250 MB of records did not fit the L2; the 125 MB 2-bit plane does.);
//   3. the hit is decided by the popcount screen of screen_bound.h (~15 instructions per 16-cell
//      block, 10 blocks) -- about 98 % of random hits end here;
//   4. undecided hits are queued per warp and walked by the persistent-lane tile loop of
//      kernels_filter.cuh (same code, same bound as before), 32 to 63 at a time.
#pragma once
#include "kernels_filter.cuh"
#include "screen_bound.h"

namespace sa {

// tuning knobs (overridable at build time: scripts/experiments/tune_filter3.sh measures the variants on the GPU)
#ifndef SA_SCR_THREADS
#define SA_SCR_THREADS 256
#endif
#ifndef SA_SCR_MIN_CTAS
#define SA_SCR_MIN_CTAS 3
#endif
#ifndef SA_SCR_STAGE_STRIDE
#define SA_SCR_STAGE_STRIDE 4
#endif
#ifndef SA_SCR_Q_CAP
#define SA_SCR_Q_CAP 64
#endif
#ifndef SA_SCR_L2_HINTS
#define SA_SCR_L2_HINTS 1 // reference records evict_last, seed positions evict_first (keeps the records L2-resident)
#endif
constexpr int SCR_THREADS = SA_SCR_THREADS;
constexpr int SCR_STAGE_STRIDE = SA_SCR_STAGE_STRIDE; // uint4 slots per hit in the staging buffer: four 16-byte chunks, XOR-swizzled
static_assert(SCR_STAGE_STRIDE == 4, "the swizzle below assumes 64-byte staging rows");
constexpr int SCR_ROW_STRIDE = SCREEN_ROW_WORDS; // 48 bytes: conflict-free for 16-byte reads
#ifndef SA_SCR_RING
#define SA_SCR_RING 256
#endif
constexpr int SCR_RING = SA_SCR_RING; // staged hits per warp (ring, power of two)
constexpr int SCR_ROWS = 64;         // aligned query rows per warp (ring, power of two)
constexpr int SCR_REFILL = 96;       // refill the ring when fewer hits than this are staged (three rounds:
                                     // the positions of round n+1 were copied in before the refill of round n)
constexpr int SCR_Q_CAP = SA_SCR_Q_CAP;
constexpr int SCR_Q_DRAIN = SCR_Q_CAP - 32;
constexpr int SCR_WARPS = SCR_THREADS / 32;
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
constexpr size_t SCR_OFF_DELTA = SCR_OFF_RROW + (size_t)SCR_WARPS * SCR_RING * 2;
constexpr size_t SCR_SMEM_BYTES = SCR_OFF_DELTA + (size_t)SCR_WARPS * SCR_ROWS * 4;
static_assert((SCR_SMEM_BYTES + 1024) * SA_SCR_MIN_CTAS <= 233472, "k_filter_hits3: shared memory of the resident blocks exceeds an SM");
static_assert(SCR_Q_DRAIN >= 1 && SCR_Q_CAP >= SCR_Q_DRAIN + 31, "a round may queue 32 hits after the drain threshold was missed");

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async8(void *smem_dst, const void *gmem_src) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async8_hint(void *smem_dst, const void *gmem_src, uint64_t policy) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global.L2::cache_hint [%0], [%1], 8, %2;" ::"r"(d), "l"(gmem_src), "l"(policy) : "memory");
}
__device__ __forceinline__ void cp_async4(void *smem_dst, const void *gmem_src) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gmem_src) : "memory");
}
// L2 cache policies: the reference records are gathered again and again by every call of a block
// (50 MB at 100 Mb: they fit the L2), the seed positions stream through once per call
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void cp_async16_hint(void *smem_dst, const void *gmem_src, uint64_t policy) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(d), "l"(gmem_src), "l"(policy) : "memory");
}
__device__ __forceinline__ void cp_async4_hint(void *smem_dst, const void *gmem_src, uint64_t policy) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global.L2::cache_hint [%0], [%1], 4, %2;" ::"r"(d), "l"(gmem_src), "l"(policy) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Shared-memory accesses of the hot loop by 32-bit shared-space address (computed once per warp): through
// generic pointers ptxas re-derives the shared window base (S2R SR_CgaCtaId + LEA ...) at every access site,
// ~20 instructions per round.
__device__ __forceinline__ uint4 lds_v4(uint32_t a) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ uint32_t lds_b32(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ uint32_t lds_b16(uint32_t a) {
    uint16_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ uint32_t lds_b8(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts_b16(uint32_t a, uint32_t x) {
    asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "h"((uint16_t)x) : "memory");
}
__device__ __forceinline__ void sts_b16_if(uint32_t a, uint32_t x, bool ok) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %2, 0;\n\t@p st.shared.u16 [%0], %1;\n\t}" ::"r"(a), "h"((uint16_t)x), "r"((uint32_t)ok) : "memory");
}
__device__ __forceinline__ void cp_async4_s_if(uint32_t smem_dst, const void *gmem_src, bool ok) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %2, 0;\n\t@p cp.async.ca.shared.global [%0], [%1], 4;\n\t}"
                 ::"r"(smem_dst), "l"(gmem_src), "r"((uint32_t)ok) : "memory");
}
__device__ __forceinline__ void cp_async4_hint_s_if(uint32_t smem_dst, const void *gmem_src, uint64_t policy, bool ok) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %3, 0;\n\t@p cp.async.ca.shared.global.L2::cache_hint [%0], [%1], 4, %2;\n\t}"
                 ::"r"(smem_dst), "l"(gmem_src), "l"(policy), "r"((uint32_t)ok) : "memory");
}
__device__ __forceinline__ void cp_async4_s(uint32_t smem_dst, const void *gmem_src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_dst), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async4_hint_s(uint32_t smem_dst, const void *gmem_src, uint64_t policy) {
    asm volatile("cp.async.ca.shared.global.L2::cache_hint [%0], [%1], 4, %2;" ::"r"(smem_dst), "l"(gmem_src), "l"(policy) : "memory");
}
__device__ __forceinline__ void sts_v2(uint32_t a, uint32_t x, uint32_t y) {
    asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(a), "r"(x), "r"(y) : "memory");
}
// predicated form: no branch (and no convergence barrier pair) around the copy
__device__ __forceinline__ void cp_async16_s_if(uint32_t smem_dst, const void *gmem_src, bool ok) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %2, 0;\n\t@p cp.async.cg.shared.global [%0], [%1], 16;\n\t}"
                 ::"r"(smem_dst), "l"(gmem_src), "r"((uint32_t)ok) : "memory");
}
__device__ __forceinline__ void cp_async16_s(uint32_t smem_dst, const void *gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_dst), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ ScreenRec as_rec(const uint4 v) {
    ScreenRec r; r.x = v.x; r.y = v.y; r.z = v.z; r.w = v.w;
    return r;
}

// Work flow of one warp (all state below is warp-uniform unless it says "lane"):
//   refill   while fewer than SCR_REFILL hits are staged: take the next group (global counter,
//            fetched one refill ahead) -- 32 seed words of a seed vector, or 32 query positions
//            with device seeding, expanded one variant at a time -- look the buckets up, align the
//            query windows into rows of the row ring, expand the buckets into the hit ring
//            (reference position + row id [+ variant] per hit; the positions are
//            copied from the seed position table with cp.async and are first read two rounds
//            later).  Hits of different groups queue up behind each other, so every round below
//            has 32 hits until the very end of the call.
//   round    the 32 oldest staged hits, one per lane: their reference records were requested by
//            the previous round (cp.async into the staging buffer); the owner lanes pull them
//            into registers, the requests of the NEXT round go out, then the screen runs -- the
//            gather latency of round n+1 hides behind the arithmetic of round n.
//   drain    undecided hits (reference anchor + seed index) wait in a per-warp queue and are
//            tile-walked SCR_Q_DRAIN at a time.
template <int SRC>
__global__ void __launch_bounds__(SCR_THREADS, SA_SCR_MIN_CTAS)
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
    uint16_t *ring_meta = reinterpret_cast<uint16_t *>(smem + SCR_OFF_RROW) + warp * SCR_RING; // row id | variant << 8, per staged hit
    // per owner (seed word with hits) of the bucket set being expanded: bucket start - exclusive hit
    // prefix, and (device seeding) the row id of its query position
    uint32_t *ownerdelta = reinterpret_cast<uint32_t *>(smem + SCR_OFF_DELTA) + warp * SCR_ROWS;
    uint8_t *ownerrow = reinterpret_cast<uint8_t *>(ownerdelta + 32);
    const uint32_t sa_stage = (uint32_t)__cvta_generic_to_shared(stage), sa_rows = (uint32_t)__cvta_generic_to_shared(rows),
                   sa_ring = (uint32_t)__cvta_generic_to_shared(ring_r), sa_meta = (uint32_t)__cvta_generic_to_shared(ring_meta),
                   sa_q = (uint32_t)__cvta_generic_to_shared(myq), sa_delta = (uint32_t)__cvta_generic_to_shared(ownerdelta);
    const uint64_t *rp2_m4 = P.rp2 - 4; // 16-byte aligned start of the plane's front padding (REC_FRONT = 4 words)
    static_assert(REC_FRONT == 4, "the window fetch below starts at word (w + 1) & ~1 of the padded plane");
#if SA_SCR_L2_HINTS
    const uint64_t pol_stream = l2_policy_evict_first();
#endif

    const uint32_t n_var = SRC == SRC_RANGE ? H.per : 1u;       // bucket sets per group (device seeding: 1 + transition variants)
    const uint32_t n_units = SRC == SRC_RANGE ? H.num_items / n_var : H.num_items; // query positions / seed words
    uint32_t key_base = 0, g_total = 0, g_done = 0, g_row_base = 0;
    uint32_t v_cur = 0, v_next = n_var, valid_mask = 0; // device seeding: variant of the current set, lanes with a valid position
    uint32_t lane_kmer = 0;        // lane: exact k-mer of its query position (device seeding)
    uint32_t pf_v = 0xFFFFFFFFu, pf_end = 0, pf_start = 0; // lane: index entries requested ahead for variant pf_v of the current positions
    uint32_t cur_end = 0, cur_start = 0;                   // lane: index entries of the set being staged
    bool lane_valid = false;
    uint32_t head = 0, tail = 0;   // hit ring: [head, tail) staged and not yet screened (monotonic counters)
    uint32_t next_c = 0;           // lane 0: the next group number, fetched one refill ahead
    if (lane == 0) next_c = atomicAdd(counters + CTR_CHUNK, 1u);
    uint32_t row_tail = 0;         // rows handed out so far (monotonic; row id = counter mod 256)
    uint32_t pre_n = 0;            // hits whose records are already requested (the next round)
    uint32_t r_next = 0;           // lane: reference anchor of its hit of that round
    bool soft_next = false;        // lane: a soft cell lies in that hit's reference records (blocks with soft cells only)
    bool exhausted = false;
    uint32_t acc_hits = 0, acc_seeds = 0, acc_last = 0;
    bool any_hits = false;
    uint32_t qcount = 0; // queued undecided hits
    uint32_t ext_tiles = 0, acc_walked = 0;

    // query anchor of a hit from the order index of its seed word
    auto query_anchor = [&](uint32_t key) -> uint32_t {
        const uint32_t qpos = SRC == SRC_RANGE ? H.j0 + key / H.per : (uint32_t)__ldg(H.seeds + key);
        return qpos + H.seed_size;
    };
    // Reference words w-3 .. w+2 (2-bit plane, 48 bytes) of the next round's n hits, fetched as 16-byte
    // chunks with cp.async.cg (L1 bypass: an 8-byte .ca copy allocates an L1 line per miss and the ~27 KB
    // of L1 left beside the shared memory cannot hold the lines of 24 warps' gathers in flight -- measured
    // 34 % slower at a 500 Mb block).  The fetch starts at the even word at or below w-3: three chunks when
    // w-3 is even, four when it is odd; the owner lane drops the leading word again (`par` below).
    // Four neighbouring lanes per hit; r_own = this lane's own hit of that round (anchor in the reference
    // block), the loader lanes get the anchors by shuffle.  Chunk c of hit h lands in slot c ^ ((h >> 1) & 3)
    // of the hit's 64-byte row: writes (two rows per quarter warp) and the owners' reads (same chunk of
    // eight rows) are both bank-conflict free without padding.
    auto request_records = [&](uint32_t r_own, uint32_t n) {
        const uint32_t w1_own = (r_own >> 5) + 1u; // = (w - 3) + REC_FRONT of this lane's own hit
        const uint32_t rc = lane & 3u, h0 = lane >> 2;
        // hit h = 8 i + h0 in iteration i: (h >> 1) & 3 = (lane >> 3) & 3 for every i
        const uint32_t dst0 = sa_stage + (h0 * SCR_STAGE_STRIDE + (rc ^ ((lane >> 3) & 3u))) * 16u;
#pragma unroll
        for (uint32_t i = 0; i < 4; i++) {
            const uint32_t hs = i * 8u + h0;
            const uint32_t w1 = __shfl_sync(0xFFFFFFFFu, w1_own, hs);
            const uint64_t *src = rp2_m4 + ((w1 & ~1u) + 2u * rc);
            // no L2 policy operand here: with `.L2::cache_hint` ptxas 12.9 emits, for the three copies whose
            // shared address is the first one's plus an immediate, an LDGSTS form that reads an unset uniform
            // register pair as descriptor (cuobjdump: `[R41+UR0+0x200], desc[UR1]`) -- the launch dies with
            // "illegal instruction".  The hint bought nothing measurable.
            cp_async16_s_if(dst0 + i * (8u * SCR_STAGE_STRIDE * 16u), src, hs < n && (rc < 3u || (w1 & 1u)));
        }
    };

    for (;;) {
        // ---------------- refill the hit ring
        const uint32_t tail_prev = tail; // positions below this were copied in by an earlier refill
        while (tail - head < (uint32_t)SCR_REFILL && !exhausted) {
            // A "bucket set" = 32 buckets, one per lane, expanded together.  Seed vectors (SRC_SEEDS): the
            // buckets of 32 consecutive seed words.  Device seeding (SRC_RANGE): the buckets of ONE
            // variant (exact word, or the transition at one care position) of 32 consecutive query
            // positions -- window, k-mer and aligned query row are then computed once per position
            // instead of once per seed word, and the index lookups of a set are independent loads.
            const bool new_set = g_done == g_total;
            if (new_set && (SRC == SRC_SEEDS || v_next == n_var)) {
                // a new group may take up to 32 rows: the rows of the unscreened hits must survive
                if (tail != head) {
                    const uint32_t live = (row_tail - (ring_meta[head & (SCR_RING - 1)] & 0xFFu)) & 0xFFu;
                    if (live > (uint32_t)(SCR_ROWS - 32)) break; // then more than 32 hits are staged: screen first
                }
                const uint32_t c = __shfl_sync(0xFFFFFFFFu, next_c, 0);
                const unsigned long long start = (unsigned long long)c * 32u;
                if (start >= n_units) { exhausted = true; break; }
                if (lane == 0) next_c = atomicAdd(counters + CTR_CHUNK, 1u);
                key_base = (uint32_t)start;
                if (SRC == SRC_RANGE) {
                    // this lane's query position: validity, exact k-mer, aligned row
                    const uint32_t pi = key_base + lane;
                    lane_valid = false; lane_kmer = 0;
                    uint32_t qa = 0;
                    if (pi < n_units) {
                        const uint32_t qpos = H.j0 + pi;
                        uint64_t W; uint32_t Tw, Sw;
                        load_window(P.qrec, (int)qpos, W, Tw, Sw);
                        const uint32_t span_mask = H.shape.span >= 32 ? 0xFFFFFFFFu : ((1u << H.shape.span) - 1u);
                        lane_valid = ((Tw | Sw) & span_mask) == 0; // all span cells upper-case ACGT (ntcoding.cpp:47-52)
                        for (int i = 0; i < H.shape.weight; i++) lane_kmer = (lane_kmer << 2) | (uint32_t)((W >> (2 * H.shape.pos[i])) & 3u);
                        qa = qpos + H.seed_size;
                    }
                    valid_mask = __ballot_sync(0xFFFFFFFFu, lane_valid);
                    acc_seeds += __popc(valid_mask) * n_var;
                    g_row_base = row_tail;
                    if (lane_valid) {
                        const uint4 *qr = P.qrec + (int)(qa >> 5) - 3;
                        const ScreenRec a[SCREEN_RECS] = {as_rec(__ldg(qr)), as_rec(__ldg(qr + 1)), as_rec(__ldg(qr + 2)),
                                                          as_rec(__ldg(qr + 3)), as_rec(__ldg(qr + 4)), as_rec(__ldg(qr + 5))};
                        uint32_t row[SCREEN_ROW_WORDS];
                        screen_align(a, qa & 31u, row);
                        const uint32_t slot = (row_tail + __popc(valid_mask & lt_mask)) & (uint32_t)(SCR_ROWS - 1);
                        uint4 *dst = reinterpret_cast<uint4 *>(rows + slot * SCR_ROW_STRIDE);
                        dst[0] = make_uint4(row[0], row[1], row[2], row[3]);
                        dst[1] = make_uint4(row[4], row[5], row[6], row[7]);
                        dst[2] = make_uint4(row[8], row[9], row[10], pi * n_var); // word 11 = seed order index of the exact word
                    }
                    row_tail += __popc(valid_mask);
                    v_next = valid_mask ? 0u : n_var; // a group without a valid position has no bucket sets
                    if (!valid_mask) continue;
                }
            }
            if (SRC == SRC_RANGE && new_set) v_cur = v_next++;
            // this lane's bucket of the set (recomputed when a set is staged in several parts)
            uint32_t b_start = 0, n = 0, qa = 0;
            bool valid = false;
            if (SRC == SRC_SEEDS) {
                const uint32_t k = key_base + lane;
                if (k < n_units) {
                    const uint64_t word = __ldg(H.seeds + k);
                    const uint32_t kmer = (uint32_t)(word >> 32);
                    valid = true;
                    // a caller-supplied word outside the table / the query block has an empty bucket
                    if (kmer < H.index_size && (unsigned long long)(uint32_t)word + H.seed_size <= H.query_len) {
                        const uint32_t b_end = __ldg(H.index_table + kmer);
                        b_start = kmer > 0 ? __ldg(H.index_table + kmer - 1) : 0u;
                        n = b_end - b_start;
                    }
                    qa = (uint32_t)word + H.seed_size;
                }
            } else if (lane_valid) {
                // The two index entries of this lane's seed word.  They are random 4-byte reads of a 64 MiB table whose
                // result the very next instruction needs (ncu: 6.8 % of the kernel's stall samples sat on that
                // subtraction), so the entries of the NEXT variant of the same 32 positions are requested here and
                // used one bucket set later; a set staged in several parts keeps its own entries in registers.
                uint32_t b_end;
                if (new_set) {
                    if (pf_v == v_cur) { b_end = pf_end; b_start = pf_start; }
                    else {
                        const uint32_t kmer = v_cur ? lane_kmer ^ (2u << (2 * H.shape.tvar[v_cur - 1])) : lane_kmer; // seeder.cpp:64-71
                        b_end = __ldg(H.index_table + kmer);
                        b_start = kmer > 0 ? __ldg(H.index_table + kmer - 1) : 0u;
                    }
                    cur_end = b_end; cur_start = b_start;
                } else {
                    b_end = cur_end; b_start = cur_start;
                }
                n = b_end - b_start;
            }
            if (SRC == SRC_RANGE && new_set) {
                pf_v = 0xFFFFFFFFu;
                if (v_cur + 1u < n_var) {
                    pf_v = v_cur + 1u;
                    if (lane_valid) {
                        const uint32_t kmer = lane_kmer ^ (2u << (2 * H.shape.tvar[v_cur])); // variant v_cur + 1
                        pf_end = __ldg(H.index_table + kmer);
                        pf_start = kmer > 0 ? __ldg(H.index_table + kmer - 1) : 0u;
                    }
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
            if (new_set) {
                g_done = 0;
                g_total = __shfl_sync(0xFFFFFFFFu, incl, 31);
                acc_hits += g_total;
                if (SRC == SRC_SEEDS) acc_seeds += __popc(__ballot_sync(0xFFFFFFFFu, valid));
                if (with_hits) { // seed order index of the last hit-bearing seed word seen by this warp
                    const uint32_t top = key_base + (31u - __clz(with_hits));
                    acc_last = max(acc_last, SRC == SRC_SEEDS ? top : top * n_var + v_cur);
                    any_hits = true;
                }
                if (g_total == 0) continue;
                if (SRC == SRC_SEEDS) {
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
                        dst[2] = make_uint4(row[8], row[9], row[10], key_base + lane); // word 11 = seed order index of the row
                    }
                    row_tail += __popc(with_hits);
                }
            }
            // owner tables of this set: the j-th lane with hits -> position-table offset and row id
            if (n > 0) {
                const uint32_t j = __popc(with_hits & lt_mask);
                ownerdelta[j] = b_start - excl; // hit f of the set sits at pos_table[f + this]
                ownerrow[j] = (uint8_t)(g_row_base + (SRC == SRC_SEEDS ? j : __popc(valid_mask & lt_mask)));
            }
            const uint32_t cnt = min(g_total - g_done, (uint32_t)SCR_RING - (tail - head));
            const uint32_t vtag = SRC == SRC_RANGE ? v_cur << 8 : 0u;
            __syncwarp();
            for (uint32_t kk = 0; kk * 32u < cnt; kk++) {
                // owner of hit f = the j-th lane with hits, j = #{lanes with hits whose inclusive prefix is
                // <= f}: one ballot for the 32 hits' common base, one OR-reduction of the prefix
                // boundaries that fall inside these 32 hits, one popcount per lane
                const uint32_t f0 = g_done + kk * 32u;
                const uint32_t j0 = __popc(__ballot_sync(0xFFFFFFFFu, n > 0 && incl <= f0));
                const uint32_t bp = incl - f0 - 1u;
                const uint32_t bounds = __reduce_or_sync(0xFFFFFFFFu, (n > 0 && incl > f0 && bp < 32u) ? 1u << bp : 0u);
                const uint32_t j = j0 + __popc(bounds & lt_mask);
                {   // branch-free: lanes past the last hit of the set run the same instructions with the copy and
                    // the store predicated off (j <= 32 keeps the table reads inside the warp's slice)
                    const bool ok = kk * 32u + lane < cnt;
                    const uint32_t slot = (tail + kk * 32u + lane) & (uint32_t)(SCR_RING - 1);
                    const uint32_t od = lds_b32(sa_delta + j * 4u), orow = lds_b8(sa_delta + 128u + j);
#if SA_SCR_L2_HINTS
                    cp_async4_hint_s_if(sa_ring + slot * 4u, H.pos_table + (od + f0 + lane), pol_stream, ok);
#else
                    cp_async4_s_if(sa_ring + slot * 4u, H.pos_table + (od + f0 + lane), ok);
#endif
                    sts_b16_if(sa_meta + slot * 2u, orow | vtag, ok);
                }
            }
            __syncwarp();
            g_done += cnt;
            tail += cnt;
        }
        cp_async_commit(); // group "A": the positions staged by this refill (possibly none)
        if (tail == head && qcount == 0) break; // exhausted, everything screened and walked

        // ---------------- one round of the screen: the 32 oldest staged hits, one per lane
        if (tail != head) {
            // cp.async groups in issue order: ... R(this round) A(this refill) | R(next round) ...
            uint32_t n1 = pre_n, safe_tail = tail_prev;
            if (!pre_n) { // cold: nothing requested yet (start of the call, or the ring ran dry)
                cp_async_wait_all();
                __syncwarp();
                n1 = min(tail - head, 32u);
                safe_tail = tail;
                r_next = lane < n1 ? lds_b32(sa_ring + ((head + lane) & (uint32_t)(SCR_RING - 1)) * 4u) + H.seed_size : 0u;
                if (P.ref_has_soft) soft_next = lane < n1 && soft_window(P.rsoft, r_next);
                request_records(r_next, n1);
                cp_async_commit();
                cp_async_wait_all();
            } else {
                cp_async_wait_group<1>(); // everything but this refill's positions has landed
            }
            __syncwarp(); // ... for every lane: a window is written by four other lanes' copies
            // Order of the round: the shared-memory loads whose results are needed first are issued first and
            // the NEXT round's gathers leave as soon as this round's staged windows sit in registers -- before the
            // ~45 instructions of window alignment, not after them (ncu at a 500 Mb block: 13 % of the round's
            // stall samples sat on the wait above, the gathers of a round being only ~250 instructions ahead).
            const uint32_t r0 = r_next;      // this lane's anchor: read when its window was requested
            const bool soft0 = soft_next;
            // (a hit outside the caller's reference window -- repeat-masker variant -- is counted, not screened)
            const bool have = lane < n1 && r0 - H.win_lo <= H.win_hi - H.win_lo;
            const uint32_t head_next = head + n1;
            const uint32_t pre_next = min(safe_tail - head_next, 32u); // only positions that are known to have landed
            uint32_t r_nn = 0;
            if (lane < pre_next) r_nn = lds_b32(sa_ring + ((head_next + lane) & (uint32_t)(SCR_RING - 1)) * 4u);
            uint32_t rowid = 0, vkey = 0;
            if (have) {
                const uint32_t meta = lds_b16(sa_meta + ((head + lane) & (uint32_t)(SCR_RING - 1)) * 2u);
                rowid = meta & 0xFFu;
                vkey = meta >> 8;
            }
            const uint32_t mine = sa_stage + lane * (SCR_STAGE_STRIDE * 16u);
            const uint32_t sw = (lane >> 1) & 3u;
            // explicit 128-bit loads: left to itself the compiler reads the seven 64-bit words it needs with
            // LDS.64, and lanes L and L+8 then meet in one bank pair (the swizzle is per 16-byte chunk)
            const uint4 m0 = lds_v4(mine + (0u ^ sw) * 16u), m1 = lds_v4(mine + (1u ^ sw) * 16u), m2 = lds_v4(mine + (2u ^ sw) * 16u),
                        m3 = lds_v4(mine + (3u ^ sw) * 16u);
            __syncwarp(); // every lane holds its window: the staging buffer is free again
            head = head_next;
            pre_n = pre_next;
            r_next = lane < pre_n ? r_nn + H.seed_size : 0u;
            if (P.ref_has_soft) soft_next = lane < pre_n && soft_window(P.rsoft, r_next);
            request_records(r_next, pre_n); // (pre_n == 0: every copy predicated off)
            cp_async_commit(); // group "R": the records of the next round
            const uint32_t qrow = sa_rows + (rowid & (uint32_t)(SCR_ROWS - 1)) * (SCR_ROW_STRIDE * 4u);
            const uint4 q_a = lds_v4(qrow), q_b = lds_v4(qrow + 16u), q_c = lds_v4(qrow + 32u);
            const uint32_t qr[SCREEN_ROW_WORDS] = {q_a.x, q_a.y, q_a.z, q_a.w, q_b.x, q_b.y, q_b.z, q_b.w, q_c.x, q_c.y, q_c.z, 0u};
            const uint32_t key = q_c.w + vkey; // seed order index of the hit's seed word
            uint32_t rr[SCREEN_ROW_WORDS];
            {
                const uint64_t c[8] = {(uint64_t)m0.x | ((uint64_t)m0.y << 32), (uint64_t)m0.z | ((uint64_t)m0.w << 32),
                                       (uint64_t)m1.x | ((uint64_t)m1.y << 32), (uint64_t)m1.z | ((uint64_t)m1.w << 32),
                                       (uint64_t)m2.x | ((uint64_t)m2.y << 32), (uint64_t)m2.z | ((uint64_t)m2.w << 32),
                                       (uint64_t)m3.x | ((uint64_t)m3.y << 32), (uint64_t)m3.z | ((uint64_t)m3.w << 32)};
                const bool par = (((r0 >> 5) + 1u) & 1u) != 0; // the fetch started one word below w-3
                uint64_t a[SCREEN_RECS];
#pragma unroll
                for (int j = 0; j < SCREEN_RECS; j++) a[j] = par ? c[j + 1] : c[j];
                screen_align_p2(a, r0 & 31u, soft0, rr);
            }
            int bound; bool decided;
            const bool push = have && !screen_reject(rr, qr, C, bound, decided);
            const unsigned pm = __ballot_sync(0xFFFFFFFFu, push);
            if (push) sts_v2(sa_q + (qcount + __popc(pm & lt_mask)) * 8u, r0, key);
            qcount += __popc(pm);
            acc_walked += __popc(pm);
            __syncwarp();
        }

        // ---------------- tile walk of the undecided hits (persistent lanes, kernels_filter.cuh)
        if (qcount >= (uint32_t)SCR_Q_DRAIN || (qcount > 0 && tail == head && exhausted)) {
            uint32_t qhead = 0, nsurv = 0;
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
                        q0 = query_anchor(key);
                        t = 0; s = 0; M = 0; right_score = 0; left = false;
                        active = true;
                    }
                    const uint32_t nneed = __popc(need);
                    qhead += nneed < avail ? nneed : avail;
                }
                if (!__any_sync(0xFFFFFFFFu, active)) break;
                bool keep = false;
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
                        keep = true;
                        active = false;
                    } else if (done) {
                        if (!left) { right_score = M; left = true; t = 0; s = 0; M = 0; }
                        else active = false;
                    } else {
                        t += 32u;
                    }
                }
                // survivors go back into the part of the queue that has already been handed out
                // (one global atomic per drain instead of one per survivor)
                const unsigned km = __ballot_sync(0xFFFFFFFFu, keep);
                if (keep) {
                    const uint32_t idx = nsurv + __popc(km & lt_mask);
                    myq[idx * 2 + 0] = r0; myq[idx * 2 + 1] = key;
                }
                nsurv += __popc(km);
            }
            __syncwarp();
            if (nsurv) {
                uint32_t base = 0;
                if (lane == 0) base = atomicAdd(counters + CTR_SURV, nsurv);
                base = __shfl_sync(0xFFFFFFFFu, base, 0);
                for (uint32_t i = lane; i < nsurv; i += 32u) {
                    SurvRec rec;
                    rec.r0 = myq[i * 2 + 0]; rec.key = myq[i * 2 + 1]; rec.q0 = query_anchor(rec.key);
                    if (base + i < surv_cap) surv[base + i] = rec;
                }
            }
            qcount = 0;
            __syncwarp();
        }
    }
    if (ext_tiles) atomicAdd(reinterpret_cast<unsigned long long *>(counters + CTR_EXT_LO), 32ull * ext_tiles);
    if (lane == 0) {
        if (acc_hits) {
            atomicAdd(counters + CTR_NHITS, acc_hits); // uint32 wrap-around like the reference's scan
            atomicAdd(reinterpret_cast<unsigned long long *>(counters + CTR_NHITS64), (unsigned long long)acc_hits);
        }
        if (acc_seeds) atomicAdd(counters + CTR_NSEEDS, acc_seeds);
        if (any_hits) atomicMax(counters + CTR_LASTKEY, acc_last);
        if (acc_walked) atomicAdd(counters + CTR_WALKED, acc_walked);
    }
}

} // namespace sa
