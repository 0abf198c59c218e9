// kernels_merge.cuh -- survivors that provably extend to the same HSP are extended once.
//
// Every seed hit inside one conserved run extends to the same HSP, and the reference extends each of
// them (src/seed_filter.cu:232-652) only to throw the copies away in its sort/unique pass
// (:776-782).  For diverged genomes the copies are few and the exact-duplicate table of
// kernels_extend.cuh handles them after the fact.  A SELF-alignment (BASELINE configs[0]) is the
// degenerate case: every query position hits the main diagonal, every one of those 4.6 M hits walks
// the whole 4.6 Mb diagonal -- 2*10^13 cell steps for ONE HSP (268 s with the reference's kernels on a
// B200, 13 s with the warp-per-hit kernel of kernels_extend_wide.cuh).
//
// Two anchors a < a' on one diagonal give the same HSP if every cell in [a, a') is a match of two
// upper-case ACGT bases and every diagonal matrix entry is positive (SURVEY A.5 semantics):
//   right walk from a : the prefix sum rises strictly over [a, a'), so at cell a'-1 the running maximum
//     equals the running sum and sits at a'-1; from a' on the walk sees exactly what the walk from a'
//     sees (stop rule and strict-maximum rule depend only on running sum minus running maximum), so
//     both end at the same absolute cell, with scores that differ by sum[a, a');
//   left walk from a' : mirrored -- it rises strictly over a'-1 .. a, then continues as the walk from
//     a does; same absolute start, scores differ by the same sum[a, a').
//   total = right + left is equal, and so are ref_start / query_start / len; the entropy counters are a
//   function of the final HSP.  (Cells past a block end score 0 and only end a walk: no effect.)
// The relation is transitive along a diagonal, so after sorting the survivors of a call by
// (diagonal, anchor) every survivor that is connected to its predecessor is dropped: its
// representative -- the first anchor of the chain -- yields the record it would have yielded, and the
// reference's unique pass would have kept one copy of it anyway (A.8).  Survivors of the call's last
// hit-bearing seed word (the reference's second iteration, A.7) keep their own dedupe scope: they are
// never dropped and never serve as a predecessor.
//
// The pass runs only for calls whose survivor count exceeds SEGALIGN_B200_MERGE_MIN (default 65536):
// stage B sees the count on the device and returns at once, the host sorts (cub radix sort, count now
// known), marks, compacts, and replays stage B on the representatives.
#pragma once
#include "kernels_extend.cuh"

namespace sa {

constexpr uint32_t MERGE_MAX_GAP = 4096; // cells between neighbouring anchors that are still checked for an all-match stretch

// sort keys: diagonal (uint32 wrap-around, as the reference's hspComp computes it) << 32 | reference anchor
__global__ void __launch_bounds__(256)
k_merge_keys(const SurvRec *__restrict__ surv, uint32_t n, unsigned long long *__restrict__ keys, uint32_t *__restrict__ idx) {
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const SurvRec r = surv[i];
        keys[i] = ((unsigned long long)(r.r0 - r.q0) << 32) | r.r0;
        idx[i] = i;
    }
}

// cells [c0, c0 + n) of the reference against [c0 - d, ...) of the query: all upper-case ACGT and equal
__device__ __forceinline__ bool all_match_stretch(const ExtendParams &P, uint32_t rc, uint32_t qc, uint32_t n) {
    for (uint32_t done = 0; done < n; done += 32u) {
        const uint32_t left = n - done;
        const uint32_t keep = left >= 32u ? 0xFFFFFFFFu : ((1u << left) - 1u);
        const uint32_t m = (load_m1_window(P.rm1, rc + done) | load_m1_window(P.qm1, qc + done)) & keep;
        if (m) return false;
        const uint64_t x = load_p2_window(P.rp2, rc + done) ^ load_p2_window(P.qp2, qc + done);
        const uint64_t keep2 = left >= 32u ? ~0ull : ((1ull << (2 * left)) - 1ull);
        if (x & keep2) return false;
    }
    return true;
}

// sorted position i keeps its survivor unless it is connected to the survivor at i-1 (see above)
__global__ void __launch_bounds__(256)
k_merge_mark(ExtendParams P, const SurvRec *__restrict__ surv, const uint32_t *__restrict__ idx, uint32_t n,
             SurvRec *__restrict__ out, uint32_t *__restrict__ counters) {
    const uint32_t stride = gridDim.x * blockDim.x;
    const uint32_t lastkey = counters[CTR_LASTKEY];
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const SurvRec cur = surv[idx[i]];
        bool keep = true;
        if (i > 0 && cur.key < lastkey) {
            const SurvRec prev = surv[idx[i - 1]];
            if (prev.key < lastkey && prev.r0 - prev.q0 == cur.r0 - cur.q0 && cur.r0 >= prev.r0 && cur.q0 >= prev.q0 &&
                cur.r0 - prev.r0 <= MERGE_MAX_GAP && cur.r0 <= P.ref_len && cur.q0 <= P.query_len)
                keep = !all_match_stretch(P, prev.r0, prev.q0, cur.r0 - prev.r0);
        }
        if (keep) out[atomicAdd(counters + CTR_MERGED, 1u)] = cur; // at most n records
    }
}

} // namespace sa
