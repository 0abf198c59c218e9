// kernels_sort.cuh -- diagonal sort, adjacent-containment dedupe, final LASTZ order.
//
// Replaces thrust::stable_sort(hspComp) / unique_copy(hspEqual) / stable_sort(hspCompLastz)
// (src/seed_filter.cu:776-782).  Every anchor carries the reference iteration it belongs to
// as the most significant key, so one pass handles all iterations of a call while dedupe
// still never crosses an iteration boundary.
#pragma once
#include "sa_common.cuh"

namespace sa {

struct CompDiag { // hspComp, seed_filter.cu:54-80, behind the iteration tag
    __device__ __forceinline__ bool operator()(const Anchor &x, const Anchor &y) const {
        if (x.tag != y.tag) return x.tag < y.tag;
        uint32_t dx = x.ref_start - x.query_start, dy = y.ref_start - y.query_start;
        if (dx != dy) return dx < dy;
        if (x.ref_start != y.ref_start) return x.ref_start < y.ref_start;
        if (x.len != y.len) return x.len < y.len;
        return x.score > y.score;
    }
};

struct CompLastz { // hspCompLastz, seed_filter.cu:82-108, behind the iteration tag
    __device__ __forceinline__ bool operator()(const Anchor &x, const Anchor &y) const {
        if (x.tag != y.tag) return x.tag < y.tag;
        if (x.query_start != y.query_start) return x.query_start < y.query_start;
        if (x.ref_start != y.ref_start) return x.ref_start < y.ref_start;
        if (x.len != y.len) return x.len < y.len;
        return x.score > y.score;
    }
};

__device__ __forceinline__ bool hsp_equal(const Anchor &x, const Anchor &y) {
    // hspEqual, seed_filter.cu:47-52 (u32 wrap-around arithmetic)
    return ((x.ref_start - x.query_start) == (y.ref_start - y.query_start)) &&
           (((x.ref_start >= y.ref_start) && ((x.ref_start + x.len) <= (y.ref_start + y.len))) ||
            ((y.ref_start >= x.ref_start) && ((y.ref_start + y.len) <= (x.ref_start + x.len))));
}

// unique_copy keeps element i iff it is not "equal" to element i-1 OF THE SORTED INPUT
// (thrust head flags).  Survivors are appended unordered; the final sort is a total order.
__global__ void __launch_bounds__(256)
k_dedupe(const Anchor *__restrict__ in, uint32_t n, Anchor *__restrict__ out,
         uint32_t *__restrict__ out_count) {
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        Anchor cur = in[i];
        bool keep = true;
        if (i > 0) {
            Anchor prev = in[i - 1];
            keep = (prev.tag != cur.tag) || !hsp_equal(prev, cur);
        }
        if (keep) out[atomicAdd(out_count, 1u)] = cur;
    }
}

__global__ void __launch_bounds__(256)
k_strip_tags(const Anchor *__restrict__ in, uint32_t n, sa_segment *__restrict__ out) {
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        Anchor a = in[i];
        sa_segment s;
        s.ref_start = a.ref_start;
        s.query_start = a.query_start;
        s.len = a.len;
        s.score = a.score;
        out[i] = s;
    }
}

} // namespace sa
