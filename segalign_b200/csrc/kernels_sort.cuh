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

namespace sa {

// ---------------------------------------------------------------------------------------------
// One-launch finalisation for the common case of few anchors per call (a 250 kb chunk yields
// ~10^2 HSPs): diagonal sort -> predecessor dedupe -> final order -> strip tags, all inside one
// thread block on shared memory, with the element count read from device memory so the host
// does not have to synchronise between the extension and the sort.  Same semantics as the
// cub path above (src/seed_filter.cu:776-782); falls back to it by reporting
// counters[CTR_OUT] = 0xFFFFFFFF when there are more than FINALIZE_CAP anchors.
constexpr int FINALIZE_CAP = 1024; // 2 x 1024 x 20 B of static shared memory
constexpr int FINALIZE_THREADS = 1024;

template <typename Comp>
__device__ __forceinline__ void block_bitonic_sort(Anchor *a, int n_pow2, Comp less) {
    for (int k = 2; k <= n_pow2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < n_pow2; i += blockDim.x) {
                int ixj = i ^ j;
                if (ixj > i) {
                    Anchor x = a[i], y = a[ixj];
                    bool up = (i & k) == 0;
                    if (up ? less(y, x) : less(x, y)) { a[i] = y; a[ixj] = x; }
                }
            }
            __syncthreads();
        }
    }
}

__global__ void __launch_bounds__(FINALIZE_THREADS)
k_finalize_small(const Anchor *__restrict__ anchors, uint32_t anchor_cap, sa_segment *__restrict__ out,
                 uint32_t *__restrict__ counters, uint32_t small_cap) {
    __shared__ Anchor a[FINALIZE_CAP];
    __shared__ Anchor b[FINALIZE_CAP];
    __shared__ uint32_t kept;
    const uint32_t n = counters[0]; // CTR_ANCHORS
    if (n > small_cap || n > anchor_cap) { // too many for one block (or the append overflowed): host takes the cub path
        if (threadIdx.x == 0) counters[6] = 0xFFFFFFFFu;
        return;
    }
    if (n == 0) {
        if (threadIdx.x == 0) counters[6] = 0;
        return;
    }
    int n2 = 1;
    while (n2 < (int)n) n2 <<= 1;
    Anchor pad; // sorts after every real anchor under both orders
    pad.tag = 0xFFFFFFFFu; pad.ref_start = 0xFFFFFFFFu; pad.query_start = 0; pad.len = 0xFFFFFFFFu; pad.score = 0;
    for (int i = threadIdx.x; i < n2; i += blockDim.x) a[i] = i < (int)n ? anchors[i] : pad;
    if (threadIdx.x == 0) kept = 0;
    __syncthreads();
    block_bitonic_sort(a, n2, CompDiag());
    // unique_copy head flags against the predecessor of the SORTED INPUT; survivors go to b[] in
    // any order (the final order is total)
    for (int i = threadIdx.x; i < (int)n; i += blockDim.x) {
        bool keep = i == 0 || a[i - 1].tag != a[i].tag || !hsp_equal(a[i - 1], a[i]);
        if (keep) b[atomicAdd(&kept, 1u)] = a[i];
    }
    __syncthreads();
    const uint32_t m = kept;
    int m2 = 1;
    while (m2 < (int)m) m2 <<= 1;
    for (int i = (int)m + threadIdx.x; i < m2; i += blockDim.x) b[i] = pad;
    __syncthreads();
    block_bitonic_sort(b, m2, CompLastz());
    for (int i = threadIdx.x; i < (int)m; i += blockDim.x) {
        sa_segment s;
        s.ref_start = b[i].ref_start; s.query_start = b[i].query_start; s.len = b[i].len; s.score = b[i].score;
        out[i] = s;
    }
    if (threadIdx.x == 0) counters[6] = m;
}

} // namespace sa

// ---------------------------------------------------------------------------------------------
// Repeat-masker variant (repeat_masker_src/seed_filter.cu:819-835): three stable sorts
// and two unique passes.  Each of the reference's comparators is extended here to a TOTAL order by
// appending the order the records had before that stable sort, which is itself a function of the
// records (never of the hit order):
//   :819 hspComp      (query_start, len desc, ref_start, score desc)            -- already total
//   :821 unique_copy(hspEqual)      exact copies (adjacent under a total order)
//   :825 hspDiagComp  (diagonal, ref_start, query_start, score desc) + ties keep the hspComp order:
//                     equal q, r, score => they differ in len only => len desc
//   :827 unique_copy(hspDiagEqual)  same diagonal + containment, against the predecessor of the input
//   :833 hspFinalComp (query_start, score desc, ref_start desc) + ties keep the order before it: equal
//                     q, r, score => one diagonal, they differ in len only => len desc
namespace sa {

struct CompRmFirst {
    __device__ __forceinline__ bool operator()(const Anchor &x, const Anchor &y) const {
        if (x.tag != y.tag) return x.tag < y.tag;
        if (x.query_start != y.query_start) return x.query_start < y.query_start;
        if (x.len != y.len) return x.len > y.len;
        if (x.ref_start != y.ref_start) return x.ref_start < y.ref_start;
        return x.score > y.score;
    }
};
struct CompRmDiag {
    __device__ __forceinline__ bool operator()(const Anchor &x, const Anchor &y) const {
        if (x.tag != y.tag) return x.tag < y.tag;
        const uint32_t dx = x.ref_start - x.query_start, dy = y.ref_start - y.query_start;
        if (dx != dy) return dx < dy;
        if (x.ref_start != y.ref_start) return x.ref_start < y.ref_start;
        if (x.query_start != y.query_start) return x.query_start < y.query_start;
        if (x.score != y.score) return x.score > y.score;
        return x.len > y.len;
    }
};
struct CompRmFinal {
    __device__ __forceinline__ bool operator()(const Anchor &x, const Anchor &y) const {
        if (x.tag != y.tag) return x.tag < y.tag;
        if (x.query_start != y.query_start) return x.query_start < y.query_start;
        if (x.score != y.score) return x.score > y.score;
        if (x.ref_start != y.ref_start) return x.ref_start > y.ref_start;
        return x.len > y.len;
    }
};
__device__ __forceinline__ bool rm_exact_equal(const Anchor &x, const Anchor &y) { // hspEqual, :79-84
    return x.ref_start == y.ref_start && x.query_start == y.query_start && x.len == y.len && x.score == y.score;
}
// minus-strand records back to forward coordinates (compress_output, :705-709)
__device__ __forceinline__ void rm_to_forward(Anchor &a, uint32_t block_len) {
    a.query_start = block_len - 1u - (a.query_start + a.len);
}

__global__ void __launch_bounds__(256)
k_rm_to_forward(Anchor *__restrict__ a, uint32_t n, uint32_t block_len) {
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) rm_to_forward(a[i], block_len);
}

// unique_copy(hspEqual) on the sorted input; survivors unordered (the next sort is total)
__global__ void __launch_bounds__(256)
k_dedupe_exact(const Anchor *__restrict__ in, uint32_t n, Anchor *__restrict__ out, uint32_t *__restrict__ out_count) {
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const Anchor cur = in[i];
        bool keep = true;
        if (i > 0) {
            const Anchor prev = in[i - 1];
            keep = prev.tag != cur.tag || !rm_exact_equal(prev, cur);
        }
        if (keep) out[atomicAdd(out_count, 1u)] = cur;
    }
}

// one-block version of the whole chain for <= FINALIZE_CAP anchors (same contract as k_finalize_small)
__global__ void __launch_bounds__(FINALIZE_THREADS)
k_finalize_small_rm(const Anchor *__restrict__ anchors, uint32_t anchor_cap, sa_segment *__restrict__ out,
                    uint32_t *__restrict__ counters, uint32_t small_cap, int rev, uint32_t block_len) {
    __shared__ Anchor a[FINALIZE_CAP];
    __shared__ Anchor b[FINALIZE_CAP];
    __shared__ uint32_t kept, kept2;
    const uint32_t n = counters[0]; // CTR_ANCHORS
    if (n > small_cap || n > anchor_cap) {
        if (threadIdx.x == 0) counters[6] = 0xFFFFFFFFu;
        return;
    }
    if (n == 0) {
        if (threadIdx.x == 0) counters[6] = 0;
        return;
    }
    Anchor pad; // sorts after every real anchor under all three orders
    pad.tag = 0xFFFFFFFFu; pad.ref_start = 0xFFFFFFFFu; pad.query_start = 0xFFFFFFFFu; pad.len = 0; pad.score = 0;
    auto pow2 = [](uint32_t v) { int p = 1; while (p < (int)v) p <<= 1; return p; };
    const int n2 = pow2(n);
    for (int i = threadIdx.x; i < n2; i += blockDim.x) {
        Anchor x = pad;
        if (i < (int)n) { x = anchors[i]; if (rev) rm_to_forward(x, block_len); }
        a[i] = x;
    }
    if (threadIdx.x == 0) { kept = 0; kept2 = 0; }
    __syncthreads();
    block_bitonic_sort(a, n2, CompRmFirst());
    for (int i = threadIdx.x; i < (int)n; i += blockDim.x) {
        const bool keep = i == 0 || a[i - 1].tag != a[i].tag || !rm_exact_equal(a[i - 1], a[i]);
        if (keep) b[atomicAdd(&kept, 1u)] = a[i];
    }
    __syncthreads();
    const uint32_t m = kept;
    const int m2 = pow2(m);
    for (int i = (int)m + threadIdx.x; i < m2; i += blockDim.x) b[i] = pad;
    __syncthreads();
    block_bitonic_sort(b, m2, CompRmDiag());
    for (int i = threadIdx.x; i < (int)m; i += blockDim.x) {
        const bool keep = i == 0 || b[i - 1].tag != b[i].tag || !hsp_equal(b[i - 1], b[i]);
        if (keep) a[atomicAdd(&kept2, 1u)] = b[i];
    }
    __syncthreads();
    const uint32_t k = kept2;
    const int k2 = pow2(k);
    for (int i = (int)k + threadIdx.x; i < k2; i += blockDim.x) a[i] = pad;
    __syncthreads();
    block_bitonic_sort(a, k2, CompRmFinal());
    for (int i = threadIdx.x; i < (int)k; i += blockDim.x) {
        sa_segment s;
        s.ref_start = a[i].ref_start; s.query_start = a[i].query_start; s.len = a[i].len; s.score = a[i].score;
        out[i] = s;
    }
    if (threadIdx.x == 0) counters[6] = k;
}

} // namespace sa


// ---------------------------------------------------------------------------------------------
// Device-wide path (more than FINALIZE_CAP anchors): stable LSD radix passes (cub::DeviceRadixSort)
// over the composite key of each order, 64 bits per pass, least significant word first, carrying a
// permutation.  Every order above is a lexicographic order of 32-bit fields (descending fields are
// stored inverted), so three 64-bit words hold it:
//   ORD_DIAG      hspComp       [tag, diagonal] [ref_start, len]      [~score]
//   ORD_LASTZ     hspCompLastz  [tag, query]    [ref_start, len]      [~score]
//   ORD_RM_FIRST  :819          [tag, query]    [~len, ref_start]     [~score]
//   ORD_RM_DIAG   :825          [tag, diagonal] [ref_start, query]    [~score, ~len]
//   ORD_RM_FINAL  :833          [tag, query]    [~score, ~ref_start]  [~len]
namespace sa {

enum { ORD_DIAG = 0, ORD_LASTZ = 1, ORD_RM_FIRST = 2, ORD_RM_DIAG = 3, ORD_RM_FINAL = 4 };

__device__ __forceinline__ unsigned long long anchor_key_word(const Anchor &a, int order, int word) {
    const uint32_t diag = a.ref_start - a.query_start;
    const uint32_t nscore = ~((uint32_t)a.score ^ 0x80000000u); // descending signed score as ascending unsigned
    uint32_t hi = 0, lo = 0;
    if (word == 0) { hi = a.tag; lo = (order == ORD_DIAG || order == ORD_RM_DIAG) ? diag : a.query_start; }
    else if (word == 1) {
        if (order == ORD_DIAG || order == ORD_LASTZ) { hi = a.ref_start; lo = a.len; }
        else if (order == ORD_RM_FIRST) { hi = ~a.len; lo = a.ref_start; }
        else if (order == ORD_RM_DIAG) { hi = a.ref_start; lo = a.query_start; }
        else { hi = nscore; lo = ~a.ref_start; }
    } else {
        if (order == ORD_RM_DIAG) { hi = nscore; lo = ~a.len; }
        else if (order == ORD_RM_FINAL) { hi = ~a.len; lo = 0; }
        else { hi = nscore; lo = 0; }
    }
    return ((unsigned long long)hi << 32) | lo;
}

// keys of one pass for the anchors in their current permutation (perm == nullptr: identity, and writes it)
__global__ void __launch_bounds__(256)
k_anchor_keys(const Anchor *__restrict__ a, const uint32_t *__restrict__ perm_in, uint32_t n, int order, int word,
              unsigned long long *__restrict__ keys, uint32_t *__restrict__ perm_out) {
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint32_t src = perm_in ? perm_in[i] : i;
        keys[i] = anchor_key_word(a[src], order, word);
        if (!perm_in) perm_out[i] = i;
    }
}

__global__ void __launch_bounds__(256)
k_anchor_gather(const Anchor *__restrict__ a, const uint32_t *__restrict__ perm, uint32_t n, Anchor *__restrict__ out) {
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = a[perm[i]];
}

} // namespace sa
