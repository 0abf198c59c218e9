// kernels_lookup.cuh -- seed position table lookup: bucket sizes, iteration plan, hit expansion.
//
// Replaces find_num_hits (src/seed_filter.cu:157-182), the lower_bound loop of SeedAndFilter
// (:718-745) and find_hits (:184-230).  The reference expands one 128-thread block per seed
// with 4 active lanes; here a warp owns 32 consecutive seeds and expands their buckets as one
// flat, load-balanced range so pos_table reads and hit writes are coalesced.
#pragma once
#include "sa_common.cuh"

namespace sa {

// bucket size per seed word: n = T[k] - (k ? T[k-1] : 0)   (seed_filter.cu:172-180)
// The seed count lives in device memory (*d_num_seeds): with device-side seeding the host only
// knows an upper bound (max_items) when it enqueues the call.  Slots past the count get 0 hits.
// A seed word whose k-mer field lies outside the table or whose span runs past the query block (only a
// caller-supplied vector can hold one; the reference's seeder never emits it) has an empty bucket.
struct SeedBounds {
    uint32_t index_size, query_len, seed_size;
    __device__ __forceinline__ bool ok(uint64_t word) const {
        return (uint32_t)(word >> 32) < index_size && (unsigned long long)(uint32_t)word + seed_size <= query_len;
    }
};
__global__ void __launch_bounds__(256)
k_count_hits(const uint64_t *__restrict__ seeds, uint32_t max_items, const uint32_t *__restrict__ d_num_seeds,
             const uint32_t *__restrict__ index_table, SeedBounds B, uint32_t *__restrict__ counts,
             unsigned long long *__restrict__ total64) {
    const uint32_t num_seeds = min(*d_num_seeds, max_items);
    const uint32_t stride = gridDim.x * blockDim.x;
    unsigned long long mine = 0;
    for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < max_items; s += stride) {
        uint32_t n = 0;
        if (s < num_seeds && B.ok(seeds[s])) {
            uint32_t kmer = (uint32_t)(seeds[s] >> 32);
            n = __ldg(index_table + kmer);
            if (kmer > 0) n -= __ldg(index_table + kmer - 1);
        }
        counts[s] = n;
        mine += n;
    }
    // 64-bit total beside the uint32 scan (the repeat masker reports it, repeat_masker_src/seed_filter.cu:856-857)
    for (int off = 16; off > 0; off >>= 1) mine += __shfl_down_sync(0xFFFFFFFFu, mine, off);
    if ((threadIdx.x & 31) == 0 && mine) atomicAdd(total64, mine);
}

// Seed words from their base words.  The reference's seeder emits, per query position, the exact
// word followed by its transition variants (src/seeder.cpp:57-74): word[g*per + v] = base[g] ^ xm[v].
// When the host finds a seed vector in that form it uploads only the base words (1/per of the
// bytes) and this kernel rebuilds the vector in HBM, bit for bit.
struct VariantMasks {
    uint64_t xm[32]; // xm[0] = 0; xm[v] = (TRANSITION_MASK << 2*t_v) << 32
};
__global__ void __launch_bounds__(256)
k_expand_bases(const uint64_t *__restrict__ bases, uint32_t num_seeds, uint32_t per, VariantMasks M,
               uint64_t *__restrict__ seeds) {
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < num_seeds; s += stride) {
        const uint32_t g = s / per, v = s - g * per;
        seeds[s] = __ldg(bases + g) ^ M.xm[v];
    }
}

__device__ __forceinline__ uint32_t lower_bound_dev(const uint32_t *a, uint32_t n, uint32_t v) {
    uint32_t lo = 0, hi = n;
    while (lo < hi) {
        uint32_t mid = lo + ((hi - lo) >> 1);
        if (a[mid] < v) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// Iteration plan (seed_filter.cu:718-745) evaluated on the device so the prefix array never
// leaves HBM.  plan[0] = num_iter, plan[1] = num_hits; limit_pos[i] = last seed of iteration i;
// hit_bound[i] = flat hit index one past iteration i.
// Reference UB zone (SURVEY A.11 i): lower_bound == 0 -> pos wraps; defined here as an empty
// iteration (limit_pos = 0xFFFFFFFF, bound 0), same as oracle/sa_oracle.c.
// plan[2] = number of seed words (input, written by the seeding step).
__global__ void k_plan_iterations(const uint32_t *__restrict__ prefix, uint32_t max_items,
                                  uint32_t max_hits, uint32_t cap, uint32_t *__restrict__ limit_pos,
                                  uint32_t *__restrict__ hit_bound, uint32_t *__restrict__ plan) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    const uint32_t num_seeds = min(plan[2], max_items);
    if (num_seeds == 0) { plan[0] = 0; plan[1] = 0; return; }
    uint32_t num_hits = prefix[num_seeds - 1];
    plan[1] = num_hits;
    if (num_hits == 0) { plan[0] = 0; return; }
    uint32_t num_iter, limit;
    if (num_hits < max_hits) { num_iter = 2; limit = num_hits; }
    else { num_iter = num_hits / max_hits + 2; limit = max_hits; }
    if (num_iter > cap) { plan[0] = 0xFFFFFFFFu; return; } // host sized the arrays from num_hits
    for (uint32_t i = 0; i + 1 < num_iter; i++) {
        uint32_t pos = lower_bound_dev(prefix, num_seeds, limit) - 1u;
        limit_pos[i] = pos;
        uint32_t base = (pos == 0xFFFFFFFFu) ? 0u : prefix[pos];
        hit_bound[i] = base;
        limit = base + max_hits;
        if (limit > num_hits) limit = num_hits;
    }
    limit_pos[num_iter - 1] = num_seeds - 1;
    hit_bound[num_iter - 1] = num_hits;
    if (limit_pos[num_iter - 1] == limit_pos[num_iter - 2]) num_iter--;
    plan[0] = num_iter;
}

// Flat hit expansion.  hits[h] = (ref anchor, query anchor) = (pos + seed_size, qpos + seed_size)
// (seed_filter.cu:204,220); flat order is seed-major, i.e. the hits of seed s occupy
// [prefix[s] - n_s, prefix[s]).  The order inside a bucket is irrelevant.
__global__ void __launch_bounds__(256)
k_expand_hits(const uint64_t *__restrict__ seeds, uint32_t max_items, const uint32_t *__restrict__ d_num_seeds,
              const uint32_t *__restrict__ index_table, const uint32_t *__restrict__ pos_table,
              const uint32_t *__restrict__ prefix, uint32_t seed_size, SeedBounds B, uint2 *__restrict__ hits,
              uint32_t hits_cap) {
    const uint32_t num_seeds = min(*d_num_seeds, max_items);
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
    const uint32_t ngroups = (num_seeds + 31u) >> 5;
    for (uint32_t g = warp; g < ngroups; g += nwarps) {
        uint32_t s = (g << 5) + lane;
        uint32_t start = 0, n = 0, q = 0, incl_global = 0;
        if (s < num_seeds) {
            uint64_t word = seeds[s];
            uint32_t kmer = (uint32_t)(word >> 32);
            if (B.ok(word)) {
                uint32_t end = __ldg(index_table + kmer);
                start = kmer > 0 ? __ldg(index_table + kmer - 1) : 0u;
                n = end - start;
            }
            q = (uint32_t)word + seed_size;
            incl_global = prefix[s];
        }
        // warp inclusive scan of n
        uint32_t incl = n;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, off);
            if (lane >= (uint32_t)off) incl += t;
        }
        const uint32_t total = __shfl_sync(0xFFFFFFFFu, incl, 31);
        const uint32_t excl = incl - n;
        const uint32_t gbase = __shfl_sync(0xFFFFFFFFu, incl_global - n, 0);
        for (uint32_t f0 = 0; f0 < total; f0 += 32) {
            uint32_t f = f0 + lane;
            // owner = first lane whose inclusive prefix exceeds f
            uint32_t lo = 0, hi = 31;
#pragma unroll
            for (int it = 0; it < 5; it++) {
                uint32_t mid = (lo + hi) >> 1;
                uint32_t v = __shfl_sync(0xFFFFFFFFu, incl, mid);
                if (v > f) hi = mid; else lo = mid + 1;
            }
            uint32_t o_excl = __shfl_sync(0xFFFFFFFFu, excl, lo);
            uint32_t o_start = __shfl_sync(0xFFFFFFFFu, start, lo);
            uint32_t o_q = __shfl_sync(0xFFFFFFFFu, q, lo);
            // hits past the buffer capacity are dropped here; the host sees num_hits > capacity at
            // its single synchronisation point, grows the buffers and replays the call
            if (f < total && gbase + f < hits_cap) {
                uint32_t r = __ldg(pos_table + o_start + (f - o_excl)) + seed_size;
                hits[(size_t)gbase + f] = make_uint2(r, o_q);
            }
        }
    }
}

} // namespace sa
