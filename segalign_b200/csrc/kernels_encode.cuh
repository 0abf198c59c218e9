// kernels_encode.cuh -- ASCII -> {b8, p2, m1} encoders, seed-word generation, seed table keys.
//
// Replaces compress_string (common/seed_filter_interface.cu:18-47) and
// compress_string_rev_comp (src/seed_filter.cu:110-155); the k-mer extraction restates
// GetKmerIndexAtPos (common/ntcoding.cpp:43-61) on the packed planes.
// All three are HBM-streaming kernels: 16-byte vector loads/stores, grid-stride.
#pragma once
#include "sa_common.cuh"

namespace sa {

// ASCII -> b8 (forward) and, if dst_rc != nullptr, the reverse complement b8 plane.
// Each thread converts 16 consecutive bases.
__global__ void __launch_bounds__(256)
k_encode_b8(const uint8_t *__restrict__ src, uint32_t len, uint8_t *__restrict__ dst,
            uint8_t *__restrict__ dst_rc) {
    const uint32_t nvec = len >> 4;
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t v = blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += stride) {
        uint4 in = __ldg(reinterpret_cast<const uint4 *>(src) + v);
        uint32_t w[4] = {in.x, in.y, in.z, in.w};
        uint32_t o[4], orc[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            uint32_t ow = 0, rw = 0;
#pragma unroll
            for (int b = 0; b < 4; b++) {
                uint8_t c = encode_ascii((w[k] >> (8 * b)) & 0xFF);
                ow |= (uint32_t)c << (8 * b);
                uint8_t rc = c < 4 ? (uint8_t)(3 - c) : c; // A<->T, C<->G, others unchanged
                rw |= (uint32_t)rc << (8 * (3 - b));
            }
            o[k] = ow;
            orc[3 - k] = rw;
        }
        reinterpret_cast<uint4 *>(dst)[v] = make_uint4(o[0], o[1], o[2], o[3]);
        if (dst_rc) {
            // bases [16v, 16v+16) land reversed at [len-16v-16, len-16v): generally unaligned
            uint8_t *p = dst_rc + (len - 16u * v - 16u);
            if ((reinterpret_cast<uintptr_t>(p) & 15u) == 0) {
                *reinterpret_cast<uint4 *>(p) = make_uint4(orc[0], orc[1], orc[2], orc[3]);
            } else {
#pragma unroll
                for (int k = 0; k < 4; k++)
#pragma unroll
                    for (int b = 0; b < 4; b++) p[4 * k + b] = (orc[k] >> (8 * b)) & 0xFF;
            }
        }
    }
    // tail (< 16 bases)
    uint32_t t = (nvec << 4) + blockIdx.x * blockDim.x + threadIdx.x;
    if (t < len) {
        uint8_t c = encode_ascii(src[t]);
        dst[t] = c;
        if (dst_rc) dst_rc[len - 1 - t] = c < 4 ? (uint8_t)(3 - c) : c;
    }
}

// b8 -> p2 + m1; one thread per 32-base word, padding words are fully masked.
__global__ void __launch_bounds__(256)
k_pack_planes(const uint8_t *__restrict__ b8, uint32_t len, uint64_t *__restrict__ p2,
              uint32_t *__restrict__ m1, uint32_t words) {
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t w = blockIdx.x * blockDim.x + threadIdx.x; w < words; w += stride) {
        uint64_t bits = 0;
        uint32_t mask = 0;
        const uint32_t base = w << 5;
        if (base + 32 <= len) {
            const uint4 *p = reinterpret_cast<const uint4 *>(b8 + base);
            uint4 a = __ldg(p), b = __ldg(p + 1);
            uint32_t x[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
            for (int k = 0; k < 8; k++) {
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    uint32_t c = (x[k] >> (8 * j)) & 0xFF;
                    int cell = 4 * k + j;
                    if (c < 4) bits |= (uint64_t)c << (2 * cell);
                    else mask |= 1u << cell;
                }
            }
        } else {
            for (int cell = 0; cell < 32; cell++) {
                uint32_t i = base + cell;
                if (base < len && i < len) {
                    uint32_t c = b8[i];
                    if (c < 4) bits |= (uint64_t)c << (2 * cell);
                    else mask |= 1u << cell;
                } else {
                    mask |= 1u << cell;
                }
            }
        }
        p2[w] = bits;
        m1[w] = mask;
    }
}

// Spaced-seed word of the window whose cell 0 is the seed start (ntcoding.cpp:54-58).
__device__ __forceinline__ uint32_t kmer_from_window(uint64_t win, const ShapeDesc &sh) {
    uint32_t k = 0;
    for (int i = 0; i < sh.weight; i++) k = (k << 2) | (uint32_t)((win >> (2 * sh.pos[i])) & 3u);
    return k;
}
// Valid iff all `span` cells (don't-care positions included) are upper-case ACGT
// (ntcoding.cpp:47-52).
__device__ __forceinline__ bool seed_valid(uint32_t mwin, int span) {
    uint32_t m = span >= 32 ? 0xFFFFFFFFu : ((1u << span) - 1u);
    return (mwin & m) == 0;
}

// Seed table pass 1 (seed_pos_table.cu:69-81): key/position pairs + bucket histogram.
// Invalid positions get key 4^w so a radix sort moves them behind every real bucket.
// b8 -> the zero-run planes of stage B (screen_bound.h: zero_run_codes): f1 / g1 = 1 bit per base, set where the cell's
// code is flat / a partner; padding words and cells past the end are 0.  counter += number of flat cells.
__global__ void __launch_bounds__(256)
k_pack_zero_planes(const uint8_t *__restrict__ b8, uint32_t len, uint32_t *__restrict__ f1, uint32_t *__restrict__ g1,
                   uint32_t words, uint32_t flat, uint32_t partners, uint32_t *__restrict__ counter) {
    const uint32_t stride = gridDim.x * blockDim.x;
    uint32_t nflat = 0;
    for (uint32_t w = blockIdx.x * blockDim.x + threadIdx.x; w < words; w += stride) {
        uint32_t f = 0, g = 0;
        const unsigned long long base = (unsigned long long)w << 5;
        if (base + 32 <= len) {
            const uint4 *p = reinterpret_cast<const uint4 *>(b8 + base);
            const uint4 a = __ldg(p), b = __ldg(p + 1);
            const uint32_t x[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
            for (int k = 0; k < 8; k++) {
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const uint32_t c = (x[k] >> (8 * j)) & 7u;
                    f |= ((flat >> c) & 1u) << (4 * k + j);
                    g |= ((partners >> c) & 1u) << (4 * k + j);
                }
            }
        } else {
            for (int cell = 0; cell < 32; cell++) {
                if (base + cell < len) {
                    const uint32_t c = b8[base + cell] & 7u;
                    f |= ((flat >> c) & 1u) << cell;
                    g |= ((partners >> c) & 1u) << cell;
                }
            }
        }
        f1[w] = f;
        g1[w] = g;
        nflat += __popc(f);
    }
    if (nflat) atomicAdd(counter, nflat);
}

// Coarse level: bit p of F1k / G1k = all 1024 bases of the aligned piece p exist and are flat / partners.  One thread
// per piece; the 32 pieces of an output word are combined by a ballot.
__global__ void __launch_bounds__(256)
k_coarse_zero_planes(const uint32_t *__restrict__ f1, const uint32_t *__restrict__ g1, uint32_t words,
                     uint32_t *__restrict__ F1k, uint32_t *__restrict__ G1k, uint32_t coarse_words) {
    const uint32_t pieces = coarse_words * 32u; // a multiple of the warp size: whole warps run the loop together
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t pc = blockIdx.x * blockDim.x + threadIdx.x; pc < pieces; pc += stride) {
        const unsigned long long first = (unsigned long long)pc * 32u;
        uint32_t af = 0, ag = 0;
        if (first + 32u <= words) {
            af = ag = 0xFFFFFFFFu;
            const uint4 *pf = reinterpret_cast<const uint4 *>(f1 + first), *pg = reinterpret_cast<const uint4 *>(g1 + first);
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const uint4 a = __ldg(pf + k), b = __ldg(pg + k);
                af &= a.x & a.y & a.z & a.w;
                ag &= b.x & b.y & b.z & b.w;
            }
        }
        const uint32_t mf = __ballot_sync(0xFFFFFFFFu, af == 0xFFFFFFFFu), mg = __ballot_sync(0xFFFFFFFFu, ag == 0xFFFFFFFFu);
        if ((threadIdx.x & 31u) == 0) { F1k[pc >> 5] = mf; G1k[pc >> 5] = mg; }
    }
}

__global__ void __launch_bounds__(256)
k_table_keys(const uint64_t *__restrict__ p2, const uint32_t *__restrict__ m1, ShapeDesc sh,
             uint32_t start_offset, uint32_t step, uint32_t num_steps, uint32_t *__restrict__ keys,
             uint32_t *__restrict__ vals, uint32_t *__restrict__ hist) {
    const uint32_t stride = gridDim.x * blockDim.x;
    const uint32_t invalid_key = 1u << (2 * sh.weight);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < num_steps; i += stride) {
        uint32_t p = start_offset + i * step;
        uint32_t mw = load_m1_window(m1, p);
        uint32_t key = invalid_key;
        if (seed_valid(mw, sh.span)) {
            key = kmer_from_window(load_p2_window(p2, p), sh);
            atomicAdd(hist + key, 1u);
        }
        keys[i] = key;
        vals[i] = p;
    }
}

// Device-side seeding (src/seeder.cpp:57-74): flag valid seed starts in [j0, j1).
__global__ void __launch_bounds__(256)
k_seed_flags(const uint32_t *__restrict__ m1, int span, uint32_t j0, uint32_t j1,
             uint32_t *__restrict__ flags) {
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t j = j0 + blockIdx.x * blockDim.x + threadIdx.x; j < j1; j += stride)
        flags[j - j0] = seed_valid(load_m1_window(m1, j), span) ? 1u : 0u;
}

// Emits the seed words in the reference's order: position ascending, exact word first, then
// the transition variants t = 0..w-1 (kmer ^ (2 << 2t), seeder.cpp:64-71).
// excl = exclusive scan of flags; words_per_pos = 1 + (#transition positions if enabled).
__global__ void __launch_bounds__(256)
k_seed_emit(const uint64_t *__restrict__ p2, const uint32_t *__restrict__ flags,
            const uint32_t *__restrict__ excl, ShapeDesc sh, int transition, uint32_t j0,
            uint32_t j1, uint64_t *__restrict__ seeds, uint32_t *__restrict__ d_num_seeds) {
    const uint32_t stride = gridDim.x * blockDim.x;
    const uint32_t per = 1u + (transition ? (uint32_t)sh.num_trans : 0u);
    for (uint32_t j = j0 + blockIdx.x * blockDim.x + threadIdx.x; j < j1; j += stride) {
        if (j == j1 - 1) *d_num_seeds = (excl[j - j0] + flags[j - j0]) * per; // total seed words
        if (!flags[j - j0]) continue;
        uint64_t kmer = kmer_from_window(load_p2_window(p2, j), sh);
        uint64_t *o = seeds + (size_t)excl[j - j0] * per;
        *o++ = (kmer << 32) + j;
        if (transition) {
            for (int t = 0; t < sh.weight; t++)
                if (sh.trans[t]) *o++ = ((kmer ^ ((uint64_t)2 << (2 * t))) << 32) + j;
        }
    }
}

} // namespace sa
