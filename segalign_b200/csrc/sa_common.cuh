// sa_common.cuh -- shared types, constants and device helpers of the B200 backend.
//
// Sequence layout in HBM (per encoded block, see DESIGN.md "Data layout"):
//   b8 : 1 byte / base, the reference's 8-symbol code (common/parameters.h:5-13)
//   p2 : 2 bits / base, 32 bases per little-endian uint64 word (base i -> bits 2*(i&31));
//        A,C,G,T = 0..3, every other symbol stored as 0
//   m1 : 1 bit / base, 32 bases per uint32 word; 1 = "not an upper-case A/C/G/T"
//        (codes L,N,X,E) and also 1 for every padding cell past the end of the block
//   rec: 16 bytes / 32 bases = {p2 lo, p2 hi, terminator bits, soft bits}: what the tile walk of the
//        filter kernels reads (one or two 16-byte loads per 32-cell window instead of four scalar
//        loads from two planes) and what the query rows of the screen are built from; REC_FRONT records
//        in front and PAD_WORDS behind are pure terminators
//   softmap: 1 bit / record, set where the record holds a "soft" cell (non-ACGT, not a terminator
//        under the current matrix).  The popcount screen reads the REFERENCE window from the bare p2
//        plane (48 bytes per hit instead of 96: a 500 Mb block is 125 MB and stays L2-resident) and
//        needs this map only for blocks that have soft cells at all (screen_bound.h)
// The hot kernels read only p2/m1 (0.375 byte per base instead of 1); b8 is touched by the
// exact extension only inside 32-base tiles that contain a masked cell.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/segalign_b200.h"
#include "zero_runs.h"

namespace sa {

constexpr int NUC = 8;
constexpr uint8_t A_NT = 0, C_NT = 1, G_NT = 2, T_NT = 3, L_NT = 4, N_NT = 5, X_NT = 6, E_NT = 7;
constexpr uint32_t INVALID_KMER = 1u << 31;
constexpr int PAD_WORDS = 4; // padding words appended to p2/m1 (all masked)

constexpr int REC_FRONT = 4; // padding records in front of record 0 (pure terminators)

struct SeqPlanes {
    uint8_t *b8 = nullptr;
    uint64_t *p2_base = nullptr; // p2 allocation: REC_FRONT zero words in front of word 0 (the screen reads words w-3 .. w+2)
    uint64_t *p2 = nullptr;      // = p2_base + REC_FRONT
    uint32_t *m1 = nullptr;
    uint32_t *softmap = nullptr; // one bit per record (index w + REC_FRONT): the record holds a soft cell; last word = soft record count
    uint32_t softmap_words = 0;  // bitmap words (the counter sits at softmap[softmap_words])
    uint32_t has_soft = 0;       // host copy of the counter: 0 = the screen never has to look at the map
    uint4 *rec_base = nullptr; // filter records {p2 lo, p2 hi, terminator bits, soft bits}, REC_FRONT + words
    uint4 *rec = nullptr;      // = rec_base + REC_FRONT (record 0 = bases 0..31)
    uint32_t term_codes = 0;   // terminator code set the records were built with
    // zero-run planes (stage B, screen_bound.h: zero_run_codes; built only under a matrix with flat codes): one
    // allocation = f1 | g1 (1 bit / base: the cell's code is flat / a partner) | F1k | G1k (1 bit / 1024 bases: every
    // cell of the aligned 1024-base piece exists and is flat / a partner) | flat-cell counter
    uint32_t *zr = nullptr, *f1 = nullptr, *g1 = nullptr, *F1k = nullptr, *G1k = nullptr;
    uint32_t coarse_words = 0;
    uint32_t zero_codes = 0xFFFFFFFFu; // F | G << 8 the planes were built with
    uint32_t has_flat = 0;     // host copy of the counter
    uint32_t len = 0;
    size_t words = 0; // p2/m1/rec words including back padding
};

// seed shape in constant-free form (passed by value to kernels)
struct ShapeDesc {
    int span;        // seed_size (19 for 12of19)
    int weight;      // kmer_size (12)
    int num_trans;   // number of care positions that allow a transition
    uint8_t pos[32];   // care positions, first = most significant
    uint8_t trans[32]; // transition_pos[t]
    uint8_t tvar[32];  // tvar[i] = care-position index t of the i-th transition variant (seeder.cpp:64-71)
};

struct ExtendParams {
    const uint8_t *rb8;
    const uint64_t *rp2;
    const uint32_t *rm1;
    uint32_t ref_len;
    const uint8_t *qb8;
    const uint64_t *qp2;
    const uint32_t *qm1;
    uint32_t query_len;
    int xdrop;
    int hspthresh;
    int noentropy;
    int diag_all_positive; // sub_mat[c][c] > 0 for c in ACGT: enables the all-match tile path
    int scores_fit_int8;   // ACGT x ACGT block within [-128,127]: enables the dp4a group path
    int soft_runs;         // lower case or N are NOT terminators under this matrix: walks pass through their runs
    uint32_t win_lo, win_hi; // repeat-masker variant: reference window of the call (0 .. 0xFFFFFFFF otherwise)
    // zero-run planes of both blocks (zskip != 0: one of the blocks has flat cells, see SeqPlanes and zero_runs.h)
    int zskip;
    ZeroPlanes rz, qz;
};

struct Anchor { // HSP + the reference iteration it belongs to (dedupe scope)
    uint32_t tag;
    uint32_t ref_start;
    uint32_t query_start;
    uint32_t len;
    int32_t score;
};

// ---------------------------------------------------------------- device helpers
// 32 bases starting at cell c of a p2 plane, cell c in bits 0..1
__device__ __forceinline__ uint64_t load_p2_window(const uint64_t *__restrict__ p2, uint32_t c) {
    uint32_t w = c >> 5, sh = (c & 31u) * 2u;
    uint64_t lo = __ldg(p2 + w);
    if (sh == 0) return lo;
    uint64_t hi = __ldg(p2 + w + 1);
    return (lo >> sh) | (hi << (64u - sh));
}
// mask bits of 32 cells starting at cell c, cell c in bit 0
__device__ __forceinline__ uint32_t load_m1_window(const uint32_t *__restrict__ m1, uint32_t c) {
    uint32_t w = c >> 5, sh = c & 31u;
    uint32_t lo = __ldg(m1 + w), hi = __ldg(m1 + w + 1);
    return __funnelshift_r(lo, hi, sh);
}

__device__ __forceinline__ uint8_t encode_ascii(uint8_t ch) {
    // common/seed_filter_interface.cu:28-45
    uint8_t d = X_NT;
    if (ch == 'A') d = A_NT;
    else if (ch == 'C') d = C_NT;
    else if (ch == 'G') d = G_NT;
    else if (ch == 'T') d = T_NT;
    else if (ch == 'a' || ch == 'c' || ch == 'g' || ch == 't') d = L_NT;
    else if (ch == 'n' || ch == 'N') d = N_NT;
    else if (ch == '&') d = E_NT;
    return d;
}

} // namespace sa
