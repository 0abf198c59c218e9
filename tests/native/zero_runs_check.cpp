// Property test of stage B's zero-run skipping (segalign_b200/csrc/zero_runs.h, compiled for the host).
//
// Random code sequences with runs of flat codes (N, IUPAC) from 1 base to hundreds of kilobases, lower case, separators,
// run starts / ends on and off the 32- and 1024-base grids, runs at both ends of a block.  The bit planes are built as
// k_pack_zero_planes / k_coarse_zero_planes build them (kernels_encode.cuh).  For random positions and both
// directions:
//   * zero_tile(rc0, qc0) must equal "all 32 cell pairs are (flat, partner) pairs in one orientation or the other";
//   * zero_jump(r, q, left) = k must be a multiple of 32 and SOUND: the next k cell pairs all exist and are such pairs
//     (checked cell by cell); and it must be USEFUL: a walk that takes 32 cells per zero tile and more where zero_jump
//     says so crosses a long stretch in a number of trips that does not grow with its length / 32.
// usage: zero_runs_check SEED N_QUERIES      prints "violations=0 ..." on success
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "../../segalign_b200/csrc/zero_runs.h"

using namespace sa;

struct Planes {
    std::vector<uint32_t> f1, g1, F1k, G1k;
    ZeroPlanes view() const { return ZeroPlanes{f1.data(), g1.data(), F1k.data(), G1k.data()}; }
};

static Planes build(const std::vector<uint8_t> &b8, uint32_t flat, uint32_t partners) {
    const size_t len = b8.size();
    const size_t words = (len + 31) / 32 + 4; // PAD_WORDS
    Planes P;
    P.f1.assign(words, 0); P.g1.assign(words, 0);
    for (size_t i = 0; i < len; i++) {
        const uint32_t c = b8[i] & 7u;
        if ((flat >> c) & 1u) P.f1[i >> 5] |= 1u << (i & 31);
        if ((partners >> c) & 1u) P.g1[i >> 5] |= 1u << (i & 31);
    }
    const size_t coarse_words = words / 1024 + 3;
    P.F1k.assign(coarse_words, 0); P.G1k.assign(coarse_words, 0);
    for (size_t pc = 0; pc < coarse_words * 32; pc++) {
        const size_t first = pc * 32;
        if (first + 32 > words) continue;
        bool af = true, ag = true;
        for (size_t k = 0; k < 32; k++) { af &= P.f1[first + k] == 0xFFFFFFFFu; ag &= P.g1[first + k] == 0xFFFFFFFFu; }
        if (af) P.F1k[pc >> 5] |= 1u << (pc & 31);
        if (ag) P.G1k[pc >> 5] |= 1u << (pc & 31);
    }
    return P;
}

int main(int argc, char **argv) {
    const unsigned seed = argc > 1 ? (unsigned)atoi(argv[1]) : 1;
    const long n_queries = argc > 2 ? atol(argv[2]) : 200000;
    std::mt19937_64 rng(seed);
    auto U = [&](uint64_t n) { return (uint64_t)(rng() % n); };
    // --ambiguous=iupac: F = {N, X}, G = everything but the separator; --ambiguous=n: F = {N}, G = ACGT + lower case + N
    const bool iupac = seed & 1;
    const uint32_t flat = iupac ? 0x60u : 0x20u, partners = iupac ? 0x7Fu : 0x3Fu;

    long violations = 0, tiles_true = 0, jumps = 0, jumped_cells = 0, deep = 0, deep_short = 0;
    for (int round = 0; round < 6; round++) {
        const size_t RL = 1000 + U(700000), QL = 1000 + U(700000);
        auto make = [&](size_t L) {
            std::vector<uint8_t> s(L);
            for (auto &c : s) { const uint64_t x = U(100); c = x < 70 ? (uint8_t)U(4) : (x < 97 ? 4 : (x < 98 ? 6 : (x < 99 ? 5 : 7))); }
            const int runs = 3 + (int)U(12);
            for (int i = 0; i < runs; i++) {
                size_t l = U(4) == 0 ? 1 + U(300000) : 1 + U(6000);
                size_t s0 = U(L);
                if (U(4) == 0) s0 &= ~(size_t)1023;            // on the coarse grid
                if (U(4) == 0) l = (l + 1023) & ~(size_t)1023;
                if (i == 0) s0 = 0;                            // a run at the start of the block
                if (i == 1) s0 = L > l ? L - l : 0;            // ... and at its end
                for (size_t k = s0; k < L && k < s0 + l; k++) s[k] = (iupac && U(50) == 0) ? 6 : 5;
            }
            return s;
        };
        const std::vector<uint8_t> rb = make(RL), qb = make(QL);
        const Planes RP = build(rb, flat, partners), QP = build(qb, flat, partners);
        const ZeroPlanes R = RP.view(), Q = QP.view();
        auto pair_ok = [&](size_t r, size_t q) {
            const uint32_t a = rb[r] & 7u, b = qb[q] & 7u;
            return (((flat >> a) & 1u) && ((partners >> b) & 1u)) || (((flat >> b) & 1u) && ((partners >> a) & 1u));
        };
        // positions biased towards flat cells of either block
        std::vector<size_t> rflat, qflat;
        for (size_t i = 0; i < RL; i += 7) if ((flat >> (rb[i] & 7u)) & 1u) rflat.push_back(i);
        for (size_t i = 0; i < QL; i += 7) if ((flat >> (qb[i] & 7u)) & 1u) qflat.push_back(i);
        for (long it = 0; it < n_queries / 6; it++) {
            size_t r = U(RL + 1), q = U(QL + 1);
            const uint64_t kind = U(4);
            if (kind == 0 && !rflat.empty()) r = rflat[U(rflat.size())];
            if (kind == 1 && !qflat.empty()) q = qflat[U(qflat.size())];
            if (kind == 2) { if (U(2)) r = U(2) ? 0 : RL; else q = U(2) ? 0 : QL; }
            // ---- tile
            if (r + 32 <= RL && q + 32 <= QL) {
                bool want = true;
                for (int j = 0; j < 32; j++) want &= pair_ok(r + j, q + j);
                const bool got = zero_tile(R, Q, (uint32_t)r, (uint32_t)q);
                tiles_true += want;
                if (got != want) { violations++; if (violations < 10) printf("tile r=%zu q=%zu want=%d got=%d\n", r, q, (int)want, (int)got); }
            }
            // ---- jump, both directions
            for (int left = 0; left < 2; left++) {
                const uint32_t k = zero_jump(R, Q, (uint32_t)r, (uint32_t)q, left != 0);
                size_t truth = 0; // consecutive good pairs from the next cell on
                if (!left) { while (r + truth < RL && q + truth < QL && pair_ok(r + truth, q + truth)) truth++; }
                else { while (truth < r && truth < q && pair_ok(r - 1 - truth, q - 1 - truth)) truth++; }
                if ((k & 31u) || k > truth) { violations++; if (violations < 10) printf("jump r=%zu q=%zu left=%d k=%u truth=%zu\n", r, q, left, k, truth); }
                if (k) { jumps++; jumped_cells += k; }
                // USEFUL: the loop of extend_dir (32 cells per zero tile, more where zero_jump says so) crosses a long good
                // stretch in a number of trips that does not grow with its length / 32: ragged ends of the runs on either
                // block's 1024-base grid cost up to 32 tiles each, the rest goes by whole pieces
                if (truth >= 16384) {
                    deep++;
                    size_t t = 0, trips = 0;
                    while (t + 32 <= truth) {
                        const size_t rr = left ? r - t : r + t, qq = left ? q - t : q + t;
                        const bool zt = left ? zero_tile(R, Q, (uint32_t)(rr - 32), (uint32_t)(qq - 32)) : zero_tile(R, Q, (uint32_t)rr, (uint32_t)qq);
                        if (!zt) { violations++; if (violations < 10) printf("tile inside a good stretch not recognised r=%zu q=%zu left=%d t=%zu\n", r, q, left, t); break; }
                        const uint32_t kk = zero_jump(R, Q, (uint32_t)rr, (uint32_t)qq, left != 0);
                        if (kk > truth - t) { violations++; break; }
                        t += kk > 32u ? kk : 32u;
                        trips++;
                    }
                    if (trips > 200 + truth / 4096) { deep_short++; violations++; if (violations < 10) printf("slow crossing r=%zu q=%zu left=%d truth=%zu trips=%zu\n", r, q, left, truth, trips); }
                }
            }
        }
    }
    printf("violations=%ld tiles_true=%ld jumps=%ld jumped_cells=%ld deep=%ld deep_short=%ld\n", violations, tiles_true, jumps,
           jumped_cells, deep, deep_short);
    return violations ? 1 : 0;
}
