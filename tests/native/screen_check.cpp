// screen_check.cpp -- host-side property test of the popcount screen (segalign_b200/csrc/screen_bound.h).
//
// Compiles the SAME screen code the CUDA filter kernel runs (screen_align / screen_walk /
// screen_reject are host+device functions) and checks, against the oracle's exact X-drop extension
// (oracle/sa_oracle.c: sao_extend_hit), that
//   (1) a rejected anchor never reaches hspthresh,
//   (2) whenever the screen calls an anchor "decided", its bound is >= the exact score.
// Inputs: random sequence pairs with planted homologies at several divergences, soft-masked runs,
// N runs, '&' separators, IUPAC letters, and anchors at the block edges.
// usage: screen_check <seed> <n_anchors> <ambiguous: ""|n|iupac> <xdrop> <hspthresh>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <vector>

#include "../../oracle/sa_oracle.h"
#include "../../segalign_b200/csrc/screen_bound.h"

using namespace sa;

// host restatement of k_pack_records (kernels_filter.cuh): records -4 .. words+3
struct Records {
    std::vector<ScreenRec> v;
    int front = 4;
    const ScreenRec &at(int w) const { return v[(size_t)(w + front)]; }
};
static Records pack(const std::vector<uint8_t> &b8, uint32_t term_codes) {
    Records R;
    const int words = (int)((b8.size() + 31) / 32) + 4;
    R.v.resize((size_t)words + R.front);
    for (int w = -R.front; w < words; w++) {
        uint64_t bits = 0; uint32_t term = 0, soft = 0;
        for (int cell = 0; cell < 32; cell++) {
            const long long pos = (long long)w * 32 + cell;
            if (pos < 0 || pos >= (long long)b8.size()) { term |= 1u << cell; continue; }
            const uint32_t c = b8[(size_t)pos];
            if (c < 4) bits |= (uint64_t)c << (2 * cell);
            else if ((term_codes >> c) & 1u) term |= 1u << cell;
            else soft |= 1u << cell;
        }
        ScreenRec r; r.x = (uint32_t)bits; r.y = (uint32_t)(bits >> 32); r.z = term; r.w = soft;
        R.v[(size_t)(w + R.front)] = r;
    }
    return R;
}
static void row_of(const Records &R, uint32_t anchor, uint32_t (&row)[SCREEN_ROW_WORDS]) {
    const int w = (int)(anchor >> 5);
    const ScreenRec a[SCREEN_RECS] = {R.at(w - 3), R.at(w - 2), R.at(w - 1), R.at(w), R.at(w + 1), R.at(w + 2)};
    screen_align(a, anchor & 31u, row);
}

// the reference side as the kernel sees it since round 2: the bare p2 words of records w-3 .. w+2 plus
// "any soft cell in those six records" from the one-bit-per-record map; terminator bits are not used
static void row_of_p2(const Records &R, uint32_t anchor, uint32_t (&row)[SCREEN_ROW_WORDS]) {
    const int w = (int)(anchor >> 5);
    uint64_t a[SCREEN_RECS];
    bool soft = false;
    for (int i = 0; i < SCREEN_RECS; i++) {
        const ScreenRec &r = R.at(w - 3 + i);
        a[i] = (uint64_t)r.x | ((uint64_t)r.y << 32);
        soft |= r.w != 0;
    }
    screen_align_p2(a, anchor & 31u, soft, row);
}

int main(int argc, char **argv) {
    const unsigned seed = argc > 1 ? (unsigned)atoi(argv[1]) : 1;
    const long n_anchors = argc > 2 ? atol(argv[2]) : 200000;
    const char *amb = argc > 3 ? argv[3] : "";
    const int xdrop = argc > 4 ? atoi(argv[4]) : 910;
    const int thresh = argc > 5 ? atoi(argv[5]) : 3000;
    std::mt19937_64 rng(seed);
    auto U = [&](uint64_t n) { return (uint64_t)(rng() % n); };
    const char ACGT[5] = "ACGT";

    const size_t RL = 60000, QL = 50000;
    std::string ref(RL, 'A'), qry(QL, 'A');
    for (auto &c : ref) c = ACGT[U(4)];
    for (auto &c : qry) c = ACGT[U(4)];
    // planted homologies: (ref pos, query pos, length, divergence)
    struct Hom { size_t r, q, len; };
    std::vector<Hom> homs;
    const double divs[] = {0.0, 0.05, 0.15, 0.25, 0.35, 0.45};
    for (int h = 0; h < 60; h++) {
        const size_t len = 30 + U(600), r = U(RL - len), q = U(QL - len);
        const double d = divs[U(6)];
        for (size_t i = 0; i < len; i++) {
            char c = ref[r + i];
            if ((double)U(1000000) / 1e6 < d) c = ACGT[(strchr(ACGT, c) - ACGT + 1 + U(3)) & 3];
            qry[q + i] = c;
        }
        homs.push_back({r, q, len});
    }
    // soft-masked runs, N runs, separators, IUPAC letters
    auto decorate = [&](std::string &s) {
        for (int k = 0; k < 40; k++) { size_t p = U(s.size() - 400), l = 1 + U(300); for (size_t i = 0; i < l; i++) s[p + i] = (char)tolower(s[p + i]); }
        for (int k = 0; k < 6; k++) { size_t p = U(s.size() - 200), l = 1 + U(100); for (size_t i = 0; i < l; i++) s[p + i] = 'N'; }
        for (int k = 0; k < 4; k++) s[U(s.size())] = '&';
        for (int k = 0; k < 30; k++) s[U(s.size())] = "RYKMSW"[U(6)];
    };
    decorate(ref); decorate(qry);

    sao_params P; memset(&P, 0, sizeof(P));
    sao_build_matrix(amb, xdrop, P.sub_mat);
    P.xdrop = xdrop; P.hspthresh = 0; P.noentropy = 1; P.seed_size = 19; P.max_hits = 1u << 30;
    std::vector<uint8_t> rb(RL), qb(QL);
    sao_encode(ref.data(), (uint32_t)RL, rb.data());
    sao_encode(qry.data(), (uint32_t)QL, qb.data());
    const uint32_t term_codes = screen_terminator_codes(P.sub_mat, xdrop); // as sa_initialize_processor
    const Records RR = pack(rb, term_codes), QR = pack(qb, term_codes);
    const ScreenConsts C = screen_consts_from_matrix(P.sub_mat, xdrop, thresh);
    if (!C.enabled) { printf("screen disabled for this matrix\n"); return 0; }

    long rejected = 0, decided_n = 0, passing = 0, bad = 0, rejected_full = 0;
    for (long it = 0; it < n_anchors; it++) {
        uint32_t r0, q0;
        const int kind = (int)U(10);
        if (kind < 4) { r0 = (uint32_t)U(RL + 1); q0 = (uint32_t)U(QL + 1); }
        else if (kind < 8) { // on or next to a planted diagonal
            const Hom &h = homs[U(homs.size())];
            const long off = (long)U(h.len + 200) - 100;
            long r = (long)h.r + off, q = (long)h.q + off + (U(8) == 0 ? 1 : 0);
            if (r < 0 || q < 0 || r > (long)RL || q > (long)QL) continue;
            r0 = (uint32_t)r; q0 = (uint32_t)q;
        } else if (kind == 8) { r0 = (uint32_t)U(140); q0 = (uint32_t)U(140); }
        else { r0 = (uint32_t)(RL - U(140)); q0 = (uint32_t)(QL - U(140)); }
        sao_segment seg;
        const int ok = sao_extend_hit(&P, rb.data(), (uint32_t)RL, qb.data(), (uint32_t)QL, r0, q0, &seg);
        const int exact = ok ? seg.score : 0;
        uint32_t rr[SCREEN_ROW_WORDS], qr[SCREEN_ROW_WORDS], rr_full[SCREEN_ROW_WORDS];
        row_of_p2(RR, r0, rr); row_of(QR, q0, qr);
        int bound; bool decided;
        const bool rej = screen_reject(rr, qr, C, bound, decided);
        {   // for comparison: the reference window with its terminator / soft bits (16-byte records)
            row_of(RR, r0, rr_full);
            int b2; bool d2;
            const bool rej_full = screen_reject(rr_full, qr, C, b2, d2);
            if (rej_full) rejected_full++;
            if ((rej_full && exact >= thresh) || (d2 && b2 < exact)) {
                if (bad < 10) fprintf(stderr, "VIOLATION (record window) r0=%u q0=%u exact=%d bound=%d\n", r0, q0, exact, b2);
                bad++;
            }
        }
        if (exact >= thresh) passing++;
        if (decided) decided_n++;
        if (rej) rejected++;
        if ((rej && exact >= thresh) || (decided && bound < exact)) {
            if (bad < 10) fprintf(stderr, "VIOLATION r0=%u q0=%u exact=%d bound=%d decided=%d rej=%d\n", r0, q0, exact, bound, (int)decided, (int)rej);
            bad++;
        }
    }
    printf("anchors=%ld decided=%ld rejected=%ld passing=%ld violations=%ld rejected_with_record_flags=%ld\n", n_anchors, decided_n, rejected, passing, bad, rejected_full);
    return bad ? 1 : 0;
}
