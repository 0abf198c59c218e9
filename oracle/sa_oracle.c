/*
 * sa_oracle.c -- CPU restatement of SegAlign's seed-filter-extend hot path (plain C, sequential).
 *
 * TEST INFRASTRUCTURE ONLY -- see sa_oracle.h.  The product path never calls into this file.
 *
 * The restatement follows the reference's observable behaviour including its quirks
 * (SURVEY.md Appendix A).  Where the reference has undefined behaviour this file documents the
 * definition it adopts; parity tests stay out of those zones.
 *
 * Numerics caveat (src/seed_filter.cu:619-623): the entropy factor uses CUDA's device
 * log(double); this file uses the host libm log().  Both are within 1 ulp, so an (int)
 * truncation can flip only if score*entropy lies within ~1e-13 of an integer.
 */
#include "sa_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ encoding */

/* common/parameters.h:5-13 */
enum { A_NT = 0, C_NT = 1, G_NT = 2, T_NT = 3, L_NT = 4, N_NT = 5, X_NT = 6, E_NT = 7 };

static inline uint8_t encode_char(char ch) {
    /* common/seed_filter_interface.cu:28-45 */
    switch (ch) {
        case 'A': return A_NT;
        case 'C': return C_NT;
        case 'G': return G_NT;
        case 'T': return T_NT;
        case 'a': case 'c': case 'g': case 't': return L_NT;
        case 'n': case 'N': return N_NT;
        case '&': return E_NT;
        default: return X_NT;
    }
}

void sao_encode(const char *src, uint32_t len, uint8_t *dst) {
    for (uint32_t i = 0; i < len; i++) dst[i] = encode_char(src[i]);
}

void sao_encode_rc(const char *src, uint32_t len, uint8_t *dst, uint8_t *dst_rc) {
    /* src/seed_filter.cu:120-154: A<->T, C<->G; L, N, E map to themselves; the rest to X */
    for (uint32_t i = 0; i < len; i++) {
        uint8_t c = encode_char(src[i]);
        dst[i] = c;
        dst_rc[len - 1 - i] = (c < 4) ? (uint8_t)(3 - c) : c;
    }
}

void sao_revcomp_ascii(char *dst, const char *src, size_t len) {
    /* common/ntcoding.cpp:63-105; characters outside the switch are skipped by the reference
     * (it prints a warning and does not advance); inputs here are restricted to its alphabet. */
    size_t r = 0;
    for (size_t i = len; i > 0; i--) {
        char c = src[i - 1], o;
        switch (c) {
            case 'a': o = 't'; break;
            case 'A': o = 'T'; break;
            case 'c': o = 'g'; break;
            case 'C': o = 'G'; break;
            case 'g': o = 'c'; break;
            case 'G': o = 'C'; break;
            case 't': o = 'a'; break;
            case 'T': o = 'A'; break;
            case 'n': o = 'n'; break;
            case 'N': o = 'N'; break;
            case '&': o = '&'; break;
            default: continue;
        }
        dst[r++] = o;
    }
}

/* ------------------------------------------------------------------ seed words */

int sao_shape_init(sao_shape *sh, const char *seed_shape) {
    /* src/main.cpp:160-178 */
    char shape[64];
    memset(sh, 0, sizeof(*sh));
    if (strcmp(seed_shape, "12of19") == 0) {
        strcpy(shape, "TTT0T00TT00T0T0TTTT");
    } else if (strcmp(seed_shape, "14of22") == 0) {
        strcpy(shape, "TTT0T0TT00TT00T0T0TTTT");
    } else {
        size_t n = strlen(seed_shape);
        if (n > 32) n = 32;
        for (size_t i = 0; i < n; i++) shape[i] = (seed_shape[i] == '1') ? 'T' : '0';
        shape[n] = 0;
    }
    /* common/ntcoding.cpp:21-37 */
    int w = 0;
    int n = (int)strlen(shape);
    for (int i = 0; i < n; i++) {
        if (shape[i] == '1' || shape[i] == 'T') {
            sh->shape_pos[w] = i;
            sh->transition_pos[w] = (shape[i] == 'T');
            w++;
        }
    }
    sh->weight = w;
    sh->span = n;
    return w;
}

uint32_t sao_kmer_at(const sao_shape *sh, const char *seq, size_t pos) {
    /* common/ntcoding.cpp:43-61: every one of the span characters (don't-care positions
     * included) must be an upper-case A/C/G/T, otherwise INVALID_KMER. */
    uint32_t nt[64];
    for (int i = 0; i < sh->span; i++) {
        switch (seq[pos + i]) {
            case 'A': nt[i] = A_NT; break;
            case 'C': nt[i] = C_NT; break;
            case 'G': nt[i] = G_NT; break;
            case 'T': nt[i] = T_NT; break;
            default: return SAO_INVALID_KMER;
        }
    }
    uint32_t kmer = 0;
    for (int i = 0; i < sh->weight; i++) kmer = (kmer << 2) + nt[sh->shape_pos[i]];
    return kmer;
}

size_t sao_chunk_seeds(const sao_shape *sh, int transition, const char *seq, size_t block_start,
                       uint32_t j0, uint32_t j1, uint64_t *out) {
    /* src/seeder.cpp:57-74 (identical loop at :94-109 for the minus strand) */
    size_t n = 0;
    for (uint32_t j = j0; j < j1; j++) {
        uint64_t kmer = sao_kmer_at(sh, seq, block_start + j);
        if (kmer != SAO_INVALID_KMER) {
            out[n++] = (kmer << 32) + j;
            if (transition) {
                for (int t = 0; t < sh->weight; t++) {
                    if (sh->transition_pos[t] == 1) {
                        uint64_t tr = kmer ^ ((uint64_t)2 << (2 * t)); /* TRANSITION_MASK */
                        out[n++] = (tr << 32) + j;
                    }
                }
            }
        }
    }
    return n;
}

/* ------------------------------------------------------------------ matrix */

void sao_build_matrix(const char *ambiguous, int xdrop, int *sub_mat) {
    /* src/main.cpp:187-268 */
    int ambiguous_reward = -100, ambiguous_penalty = -100;
    const int fill_score = -100, bad_score = -1000;
    char field[32] = "x";
    const char *amb = ambiguous ? ambiguous : "";
    /* boost::split on ',' (:193-199): three fields => x,R,P */
    const char *c1 = strchr(amb, ',');
    size_t flen = c1 ? (size_t)(c1 - amb) : strlen(amb);
    if (flen >= sizeof(field)) flen = sizeof(field) - 1;
    memcpy(field, amb, flen);
    field[flen] = 0;
    if (c1) {
        const char *c2 = strchr(c1 + 1, ',');
        if (c2 && !strchr(c2 + 1, ',')) {
            ambiguous_reward = atoi(c1 + 1);
            ambiguous_penalty = -1 * atoi(c2 + 1);
        }
    } else if (strcmp(amb, "n") == 0 || strcmp(amb, "iupac") == 0) {
        ambiguous_reward = 0;
        ambiguous_penalty = 0;
    }
    memset(sub_mat, 0, 64 * sizeof(int));
    static const int acgt[4][4] = {{91, -114, -31, -123},
                                   {-114, 100, -125, -31},
                                   {-31, -125, 100, -114},
                                   {-123, -31, -114, 91}};
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) sub_mat[i * 8 + j] = acgt[i][j];
    for (int i = 0; i < L_NT; i++) {
        sub_mat[i * 8 + L_NT] = bad_score;
        sub_mat[L_NT * 8 + i] = bad_score;
    }
    sub_mat[L_NT * 8 + L_NT] = bad_score;
    if (strcmp(field, "n") == 0 || strcmp(field, "iupac") == 0) {
        for (int i = 0; i < N_NT; i++) {
            sub_mat[i * 8 + N_NT] = ambiguous_penalty;
            sub_mat[N_NT * 8 + i] = ambiguous_penalty;
        }
        sub_mat[N_NT * 8 + N_NT] = ambiguous_reward;
    } else {
        for (int i = 0; i < N_NT; i++) {
            sub_mat[i * 8 + N_NT] = bad_score;
            sub_mat[N_NT * 8 + i] = bad_score;
        }
        sub_mat[N_NT * 8 + N_NT] = bad_score;
    }
    if (strcmp(field, "iupac") == 0) {
        for (int i = 0; i < X_NT; i++) {
            sub_mat[i * 8 + X_NT] = ambiguous_penalty;
            sub_mat[X_NT * 8 + i] = ambiguous_penalty;
        }
        sub_mat[X_NT * 8 + X_NT] = ambiguous_reward;
    } else {
        for (int i = 0; i < L_NT; i++) {
            sub_mat[i * 8 + X_NT] = fill_score;
            sub_mat[X_NT * 8 + i] = fill_score;
        }
        for (int i = L_NT; i < X_NT; i++) {
            sub_mat[i * 8 + X_NT] = bad_score;
            sub_mat[X_NT * 8 + i] = bad_score;
        }
        sub_mat[X_NT * 8 + X_NT] = fill_score;
    }
    for (int i = 0; i < E_NT; i++) {
        sub_mat[i * 8 + E_NT] = -10 * xdrop;
        sub_mat[E_NT * 8 + i] = -10 * xdrop;
    }
    sub_mat[E_NT * 8 + E_NT] = -10 * xdrop;
}

/* ------------------------------------------------------------------ seed position table */

int sao_table_build(sao_table *t, const sao_shape *sh, const char *ref, size_t start_addr,
                    uint32_t ref_length, uint32_t step) {
    /* common/seed_pos_table.cu:58-64 */
    uint32_t shape_size = (uint32_t)sh->span;
    uint32_t offset = (shape_size + 1) % step;
    uint32_t start_offset = step - offset;
    uint32_t index_table_size = ((uint32_t)1 << (2 * sh->weight)) + 1;
    uint32_t num_steps = (ref_length - shape_size + offset) / step;
    if (ref_length < shape_size) num_steps = 0; /* reference would wrap; out of scope */

    uint32_t *index_table = (uint32_t *)calloc(index_table_size, sizeof(uint32_t));
    uint32_t *kmers = (uint32_t *)malloc((size_t)(num_steps ? num_steps : 1) * sizeof(uint32_t));
    if (!index_table || !kmers) return -1;
    /* pass 1 (:69-81): histogram at index+1 */
    for (uint32_t i = 0; i < num_steps; i++) {
        uint32_t k = sao_kmer_at(sh, ref, start_addr + start_offset + (size_t)i * step);
        kmers[i] = k;
        if (k != SAO_INVALID_KMER) index_table[k + 1]++;
    }
    /* :83 inclusive scan over the whole table */
    for (uint32_t i = 1; i < index_table_size; i++) index_table[i] += index_table[i - 1];
    uint32_t num_index = index_table[index_table_size - 1];
    uint32_t *pos_table = (uint32_t *)malloc((size_t)(num_index ? num_index : 1) * sizeof(uint32_t));
    uint32_t *cursor = (uint32_t *)malloc((size_t)index_table_size * sizeof(uint32_t));
    if (!pos_table || !cursor) return -1;
    memcpy(cursor, index_table, (size_t)index_table_size * sizeof(uint32_t));
    /* pass 2 (:89-101).  The reference's order inside a bucket depends on TBB scheduling;
     * any order is conformant.  Here: ascending position. */
    for (uint32_t i = 0; i < num_steps; i++) {
        uint32_t k = kmers[i];
        if (k != SAO_INVALID_KMER) pos_table[cursor[k]++] = start_offset + i * step;
    }
    free(cursor);
    free(kmers);
    /* :103 the device sees index_table+1 */
    t->index_size = index_table_size - 1;
    t->index = (uint32_t *)malloc((size_t)t->index_size * sizeof(uint32_t));
    if (!t->index) return -1;
    memcpy(t->index, index_table + 1, (size_t)t->index_size * sizeof(uint32_t));
    free(index_table);
    t->pos = pos_table;
    t->num_pos = num_index;
    return 0;
}

void sao_table_free(sao_table *t) {
    free(t->index);
    free(t->pos);
    t->index = t->pos = NULL;
}

/* ------------------------------------------------------------------ extension */

/*
 * One direction of find_hsps, kept in the reference's 32-cell tiles because the entropy
 * counters are updated per tile (src/seed_filter.cu:436-451, :587-602).
 * frame[0..3] = count[], frame[4..7] = count_del[]: the reference indexes count[] and
 * count_del[] (adjacent short[4] arrays in one 16-byte local frame) with codes up to 7, so
 * count[c] for c>=4 lands in count_del[c-4] and count_del[c] for c>=4 falls outside the
 * frame.
 */
static void extend_direction(const sao_params *p, const uint8_t *ref, uint32_t ref_len,
                             const uint8_t *qry, uint32_t query_len, uint32_t r0, uint32_t q0,
                             int left, short frame[8], int *best_score, int *best_pos) {
    int prev_score = 0, prev_max_score = 0;
    int prev_max_pos = left ? 0 : -1; /* :310 / :467 */
    uint32_t tile = 0;
    frame[4] = frame[5] = frame[6] = frame[7] = 0; /* :318-321 / :471-474 */

    for (;;) {
        uint8_t rc[32], qc[32];
        int inb[32];
        int s = prev_score, M = prev_max_score, mp = prev_max_pos;
        int xdrop_done = 0;
        int last_oob = 0;
        int pos_of[32];
        for (int lane = 0; lane < 32; lane++) {
            int v = 0;
            uint32_t pos_offset = left ? (uint32_t)lane + 1 + tile : (uint32_t)lane + tile;
            pos_of[lane] = (int)pos_offset;
            int in;
            if (!left) {
                uint32_t rp = r0 + pos_offset, qp = q0 + pos_offset; /* :328-332 */
                in = (rp < ref_len && qp < query_len);
                if (in) { rc[lane] = ref[rp]; qc[lane] = qry[qp]; }
            } else {
                in = (r0 >= pos_offset && q0 >= pos_offset); /* :482 */
                if (in) { rc[lane] = ref[r0 - pos_offset]; qc[lane] = qry[q0 - pos_offset]; }
            }
            inb[lane] = in;
            if (in) v = p->sub_mat[rc[lane] * 8 + qc[lane]];
            if (lane == 31) last_oob = !in;
            if (!xdrop_done) {
                /* sequential equivalent of the four shuffle scans (:339-403) */
                s += v;
                if (s > M) { M = s; mp = (int)pos_offset; }
                if (M - s > p->xdrop) xdrop_done = 1;
            }
        }
        /* when an x-drop fired, s stopped at the dropping cell; the running prefix at lane 31
         * is only needed when the loop continues, i.e. when no x-drop fired */
        int new_max_found = (mp > prev_max_pos); /* :408-411 */
        int stop = 0;
        if (xdrop_done || last_oob) { /* :413-426 */
            stop = 1;
            prev_max_pos = mp;
        } else { /* :427-432 */
            prev_score = s;
            prev_max_score = M;
            prev_max_pos = mp;
            tile += 32;
        }
        if (new_max_found) { /* :436-441 */
            for (int i = 0; i < 4; i++) {
                frame[i] = (short)(frame[i] + frame[4 + i]);
                frame[4 + i] = 0;
            }
        }
        for (int lane = 0; lane < 32; lane++) { /* :444-451 */
            if (!inb[lane]) continue; /* stale registers on out-of-bounds lanes only ever reach
                                         count_del of a direction's final tile: no effect */
            if (rc[lane] == qc[lane]) {
                int c = rc[lane];
                int idx = (pos_of[lane] <= prev_max_pos) ? c : 4 + c;
                if (idx < 8) frame[idx] = (short)(frame[idx] + 1);
            }
        }
        if (stop) {
            *best_score = M;
            *best_pos = mp;
            return;
        }
    }
}

int sao_extend_hit(const sao_params *p, const uint8_t *ref, uint32_t ref_len, const uint8_t *qry,
                   uint32_t query_len, uint32_t r0, uint32_t q0, sao_segment *out) {
    short frame[8] = {0, 0, 0, 0, 0, 0, 0, 0}; /* :314-321 */
    int right_score, right_pos, left_score, left_pos;
    extend_direction(p, ref, ref_len, qry, query_len, r0, q0, 0, frame, &right_score, &right_pos);
    extend_direction(p, ref, ref_len, qry, query_len, r0, q0, 1, frame, &left_score, &left_pos);
    int total = right_score + left_score;          /* :414/:421 + :563/:571 */
    uint32_t left_extent = (uint32_t)left_pos;     /* :565 */
    int extent = right_pos + (int)left_extent;     /* :416 + :566 */

    double entropy = 1.0; /* :307 */
    if (total >= p->hspthresh && total <= 3 * p->hspthresh && !p->noentropy) { /* :608 */
        int c0 = frame[0], c1 = frame[1], c2 = frame[2], c3 = frame[3];
        if (c0 + c1 + c2 + c3 >= 20) { /* :617 */
            double L = (double)(extent + 1);
            int cnt[4] = {c0, c1, c2, c3};
            double e = 0.0;
            for (int i = 0; i < 4; i++) { /* :620-622, fused multiply-add as nvcc contracts it */
                double pr = (double)cnt[i] / L;
                double lg = (cnt[i] != 0) ? log(pr) : 0.0;
                e = fma(pr, lg, e);
            }
            /* :623 log(4.0f) is the float overload: divide by -(double)(float)ln4 */
            entropy = -e / (double)1.38629436111989061883f;
        }
    }
    if ((int)(((float)total) * entropy) >= p->hspthresh) { /* :633 */
        out->ref_start = r0 - left_extent;
        out->query_start = q0 - left_extent;
        out->len = (uint32_t)extent;
        out->score = 0;
        if (entropy > 0) out->score = (int)((double)total * entropy); /* :637-638 */
        return 1;
    }
    return 0;
}

/* ------------------------------------------------------------------ sort / dedupe */

static inline uint32_t diag_of(const sao_segment *x) { return x->ref_start - x->query_start; }

static int cmp_hspComp(const void *a, const void *b) {
    /* src/seed_filter.cu:54-80: (diag, ref_start, len, score desc), all unsigned but score */
    const sao_segment *x = (const sao_segment *)a, *y = (const sao_segment *)b;
    uint32_t dx = diag_of(x), dy = diag_of(y);
    if (dx != dy) return dx < dy ? -1 : 1;
    if (x->ref_start != y->ref_start) return x->ref_start < y->ref_start ? -1 : 1;
    if (x->len != y->len) return x->len < y->len ? -1 : 1;
    if (x->score != y->score) return x->score > y->score ? -1 : 1;
    return 0;
}

static int cmp_hspCompLastz(const void *a, const void *b) {
    /* src/seed_filter.cu:82-108: (query_start, ref_start, len, score desc) */
    const sao_segment *x = (const sao_segment *)a, *y = (const sao_segment *)b;
    if (x->query_start != y->query_start) return x->query_start < y->query_start ? -1 : 1;
    if (x->ref_start != y->ref_start) return x->ref_start < y->ref_start ? -1 : 1;
    if (x->len != y->len) return x->len < y->len ? -1 : 1;
    if (x->score != y->score) return x->score > y->score ? -1 : 1;
    return 0;
}

static int hspEqual(const sao_segment *x, const sao_segment *y) {
    /* src/seed_filter.cu:47-52, u32 wrap-around arithmetic */
    return (diag_of(x) == diag_of(y)) &&
           (((x->ref_start >= y->ref_start) &&
             ((uint32_t)(x->ref_start + x->len) <= (uint32_t)(y->ref_start + y->len))) ||
            ((y->ref_start >= x->ref_start) &&
             ((uint32_t)(y->ref_start + y->len) <= (uint32_t)(x->ref_start + x->len))));
}

size_t sao_sort_dedupe(sao_segment *a, size_t n) {
    if (n == 0) return 0;
    /* :776 -- the comparator is a total order on distinct records, so stability is moot */
    qsort(a, n, sizeof(sao_segment), cmp_hspComp);
    /* :778 thrust::unique_copy compares each element with its predecessor IN THE INPUT
     * (head flags), not with the last element kept */
    sao_segment *tmp = (sao_segment *)malloc(n * sizeof(sao_segment));
    size_t m = 0;
    for (size_t i = 0; i < n; i++) {
        if (i == 0 || !hspEqual(&a[i - 1], &a[i])) tmp[m++] = a[i];
    }
    /* :782 */
    qsort(tmp, m, sizeof(sao_segment), cmp_hspCompLastz);
    memcpy(a, tmp, m * sizeof(sao_segment));
    free(tmp);
    return m;
}

/* ------------------------------------------------------------------ SeedAndFilter */

static uint32_t lower_bound_u32(const uint32_t *a, uint32_t n, uint32_t v) {
    uint32_t lo = 0, hi = n;
    while (lo < hi) {
        uint32_t mid = lo + (hi - lo) / 2;
        if (a[mid] < v) lo = mid + 1; else hi = mid;
    }
    return lo;
}

int sao_iteration_plan(const uint32_t *prefix, uint32_t num_seeds, uint32_t max_hits,
                       uint32_t *limit_pos) {
    /* src/seed_filter.cu:718-745.
     * Reference UB (SURVEY A.11 i): lower_bound returning 0 makes pos wrap to 0xFFFFFFFF and
     * reads out of range.  Definition adopted here: such an iteration is empty, i.e. the
     * boundary is "before seed 0" (limit_pos = 0xFFFFFFFF, next iteration starts at seed 0)
     * and the hit limit restarts from 0. */
    if (num_seeds == 0) return 0;
    uint32_t num_hits = prefix[num_seeds - 1];
    if (num_hits == 0) return 0; /* :754 guard */
    int num_iter;
    uint32_t iter_hit_limit;
    if (num_hits < max_hits) { num_iter = 2; iter_hit_limit = num_hits; }
    else { num_iter = (int)(num_hits / max_hits) + 2; iter_hit_limit = max_hits; }
    for (int i = 0; i < num_iter - 1; i++) {
        uint32_t pos = lower_bound_u32(prefix, num_seeds, iter_hit_limit) - 1;
        limit_pos[i] = pos;
        uint32_t base = (pos == 0xFFFFFFFFu) ? 0 : prefix[pos];
        iter_hit_limit = base + max_hits;
        if (iter_hit_limit > num_hits) iter_hit_limit = num_hits;
    }
    limit_pos[num_iter - 1] = num_seeds - 1;
    if (limit_pos[num_iter - 1] == limit_pos[num_iter - 2]) num_iter--;
    return num_iter;
}

sao_segment *sao_seed_and_filter(const sao_params *p, const sao_table *t, const uint8_t *ref,
                                 uint32_t ref_len, const uint8_t *qry, uint32_t query_len,
                                 const uint64_t *seeds, uint32_t num_seeds, size_t *out_n) {
    /* :712-714 bucket sizes + inclusive scan (uint32 wrap-around) */
    uint32_t *prefix = (uint32_t *)malloc((size_t)(num_seeds ? num_seeds : 1) * sizeof(uint32_t));
    uint32_t acc = 0;
    for (uint32_t i = 0; i < num_seeds; i++) {
        uint32_t seed = (uint32_t)(seeds[i] >> 32);
        uint32_t n = t->index[seed];
        if (seed > 0) n -= t->index[seed - 1];
        acc += n;
        prefix[i] = acc;
    }
    uint32_t num_hits = num_seeds ? prefix[num_seeds - 1] : 0; /* :716 */
    uint32_t cap_iter = (p->max_hits ? num_hits / p->max_hits : 0) + 2;
    uint32_t *limit_pos = (uint32_t *)malloc((size_t)cap_iter * sizeof(uint32_t));
    int num_iter = sao_iteration_plan(prefix, num_seeds, p->max_hits, limit_pos);

    size_t cap = 1024, n_out = 1;
    sao_segment *out = (sao_segment *)malloc(cap * sizeof(sao_segment));
    uint32_t total_anchors = 0;

    uint32_t start_seed = 0;
    for (int it = 0; it < num_iter; it++) {
        /* :756-793; an empty iteration (reference UB, A.11 ii) yields no anchors here */
        uint32_t end_seed = limit_pos[it] + 1; /* 0xFFFFFFFF+1 == 0: empty */
        size_t acap = 1024, an = 0;
        sao_segment *anch = (sao_segment *)malloc(acap * sizeof(sao_segment));
        for (uint32_t s = start_seed; s < end_seed; s++) {
            uint32_t seed = (uint32_t)(seeds[s] >> 32);
            uint32_t q0 = (uint32_t)(seeds[s] & 0xFFFFFFFFu) + p->seed_size; /* :204 */
            uint32_t end = t->index[seed];
            uint32_t start = seed > 0 ? t->index[seed - 1] : 0;
            for (uint32_t e = start; e < end; e++) {
                uint32_t r0 = t->pos[e] + p->seed_size; /* :220 */
                sao_segment seg;
                if (sao_extend_hit(p, ref, ref_len, qry, query_len, r0, q0, &seg)) {
                    if (an == acap) {
                        acap *= 2;
                        anch = (sao_segment *)realloc(anch, acap * sizeof(sao_segment));
                    }
                    anch[an++] = seg;
                }
            }
        }
        size_t kept = sao_sort_dedupe(anch, an);
        if (n_out + kept > cap) {
            while (n_out + kept > cap) cap *= 2;
            out = (sao_segment *)realloc(out, cap * sizeof(sao_segment));
        }
        memcpy(out + n_out, anch, kept * sizeof(sao_segment));
        n_out += kept;
        total_anchors += (uint32_t)kept;
        free(anch);
        start_seed = end_seed;
    }
    /* :806-809 header; ref_start/query_start are uninitialised in the reference */
    out[0].ref_start = 0;
    out[0].query_start = 0;
    out[0].len = total_anchors;
    out[0].score = (int32_t)num_hits;
    free(prefix);
    free(limit_pos);
    *out_n = n_out;
    return out;
}

/* ------------------------------------------------------------------ repeat-masker variant
 *
 * repeat_masker_src/seed_filter.cu: the sequence block is aligned against itself (plus strand) or
 * against its own reverse complement built on the device (:138-168, :951-961); hits whose reference
 * anchor lies outside the caller's window [ref_start, ref_end] are enumerated and counted but not
 * extended (:239-244, :305-310, :328-333); minus-strand records are mapped back to forward coordinates
 * (:705-709); three sorts and two unique passes (:819-835); 64-bit hit and anchor totals in the
 * header (:856-861). */

void sao_rm_revcomp_codes(const uint8_t *src, uint32_t len, uint8_t *dst) {
    /* :138-168 rev_comp_string: on the 8-symbol codes, A<->T, C<->G, every other code unchanged */
    for (uint32_t i = 0; i < len; i++) {
        uint8_t c = src[i], r = c;
        if (c == 0) r = 3; else if (c == 1) r = 2; else if (c == 2) r = 1; else if (c == 3) r = 0;
        dst[len - 1 - i] = r;
    }
}

static int rm_cmp_hspComp(const void *a, const void *b) {
    /* :109-135: (query_start, len desc, ref_start, score desc) -- all four fields: a total order */
    const sao_segment *x = (const sao_segment *)a, *y = (const sao_segment *)b;
    if (x->query_start != y->query_start) return x->query_start < y->query_start ? -1 : 1;
    if (x->len != y->len) return x->len > y->len ? -1 : 1;
    if (x->ref_start != y->ref_start) return x->ref_start < y->ref_start ? -1 : 1;
    if (x->score != y->score) return x->score > y->score ? -1 : 1;
    return 0;
}
static int rm_hspEqual(const sao_segment *x, const sao_segment *y) {
    /* :79-84 */
    return x->ref_start == y->ref_start && x->query_start == y->query_start && x->len == y->len && x->score == y->score;
}
static int rm_less_hspDiagComp(const sao_segment *x, const sao_segment *y) {
    /* :52-77: (diagonal [u32 wrap-around], ref_start, query_start, score desc) */
    uint32_t dx = diag_of(x), dy = diag_of(y);
    if (dx != dy) return dx < dy;
    if (x->ref_start != y->ref_start) return x->ref_start < y->ref_start;
    if (x->query_start != y->query_start) return x->query_start < y->query_start;
    return x->score > y->score;
}
static int rm_less_hspFinalComp(const sao_segment *x, const sao_segment *y) {
    /* :86-107: (query_start, score desc, ref_start desc) */
    if (x->query_start != y->query_start) return x->query_start < y->query_start;
    if (x->score != y->score) return x->score > y->score;
    return x->ref_start > y->ref_start;
}
/* thrust::stable_sort: bottom-up merge sort, equal elements keep their order */
static void stable_sort_seg(sao_segment *a, size_t n, int (*less)(const sao_segment *, const sao_segment *)) {
    if (n < 2) return;
    sao_segment *tmp = (sao_segment *)malloc(n * sizeof(sao_segment));
    sao_segment *src = a, *dst = tmp;
    for (size_t w = 1; w < n; w *= 2) {
        for (size_t lo = 0; lo < n; lo += 2 * w) {
            size_t mid = lo + w < n ? lo + w : n, hi = lo + 2 * w < n ? lo + 2 * w : n;
            size_t i = lo, j = mid, k = lo;
            while (i < mid && j < hi) dst[k++] = less(&src[j], &src[i]) ? src[j++] : src[i++];
            while (i < mid) dst[k++] = src[i++];
            while (j < hi) dst[k++] = src[j++];
        }
        sao_segment *t = src; src = dst; dst = t;
    }
    if (src != a) memcpy(a, src, n * sizeof(sao_segment));
    free(tmp);
}

size_t sao_rm_sort_dedupe(sao_segment *a, size_t n) {
    /* :819-835 */
    if (n == 0) return 0;
    qsort(a, n, sizeof(sao_segment), rm_cmp_hspComp);            /* :819 (total order: stability moot) */
    sao_segment *tmp = (sao_segment *)malloc(n * sizeof(sao_segment));
    size_t m = 0;
    for (size_t i = 0; i < n; i++)                                /* :821 unique_copy(hspEqual): predecessor in the input */
        if (i == 0 || !rm_hspEqual(&a[i - 1], &a[i])) tmp[m++] = a[i];
    stable_sort_seg(tmp, m, rm_less_hspDiagComp);                 /* :825 */
    size_t k = 0;
    for (size_t i = 0; i < m; i++)                                /* :827 unique_copy(hspDiagEqual) */
        if (i == 0 || !hspEqual(&tmp[i - 1], &tmp[i])) a[k++] = tmp[i];
    stable_sort_seg(a, k, rm_less_hspFinalComp);                  /* :833 */
    free(tmp);
    return k;
}

static uint32_t lower_bound_u64(const uint64_t *a, uint32_t n, uint64_t v) {
    uint32_t lo = 0, hi = n;
    while (lo < hi) {
        uint32_t mid = lo + (hi - lo) / 2;
        if (a[mid] < v) lo = mid + 1; else hi = mid;
    }
    return lo;
}

sao_segment *sao_rm_seed_and_filter(const sao_params *p, const sao_table *t, const uint8_t *seq,
                                    const uint8_t *seq_rc, uint32_t len, const uint64_t *seeds,
                                    uint32_t num_seeds, int rev, uint32_t ref_start, uint32_t ref_end, size_t *out_n) {
    /* :756-758 bucket sizes + inclusive scan, 64-bit */
    uint64_t *prefix = (uint64_t *)malloc((size_t)(num_seeds ? num_seeds : 1) * sizeof(uint64_t));
    uint64_t acc = 0;
    for (uint32_t i = 0; i < num_seeds; i++) {
        uint32_t seed = (uint32_t)(seeds[i] >> 32);
        uint32_t n = t->index[seed];
        if (seed > 0) n -= t->index[seed - 1];
        acc += n;
        prefix[i] = acc;
    }
    const uint64_t num_hits = num_seeds ? prefix[num_seeds - 1] : 0;
    /* :760-790 iteration plan (same rule as sao_iteration_plan, 64-bit; same definition of its UB zone) */
    uint32_t num_iter = 0;
    uint64_t *limit_pos = (uint64_t *)malloc(((size_t)(p->max_hits ? num_hits / p->max_hits : 0) + 2) * sizeof(uint64_t));
    if (num_hits > 0) {
        uint64_t iter_hit_limit;
        if (num_hits < p->max_hits) { num_iter = 2; iter_hit_limit = num_hits; }
        else { num_iter = (uint32_t)(num_hits / p->max_hits + 2); iter_hit_limit = p->max_hits; }
        for (uint32_t i = 0; i + 1 < num_iter; i++) {
            uint64_t pos = (uint64_t)lower_bound_u64(prefix, num_seeds, iter_hit_limit) - 1;
            limit_pos[i] = pos;
            uint64_t base = (pos == (uint64_t)-1) ? 0 : prefix[pos];
            iter_hit_limit = base + p->max_hits;
            if (iter_hit_limit > num_hits) iter_hit_limit = num_hits;
        }
        limit_pos[num_iter - 1] = num_seeds - 1;
        if (limit_pos[num_iter - 1] == limit_pos[num_iter - 2]) num_iter--;
    }
    const uint8_t *qry = rev ? seq_rc : seq; /* :805-810 */
    size_t cap = 1024, n_out = 1;
    sao_segment *out = (sao_segment *)malloc(cap * sizeof(sao_segment));
    uint64_t total_anchors = 0;
    uint32_t start_seed = 0;
    for (uint32_t it = 0; it < num_iter; it++) {
        uint32_t end_seed = (uint32_t)(limit_pos[it] + 1);
        size_t acap = 1024, an = 0;
        sao_segment *anch = (sao_segment *)malloc(acap * sizeof(sao_segment));
        for (uint32_t s = start_seed; s < end_seed; s++) {
            uint32_t seed = (uint32_t)(seeds[s] >> 32);
            uint32_t q0 = (uint32_t)(seeds[s] & 0xFFFFFFFFu) + p->seed_size;
            uint32_t end = t->index[seed];
            uint32_t start = seed > 0 ? t->index[seed - 1] : 0;
            for (uint32_t e = start; e < end; e++) {
                uint32_t r0 = t->pos[e] + p->seed_size;
                if (!(r0 >= ref_start && r0 <= ref_end)) continue; /* :239-244: score = -1, never extended */
                sao_segment seg;
                if (sao_extend_hit(p, seq, len, qry, len, r0, q0, &seg)) {
                    if (rev) seg.query_start = len - 1 - (seg.query_start + seg.len); /* :705-709 */
                    if (an == acap) { acap *= 2; anch = (sao_segment *)realloc(anch, acap * sizeof(sao_segment)); }
                    anch[an++] = seg;
                }
            }
        }
        size_t kept = sao_rm_sort_dedupe(anch, an);
        if (n_out + kept > cap) {
            while (n_out + kept > cap) cap *= 2;
            out = (sao_segment *)realloc(out, cap * sizeof(sao_segment));
        }
        memcpy(out + n_out, anch, kept * sizeof(sao_segment));
        n_out += kept;
        total_anchors += kept;
        free(anch);
        start_seed = end_seed;
    }
    /* :856-861 */
    out[0].ref_start = (uint32_t)(num_hits & 0xFFFFFFFFu);
    out[0].query_start = (uint32_t)(num_hits >> 32);
    out[0].len = (uint32_t)(total_anchors & 0xFFFFFFFFu);
    out[0].score = (int32_t)(total_anchors >> 32);
    free(prefix);
    free(limit_pos);
    *out_n = n_out;
    return out;
}

void sao_free(void *p) { free(p); }
