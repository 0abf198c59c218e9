/*
 * sa_oracle.h -- CPU restatement of SegAlign's seed-filter-extend hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg may load it, and only as the
 * checker.  The product (segalign_b200/csrc) never links or calls this code.
 *
 * Every function cites the reference file:line it restates (paths relative to the
 * gsneha26/SegAlign checkout).  Parity status: pinned against the reference's own CUDA
 * implementation run on a B200 (oracle/_ref/oracle_runner, built from the unmodified
 * reference sources by oracle/Makefile); the resulting golden vectors live in tests/golden/.
 * The reference repository itself ships no tests or golden vectors for this path.
 */
#ifndef SA_ORACLE_H
#define SA_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* src/graph.h:25-30 */
typedef struct {
    uint32_t ref_start;
    uint32_t query_start;
    uint32_t len;
    int32_t score;
} sao_segment;

#define SAO_NUC 8
#define SAO_INVALID_KMER (1u << 31) /* common/ntcoding.h:1 */

/* Seed shape state; restates the file-static shape_pos/shape_size/transition_pos of
 * common/ntcoding.cpp:6-8. */
typedef struct {
    int shape_pos[32];
    int transition_pos[32];
    int weight; /* number of care positions ("kmer_size") */
    int span;   /* length of the shape string ("seed_size") */
} sao_shape;

/* Scalars handed to InitializeProcessor (src/seed_filter.cu:830-846) + MAX_HITS. */
typedef struct {
    int sub_mat[64];
    int xdrop;
    int hspthresh;
    int noentropy;
    uint32_t seed_size; /* span */
    uint32_t max_hits;  /* MAX_HITS, src/seed_filter.cu:841 */
} sao_params;

/* CSR seed position table (common/seed_pos_table.cu:49-109).  index[k] is the inclusive end
 * offset of bucket k, i.e. what the device receives as index_table+1 (:103). */
typedef struct {
    uint32_t *index;    /* 4^weight entries */
    uint32_t index_size;
    uint32_t *pos;      /* num_pos entries; buckets in ascending position order */
    uint32_t num_pos;
} sao_table;

/* common/seed_filter_interface.cu:18-47 */
void sao_encode(const char *src, uint32_t len, uint8_t *dst);
/* src/seed_filter.cu:110-155 */
void sao_encode_rc(const char *src, uint32_t len, uint8_t *dst, uint8_t *dst_rc);
/* common/ntcoding.cpp:63-105 (host ASCII reverse complement used for minus-strand seeding) */
void sao_revcomp_ascii(char *dst, const char *src, size_t len);

/* src/main.cpp:159-180 + common/ntcoding.cpp:21-37.  seed_shape is the user string
 * ("12of19", "14of22" or a custom 0/1 pattern).  Returns the weight. */
int sao_shape_init(sao_shape *sh, const char *seed_shape);
/* common/ntcoding.cpp:43-61 */
uint32_t sao_kmer_at(const sao_shape *sh, const char *seq, size_t pos);

/* src/main.cpp:187-268.  ambiguous = "", "n", "iupac" or "x,R,P"-style triple. */
void sao_build_matrix(const char *ambiguous, int xdrop, int *sub_mat);

/* common/seed_pos_table.cu:49-109 */
int sao_table_build(sao_table *t, const sao_shape *sh, const char *ref, size_t start_addr,
                    uint32_t ref_length, uint32_t step);
void sao_table_free(sao_table *t);

/* src/seeder.cpp:48-74 (plus) / :89-109 (minus): seed words of one chunk [j0,j1).
 * seq is the ASCII block buffer (query_DRAM or query_rc_DRAM), block_start its offset.
 * out must hold (j1-j0)*(1+weight) words.  Returns the number of words. */
size_t sao_chunk_seeds(const sao_shape *sh, int transition, const char *seq, size_t block_start,
                       uint32_t j0, uint32_t j1, uint64_t *out);

/* src/seed_filter.cu:232-652: one hit.  Returns 1 if it passes (d_done=1) and fills *out. */
int sao_extend_hit(const sao_params *p, const uint8_t *ref, uint32_t ref_len, const uint8_t *qry,
                   uint32_t query_len, uint32_t r0, uint32_t q0, sao_segment *out);

/* src/seed_filter.cu:682-828: one SeedAndFilter call.  Returns a malloc'd array whose element
 * 0 is the header {0,0,len=total_anchors,score=num_hits}; *out_n = number of elements. */
sao_segment *sao_seed_and_filter(const sao_params *p, const sao_table *t, const uint8_t *ref,
                                 uint32_t ref_len, const uint8_t *qry, uint32_t query_len,
                                 const uint64_t *seeds, uint32_t num_seeds, size_t *out_n);

/* src/seed_filter.cu:718-745: iteration plan.  limit_pos must hold num_hits/max_hits+2
 * entries.  Returns num_iter (0 if num_hits == 0). */
int sao_iteration_plan(const uint32_t *prefix, uint32_t num_seeds, uint32_t max_hits,
                       uint32_t *limit_pos);

/* src/seed_filter.cu:776-782 applied to n records in place; returns the surviving count. */
size_t sao_sort_dedupe(sao_segment *a, size_t n);

/* ---- repeat-masker variant (repeat_masker_src/seed_filter.cu) ---- */
/* :138-168 reverse complement on the 8-symbol codes (what SendQueryWriteRequest() builds on the device) */
void sao_rm_revcomp_codes(const uint8_t *src, uint32_t len, uint8_t *dst);
/* :819-835 applied to n records in place; returns the surviving count */
size_t sao_rm_sort_dedupe(sao_segment *a, size_t n);
/* :724-870: one SeedAndFilter(seeds, rev, ref_start, ref_end) call of the repeat masker.  seq / seq_rc =
 * the encoded block and its device-style reverse complement.  Header (element 0): ref_start/query_start
 * = low/high word of the 64-bit hit total, len/score = low/high word of the anchor total. */
sao_segment *sao_rm_seed_and_filter(const sao_params *p, const sao_table *t, const uint8_t *seq,
                                    const uint8_t *seq_rc, uint32_t len, const uint64_t *seeds,
                                    uint32_t num_seeds, int rev, uint32_t ref_start, uint32_t ref_end, size_t *out_n);

void sao_free(void *p);

#ifdef __cplusplus
}
#endif
#endif
