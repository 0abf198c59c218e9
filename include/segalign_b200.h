/*
 * segalign_b200.h -- C ABI of the B200-native seed-filter-extend backend.
 *
 * One entry point per function of SegAlign's GPU backend boundary.  Each declaration cites the
 * reference interface it replaces (paths relative to the gsneha26/SegAlign checkout).  The
 * signatures use plain pointers and sizes only; the C++ shim that re-exports the reference's
 * own symbols (g_InitializeInterface ... g_SeedAndFilter, GenerateSeedPosTable) on top of this
 * ABI is segalign_b200/csrc/shim.cpp, see INTEGRATION.md.
 *
 * Error convention (reference: common/cuda_utils.h:4-37, seed_filter_interface.cu:54-69):
 * the reference prints to stderr and exit()s.  Here every function returns 0 on success or a
 * negative SA_ERR_* code and records a message retrievable with sa_last_error(); the shim
 * turns these back into the reference's "print + exit(code)" behaviour.
 *
 * Threading: sa_seed_and_filter* may be called concurrently from many host
 * threads; every other function is called from one control thread, and ref-level calls only
 * while no seed_and_filter call is in flight.  sa_send_query / sa_clear_query on one buffer
 * slot are safe while calls run on the other slot.
 */
#ifndef SEGALIGN_B200_H
#define SEGALIGN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* HSP / hit record -- src/graph.h:25-30 (segmentPair), 16 bytes */
typedef struct sa_segment {
    uint32_t ref_start;   /* block-relative, 0-based */
    uint32_t query_start; /* block-relative, 0-based (rev-comp coordinates when rev) */
    uint32_t len;         /* number of bases - 1 */
    int32_t score;
} sa_segment;

#define SA_BUFFER_DEPTH 2 /* src/graph.h:14 */

/* exit codes of the reference (scripts/run_segalign:3-13), negated */
#define SA_OK 0
#define SA_ERR_NO_GPU (-1)        /* exit(1)  seed_filter_interface.cu:54-57 */
#define SA_ERR_TOO_MANY_GPUS (-10) /* exit(10) seed_filter_interface.cu:66-69 */
#define SA_ERR_SET_DEVICE (-11)    /* exit(11) cuda_utils.h:4-10 */
#define SA_ERR_MALLOC (-12)        /* exit(12) cuda_utils.h:13-19 */
#define SA_ERR_MEMCPY (-13)        /* exit(13) cuda_utils.h:22-28 */
#define SA_ERR_FREE (-14)          /* exit(14) cuda_utils.h:31-37 */
#define SA_ERR_MAX_SEEDS (-20)     /* assert(num_seeds <= MAX_SEEDS) seed_filter.cu:688-692 */
#define SA_ERR_STATE (-21)         /* call order violated (no ref / table / query loaded) */
#define SA_ERR_ARG (-22)
#define SA_ERR_KERNEL (-23)        /* a launch or a stream sync failed */

const char *sa_last_error(void);

/* InitializeInterface -- common/seed_filter_interface.cu:49-80.
 * num_gpu == -1 uses every visible device.  Returns the device count (>0) or SA_ERR_*.
 * first_device lets a one-process-per-GPU launcher (torchrun) bind rank r to device r; the
 * reference always starts at device 0. */
int sa_initialize_interface(int num_gpu);
int sa_initialize_interface_at(int first_device, int num_gpu);

/* InitializeProcessor -- src/seed_filter.cu:830-897.  sub_mat is the 8x8 matrix, row = ref
 * code.  MAX_SEEDS = (transition ? 13 : 1) * wga_chunk; MAX_HITS = 4194304 * GiB(device 0). */
int sa_initialize_processor(int transition, uint32_t wga_chunk, uint32_t seed_size,
                            const int *sub_mat, int xdrop, int hspthresh, int noentropy);

/* Test/oracle knob: the reference's MAX_HITS is a non-static global (seed_filter.cu:21) that a
 * harness can lower to force the multi-iteration path; 0 restores the device-derived value. */
int sa_set_max_hits(uint32_t max_hits);
uint32_t sa_get_max_hits(void);
/* Test/measurement knob: which filter kernel the fused path launches from the next call on.
 * 3 = popcount screen + tile walk (default), 2 = tile walk only.  Both return identical results;
 * any other value restores the default.  Returns the previous value. */
int sa_set_filter_kernel(int kernel);

/* GenerateShapePos -- common/ntcoding.cpp:21-37.  pattern uses 'T'/'1' for care positions
 * ("TTT0T00TT00T0T0TTTT" for 12of19).  Returns the weight.  The reference keeps this state in
 * ntcoding.cpp's globals; the shim forwards it. */
int sa_set_seed_shape(const char *pattern);

/* SendRefWriteRequest -- common/seed_filter_interface.cu:82-101.  Uploads seq[start,start+len)
 * (ASCII) to every GPU and encodes it. */
int sa_send_ref(const char *seq, size_t start_addr, uint32_t len);

/* GenerateSeedPosTable -- common/seed_pos_table.cu:49-109.  The table is built ON THE GPU
 * from the encoded reference block already resident there (sa_send_ref must precede it, as in
 * src/main.cpp:615-621); ref_str/start_addr are accepted for signature parity and only used
 * to validate the call.  shape_size = seed span, kmer_size = weight. */
int sa_generate_seed_pos_table(const char *ref_str, size_t start_addr, uint32_t ref_length,
                               uint32_t step, int shape_size, int kmer_size);

/* ClearRef -- common/seed_filter_interface.cu:103-113 (frees ref + table on every GPU) */
int sa_clear_ref(void);

/* SendQueryWriteRequest -- src/seed_filter.cu:899-919.  The reference reads the global
 * query_DRAM->buffer + start_addr; here the base pointer is explicit. */
int sa_send_query(const char *query_base, size_t start_addr, uint32_t len, uint32_t buffer);

/* ClearQuery -- src/seed_filter.cu:921-930 */
int sa_clear_query(uint32_t buffer);

/* SeedAndFilter -- src/seed_filter.cu:682-828.
 * seeds[i] = (kmer << 32) + query_pos.  On return *out points to a library-owned array of
 * *out_count records whose element 0 is the header {0, 0, len = total_anchors,
 * score = num_hits}; release it with sa_release_result().  Returns 0 or SA_ERR_*. */
int sa_seed_and_filter(const uint64_t *seeds, uint32_t num_seeds, int rev, uint32_t buffer,
                       sa_segment **out, uint32_t *out_count);
void sa_release_result(sa_segment *out);

/* ---- repeat-masker variant (segalign_repeat_masker) -------------------------------------
 * repeat_masker_src/seed_filter.h:5-14: the same backend boundary with three signatures changed.  One
 * sequence block is aligned against itself (plus strand) and against its own reverse complement, which
 * the backend builds on the device (minus strand).  InitializeInterface / InitializeProcessor /
 * SendRefWriteRequest / GenerateSeedPosTable / ClearRef / ShutdownProcessor are the calls above. */
/* SendQueryWriteRequest() -- repeat_masker_src/seed_filter.cu:951-961.  Call after sa_send_ref. */
int sa_rm_send_query(void);
/* ClearQuery() -- repeat_masker_src/seed_filter.cu:963-971 */
int sa_rm_clear_query(void);
/* SeedAndFilter(seed_offset_vector, rev, ref_start, ref_end) -- repeat_masker_src/seed_filter.cu:724-870.
 * Seed hits whose reference anchor (position + seed span) lies outside [ref_start, ref_end] are counted
 * but not extended (:239-244); minus-strand records come back in forward coordinates (:705-709); the
 * records of an iteration are ordered and thinned by the three-sort / two-unique chain of :819-835.
 * out[0] = header {ref_start, query_start} = low / high word of the 64-bit hit total, {len, score} =
 * low / high word of the anchor total (:856-861); out[1..] = records.  Release with sa_release_result. */
int sa_rm_seed_and_filter(const uint64_t *seeds, uint32_t num_seeds, int rev, uint32_t ref_start, uint32_t ref_end,
                          sa_segment **out, uint32_t *out_count);
/* The same call with the seed words of positions [q_start, q_end) of the block (rev = 0) or of its reverse
 * complement (rev = 1) generated on the device (repeat_masker_src/seeder.cpp:69-150 builds them on the host). */
int sa_rm_seed_and_filter_range(uint32_t q_start, uint32_t q_end, int transition, int rev, uint32_t ref_start,
                                uint32_t ref_end, sa_segment **out, uint32_t *out_count, uint32_t *out_num_seeds);

/* Device-side seeding: generates the seed words of src/seeder.cpp:57-74 for
 * query positions [q_start, q_end) of the resident (fwd or rev-comp) query block on the GPU,
 * then runs the same pipeline as sa_seed_and_filter.  *out_num_seeds (optional) receives the
 * number of seed words.  A range without any valid seed returns a header-only result with
 * *out_num_seeds = 0 (the reference's seeder skips the call, seeder.cpp:76). */
int sa_seed_and_filter_range(uint32_t q_start, uint32_t q_end, int transition, int rev,
                             uint32_t buffer, sa_segment **out, uint32_t *out_count,
                             uint32_t *out_num_seeds);

/* Host seed words of one chunk -- src/seeder.cpp:57-74 + common/ntcoding.cpp:43-61, for callers
 * that keep the reference's seed-vector ABI.  seq + block_start is the ASCII block
 * (query_DRAM->buffer or query_rc_DRAM->buffer), positions [j0, j1); out must hold
 * (j1-j0)*(1+weight) words.  Uses the shape of sa_set_seed_shape.  Thread-safe.  Returns the
 * number of words written. */
size_t sa_host_chunk_seeds(const char *seq, size_t block_start, uint32_t j0, uint32_t j1,
                           int transition, uint64_t *out);

/* Segments writer -- the record formatting of src/segment_printer.cpp:72-94 (plus)
 * and :125-149 (minus): block-relative HSPs -> "name1 start1 end1 name2 start2 end2 strand score",
 * origin-one closed, the wire format LASTZ reads with --segments.  chroms: the chromosomes of the
 * block in block order (r_chr_* / q_chr_* of src/store.h:9-20; for minus != 0 the rc_q_chr_* table);
 * starts are buffer offsets like r_block_start / q_block_start.  Minus-strand records are written in
 * reverse order, as the reference does.  Host only (no CUDA). */
typedef struct sa_chrom_table {
    const char *const *names;
    const uint64_t *starts;
    const uint32_t *lens;
    uint32_t count;
} sa_chrom_table;
int sa_write_segments(const char *path, const sa_segment *hsps, uint32_t n, int minus, uint64_t r_block_start,
                      uint64_t q_block_start, const sa_chrom_table *ref_chroms, const sa_chrom_table *query_chroms);

/* Whole-genome driver -- what src/main.cpp does around the hot path, without Boost
 * or TBB: FASTA records -> '&'-separated blocks (src/main.cpp:336-415, :479-541), seeding
 * intervals (:380-393), the block schedule of the reader node (:600-741: every reference block
 * against every query block, query blocks double-buffered in the two device slots), and the
 * output of segment_printer_body (src/segment_printer.cpp:38-170) into out_dir:
 *   tmp<interval>.block<q>.r<ref block offset>.{plus,minus}.segments, ref_block<i>.name,
 *   query_block<i>.name, lastz_commands.txt (one LASTZ command per segments file; also printed to
 *   stdout when gapped != 0, as the reference does).
 * Zero / NULL fields take the reference's defaults (src/main.cpp:60-120, src/graph.h:10-12).
 * Calls sa_initialize_interface ... sa_shutdown_processor itself.  Plain-text FASTA only. */
typedef struct sa_pipeline_config {
    const char *ref_fasta, *query_fasta, *out_dir;
    const char *data_folder;   /* prefix of ref.2bit / query.2bit in the LASTZ command lines */
    const char *seed_shape;    /* "12of19" (default), "14of22" or a 0/1 pattern */
    const char *strand;        /* "plus", "minus", "both" (default) */
    const char *ambiguous;     /* "", "n", "iupac" or "x,R,P" */
    const char *output_format; /* default "maf-" */
    const char *scoring_file;  /* only forwarded to the command lines */
    const int *sub_mat;        /* 8x8 matrix; NULL = built by sa_build_matrix(ambiguous, xdrop) */
    int transition, noentropy, gapped, notrivial;
    int xdrop, ydrop, hspthresh, gappedthresh;
    uint32_t step, wga_chunk, lastz_interval;
    uint64_t seq_block_size;
    int num_gpu, num_threads;
} sa_pipeline_config;
typedef struct sa_pipeline_report {
    uint64_t ref_blocks, query_blocks, intervals, calls, seeds, hits, hsps, segment_files;
    double seconds;
    /* wall milliseconds summed over the run's block-level calls (all GPUs of the pool work concurrently) */
    double ms_ref_upload, ms_table_build, ms_query_upload;
    /* parts of `seconds`: reading + blocking the two FASTA inputs; device discovery + processor set-up; everything from the
     * first reference block's upload to the last segment file (the part the metric of bench.py corresponds to) */
    double seconds_read_input, seconds_device_init, seconds_align;
} sa_pipeline_report;
int sa_pipeline_run(const sa_pipeline_config *cfg, sa_pipeline_report *report);
/* The host-only part of sa_pipeline_run: inputs -> blocks -> intervals, block name files, and
 * report->{ref_blocks, query_blocks, intervals}.  Touches no GPU. */
int sa_pipeline_plan(const sa_pipeline_config *cfg, sa_pipeline_report *report);
/* src/main.cpp:187-268: the substitution matrix of --ambiguous / --xdrop (no --scoring file) */
int sa_build_matrix(const char *ambiguous, int xdrop, int *sub_mat);

/* ShutdownProcessor -- src/seed_filter.cu:932-940 */
int sa_shutdown_processor(void);

/* ------------------------------------------------------------------ introspection (tests, bench) */

/* Copies the device seed position table of GPU 0 back to the host.  index_out holds 4^weight
 * inclusive end offsets (what the reference uploads as index_table+1); pos_out holds num_pos
 * positions.  Pass NULL to query sizes only. */
int sa_debug_get_table(uint32_t *index_size, uint32_t *num_pos, uint32_t *index_out,
                       uint32_t *pos_out);
/* Copies the encoded (1 byte/base) ref block (which=0), query fwd (1) or query rc (2) of slot
 * `buffer` from GPU 0. */
int sa_debug_get_encoded(int which, uint32_t buffer, uint8_t *out, uint32_t len);

/* Per-phase device timings (CUDA events on the launching stream) and counters accumulated
 * since the last reset, summed over all GPUs of this process. */
typedef struct sa_stats {
    uint64_t calls, seeds, hits, survivors, anchors_pre_dedupe, hsps;
    uint64_t ext_cells;        /* ref bases scanned by the exact extension beyond 32/side */
    double ms_h2d, ms_count_scan, ms_lookup, ms_prefilter, ms_extend, ms_sort, ms_d2h;
    double ms_ref_encode, ms_table_build, ms_query_encode;
    uint64_t launches;         /* kernels launched by this library */
    uint64_t walked;           /* hits the popcount screen left to the tile walk (0 for the tile-walk-only kernels) */
    uint64_t h2d_bytes;        /* bytes sa_seed_and_filter / sa_send_query actually copied host -> device */
    uint64_t merge_calls;      /* calls that took the merge pass (more filter survivors than SEGALIGN_B200_MERGE_MIN) */
    uint64_t merge_dropped;    /* survivors of those calls dropped as provable copies of a neighbour on their diagonal */
} sa_stats;
int sa_get_stats(sa_stats *out);
int sa_reset_stats(void);
/* 1 = record per-phase CUDA events (adds stream syncs; off by default) */
int sa_set_profiling(int enabled);
/* SeedAndFilter calls served by each GPU of the pool since sa_reset_stats (the reference hands a call to
 * whichever GPU is free, src/seed_filter.cu:699-708 / :798-803): out[i] for GPU i, at most cap entries.
 * Returns the number of GPUs in the pool. */
int sa_get_gpu_calls(uint64_t *out, int cap);

const char *sa_version(void);

#ifdef __cplusplus
}
#endif
#endif
