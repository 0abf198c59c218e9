// rm_driver.cpp -- Boost-free driver around the UNMODIFIED repeat-masker backend.
//
// TEST INFRASTRUCTURE ONLY (see sa_oracle.h).  Our own code; linked by oracle/Makefile against objects
// compiled straight from the reference sources where they lie under /root/reference
// (repeat_masker_src/seed_filter.cu, common/seed_filter_interface.cu, common/seed_pos_table.cu,
// common/ntcoding.cpp, common/DRAM.cpp).  It calls the reference's g_* entry points in the order
// repeat_masker_src/main.cpp does (:256-257, :494-505), builds the interval list with its reference
// windows like main.cpp:323-420 (one sequence block) and the per-chunk seed vectors of both strands
// like repeat_masker_src/seeder.cpp:69-150, and dumps every SeedAndFilter return value.
// With NEW_BACKEND the same driver runs on segalign_b200's shim (rm_new_runner: the drop-in check).
//
// usage: rm_oracle_runner CASE_FILE OUT_FILE [--neigh-prop P]
//   CASE_FILE = SACASE01 (oracle/ref_driver.cpp); only its reference sequence is used (self-alignment).
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "graph.h"
#include "store.h"
#include "ntcoding.h"
#include "seed_filter_interface.h"
#include "seed_filter.h"
#ifdef NEW_BACKEND
#include "segalign_b200.h"
#endif

// globals the reference objects expect from repeat_masker_src/main.cpp (:26-33)
Configuration cfg;
DRAM *seq_DRAM = nullptr;
DRAM *seq_rc_DRAM = nullptr;
std::vector<std::string> chr_name;
std::vector<size_t> chr_start;
std::vector<uint32_t> chr_len;

#ifndef NEW_BACKEND
extern int MAX_HITS;  // repeat_masker_src/seed_filter.cu (non-static global)
#endif

static void die(const char *msg) { fprintf(stderr, "rm_oracle_runner: %s\n", msg); exit(2); }
template <typename T> static T rd(FILE *f) { T v; if (fread(&v, sizeof(T), 1, f) != 1) die("short read"); return v; }

struct CallRecord {
    uint32_t rev, chunk_start, chunk_end, num_seeds, ref_start, ref_end;
    std::vector<segmentPair> out; // element 0 = header
};

int main(int argc, char **argv) {
    if (argc < 3) die("usage: rm_oracle_runner CASE OUT [--neigh-prop P]");
    float prop = 1.0f;
    for (int i = 3; i < argc; i++)
        if (!strcmp(argv[i], "--neigh-prop") && i + 1 < argc) prop = (float)atof(argv[++i]);
    FILE *f = fopen(argv[1], "rb");
    if (!f) die("cannot open case file");
    char magic[8];
    if (fread(magic, 1, 8, f) != 8 || memcmp(magic, "SACASE01", 8) != 0) die("bad case magic");
    uint32_t n = rd<uint32_t>(f);
    std::string seed_shape(n, 0);
    if (n && fread(&seed_shape[0], 1, n, f) != n) die("short read");
    const int32_t transition = rd<int32_t>(f);
    cfg.step = rd<uint32_t>(f);
    cfg.xdrop = rd<int32_t>(f);
    cfg.hspthresh = rd<int32_t>(f);
    cfg.noentropy = rd<int32_t>(f) != 0;
    cfg.wga_chunk_size = rd<uint32_t>(f);
    cfg.lastz_interval_size = rd<uint32_t>(f);
    const int32_t max_hits_override = rd<int32_t>(f);
    const int32_t strand = rd<int32_t>(f);
    for (int i = 0; i < 64; i++) cfg.sub_mat[i] = rd<int32_t>(f);
    const uint64_t rl = rd<uint64_t>(f);
    std::vector<char> seq(rl);
    if (rl && fread(seq.data(), 1, rl, f) != rl) die("short read");
    fclose(f);

    // repeat_masker_src/main.cpp:130-150 (seed), :256-257
    cfg.seed.transition = transition != 0;
    cfg.seed_shape = seed_shape;
    if (seed_shape == "12of19") { cfg.seed.shape = "TTT0T00TT00T0T0TTTT"; cfg.seed.size = 19; }
    else if (seed_shape == "14of22") { cfg.seed.shape = "TTT0T0TT00TT00T0T0TTTT"; cfg.seed.size = 22; }
    else {
        cfg.seed.shape = seed_shape;
        for (size_t i = 0; i < seed_shape.size(); i++) cfg.seed.shape[i] = seed_shape[i] == '1' ? 'T' : '0';
        cfg.seed.size = (int)seed_shape.size();
    }
    cfg.seed.kmer_size = GenerateShapePos(cfg.seed.shape);
    cfg.strand = strand == 0 ? "both" : (strand == 1 ? "plus" : "minus");
    cfg.prop_neigh_interval = prop;
    cfg.seq_block_size = DEFAULT_SEQ_BLOCK_SIZE;
    cfg.num_gpu = g_InitializeInterface(1);
    g_InitializeProcessor(cfg.seed.transition, cfg.wga_chunk_size, cfg.seed.size, cfg.sub_mat, cfg.xdrop, cfg.hspthresh, cfg.noentropy);
#ifndef NEW_BACKEND
    const int ref_max_hits = MAX_HITS;
    if (max_hits_override > 0) MAX_HITS = max_hits_override;
#else
    const int ref_max_hits = (int)sa_get_max_hits();
    if (max_hits_override > 0) sa_set_max_hits((uint32_t)max_hits_override);
#endif

    // main.cpp:264-311: the sequence and its host reverse complement
    seq_DRAM = new DRAM;
    seq_rc_DRAM = new DRAM;
    memcpy(seq_DRAM->buffer, seq.data(), rl);
    seq_DRAM->bufferPosition = rl;
    cfg.seq_len = rl;
    RevComp(seq_rc_DRAM->buffer, seq_DRAM->buffer, seq_rc_DRAM->bufferPosition, 0, cfg.seq_len);

    // main.cpp:323-331, :348-420 for one block that holds the whole sequence
    const uint32_t total_query_intervals = (uint32_t)ceil((float)cfg.seq_len / cfg.lastz_interval_size);
    cfg.num_neigh_interval = (uint32_t)ceil((float)cfg.prop_neigh_interval * total_query_intervals);
    const uint32_t left_intervals = (uint32_t)ceil((float)(cfg.num_neigh_interval - 1) / 2);
    const uint32_t right_intervals = cfg.num_neigh_interval - 1 - left_intervals;
    const uint32_t left_overlap = left_intervals * cfg.lastz_interval_size;
    const uint32_t right_overlap = right_intervals * cfg.lastz_interval_size;
    const uint32_t max_interval_seq_len = left_overlap + cfg.lastz_interval_size + right_overlap;
    const size_t block_start = 0;
    const uint32_t block_len = (uint32_t)cfg.seq_len;
    std::vector<seed_interval> intervals;
    {
        uint32_t start_pos = 0, end_pos = block_len - cfg.seed.size; // seq_block_len < seq_block_size branch (:366-367)
        while (start_pos < end_pos) {
            seed_interval inter;
            inter.start = start_pos;
            inter.end = std::min(end_pos, start_pos + cfg.lastz_interval_size);
            const bool left_limit = inter.start < left_overlap;
            const bool right_limit = (inter.end + right_overlap) > block_len;
            if (left_limit) {
                inter.ref_start = 0;
                inter.ref_end = right_limit ? block_len : (max_interval_seq_len > block_len ? block_len : max_interval_seq_len);
            } else if (right_limit) {
                inter.ref_end = block_len;
                inter.ref_start = block_len < max_interval_seq_len ? 0 : block_len - max_interval_seq_len;
            } else {
                inter.ref_start = inter.start - left_overlap;
                inter.ref_end = inter.end + right_overlap;
            }
            inter.num_invoked = inter.num_intervals = 0;
            intervals.push_back(inter);
            start_pos += cfg.lastz_interval_size;
        }
    }

    // main.cpp:494-505
    auto t0 = std::chrono::steady_clock::now();
    g_SendRefWriteRequest(seq_DRAM->buffer, block_start, block_len);
    g_SendQueryWriteRequest();
    cudaDeviceSynchronize();
    auto t1 = std::chrono::steady_clock::now();
    GenerateSeedPosTable(seq_DRAM->buffer, block_start, block_len, cfg.step, cfg.seed.size, cfg.seed.kmer_size);
    cudaDeviceSynchronize();
    auto t2 = std::chrono::steady_clock::now();

    std::vector<CallRecord> calls;
    double saf_s = 0;
    uint64_t total_seeds = 0, total_hits = 0, total_hsps = 0;
    const size_t rc_block_start = cfg.seq_len - 1 - block_start - (block_len - 1); // seeder.cpp:44
    for (const seed_interval &inter : intervals) {
        const uint32_t start_pos = inter.start, end_pos = inter.end;
        const uint32_t end_pos_rc = block_len - 1 - start_pos; // seeder.cpp:43
        for (uint32_t i = start_pos; i < end_pos; i += cfg.wga_chunk_size) { // seeder.cpp:69-150
            int32_t start = (int32_t)i;
            int32_t end = (int32_t)std::min((uint32_t)start + cfg.wga_chunk_size, end_pos);
            for (int rev = 0; rev < 2; rev++) {
                if (rev == 0 && !(cfg.strand == "plus" || cfg.strand == "both")) continue;
                if (rev == 1 && !(cfg.strand == "minus" || cfg.strand == "both")) continue;
                if (rev == 1) { // seeder.cpp:110-111 (reuses the plus chunk's `end`)
                    start = (int32_t)(block_len - 1 - (uint32_t)end);
                    end = (int32_t)std::min((uint32_t)start + cfg.wga_chunk_size, end_pos_rc);
                }
                char *buf = rev ? seq_rc_DRAM->buffer : seq_DRAM->buffer;
                const size_t base = rev ? rc_block_start : block_start;
                std::vector<uint64_t> seed_offset_vector;
                for (uint32_t j = (uint32_t)start; j < (uint32_t)end; j++) {
                    const uint64_t kmer_index = GetKmerIndexAtPos(buf, base + j, cfg.seed.size);
                    if (kmer_index != ((uint32_t)1 << 31)) {
                        seed_offset_vector.push_back((kmer_index << 32) + j);
                        if (cfg.seed.transition)
                            for (int t = 0; t < cfg.seed.kmer_size; t++)
                                if (IsTransitionAtPos(t) == 1)
                                    seed_offset_vector.push_back(((kmer_index ^ (TRANSITION_MASK << (2 * t))) << 32) + j);
                    }
                }
                if (seed_offset_vector.size() > 0) {
                    CallRecord cr;
                    cr.rev = rev; cr.chunk_start = (uint32_t)start; cr.chunk_end = (uint32_t)end;
                    cr.num_seeds = (uint32_t)seed_offset_vector.size();
                    cr.ref_start = inter.ref_start; cr.ref_end = inter.ref_end;
                    auto s1 = std::chrono::steady_clock::now();
                    cr.out = g_SeedAndFilter(seed_offset_vector, rev != 0, inter.ref_start, inter.ref_end);
                    saf_s += std::chrono::duration<double>(std::chrono::steady_clock::now() - s1).count();
                    total_seeds += cr.num_seeds;
                    total_hits += ((uint64_t)cr.out[0].query_start << 32) + cr.out[0].ref_start; // seeder.cpp:96
                    total_hsps += cr.out.size() - 1;
                    calls.push_back(std::move(cr));
                }
            }
        }
    }

    FILE *o = fopen(argv[2], "wb");
    if (!o) die("cannot open output");
    fwrite("SARMO001", 1, 8, o);
    const uint32_t ncalls = (uint32_t)calls.size();
    fwrite(&ncalls, 4, 1, o);
    for (auto &cr : calls) {
        const uint32_t nseg = (uint32_t)cr.out.size() - 1;
        const uint32_t hdr[11] = {cr.rev, cr.chunk_start, cr.chunk_end, cr.num_seeds, cr.ref_start, cr.ref_end, nseg,
                                  cr.out[0].ref_start, cr.out[0].query_start, cr.out[0].len, (uint32_t)cr.out[0].score};
        fwrite(hdr, 4, 11, o);
        if (nseg) fwrite(cr.out.data() + 1, 16, nseg, o);
    }
    const double times[3] = {std::chrono::duration<double>(t1 - t0).count(), std::chrono::duration<double>(t2 - t1).count(), saf_s};
    fwrite(times, 8, 3, o);
    const uint64_t counters[4] = {total_seeds, total_hits, total_hsps, (uint64_t)ref_max_hits};
    fwrite(counters, 8, 4, o);
    fclose(o);
    fprintf(stderr, "rm_oracle_runner: intervals=%zu calls=%u seeds=%lu hits=%lu hsps=%lu upload=%.3fs table=%.3fs seed_and_filter=%.3fs\n",
            intervals.size(), ncalls, (unsigned long)total_seeds, (unsigned long)total_hits, (unsigned long)total_hsps,
            times[0], times[1], times[2]);
    g_ShutdownProcessor();
    return 0;
}
