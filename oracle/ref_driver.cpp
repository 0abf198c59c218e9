// ref_driver.cpp -- Boost-free driver around the UNMODIFIED reference backend.
//
// TEST INFRASTRUCTURE ONLY (see sa_oracle.h).  This translation unit is our own code; it is
// linked by oracle/Makefile against objects compiled straight from the reference sources
// where they lie under /root/reference (common/seed_filter_interface.cu,
// common/seed_pos_table.cu, src/seed_filter.cu, common/ntcoding.cpp, common/DRAM.cpp,
// src/seeder.cpp).  It calls the reference's g_* entry points in the order src/main.cpp does
// (main.cpp:297-298, :613-661) and builds the per-chunk seed vectors exactly like
// src/seeder.cpp:48-120, then dumps every SeedAndFilter return value so that tests can compare
// the new backend byte for byte.  With --check-seeder it additionally runs the reference's own
// seeder_body functor over the same interval and asserts that it returns the same records.
//
// usage: oracle_runner CASE_FILE OUT_FILE [--dump-table] [--check-seeder] [--repeat N]
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <chrono>

#include <cuda_runtime.h>

#include "graph.h"
#include "store.h"
#include "ntcoding.h"
#include "seed_filter_interface.h"
#include "seed_filter.h"
#ifndef NEW_BACKEND
#include "store_gpu.h"
#else
#include "segalign_b200.h" // new_runner: same driver, B200 backend behind the reference's symbols
#endif

// globals the reference objects expect from main.cpp (main.cpp:26-54)
Configuration cfg;
DRAM *ref_DRAM = nullptr;
DRAM *query_DRAM = nullptr;
DRAM *query_rc_DRAM = nullptr;
std::vector<std::string> q_chr_name;
std::vector<uint32_t> q_chr_file_name;
std::vector<size_t> q_chr_start;
std::vector<uint32_t> q_chr_len;
std::vector<std::string> rc_q_chr_name;
std::vector<uint32_t> rc_q_chr_file_name;
std::vector<size_t> rc_q_chr_start;
std::vector<uint32_t> rc_q_chr_len;
std::vector<std::string> r_chr_name;
std::vector<uint32_t> r_chr_file_name;
std::vector<size_t> r_chr_start;
std::vector<uint32_t> r_chr_len;

#ifndef NEW_BACKEND
extern int MAX_HITS;  // src/seed_filter.cu:21 (non-static global)
extern int MAX_SEEDS; // src/seed_filter.cu:20
#endif

struct CaseFile {
    std::string seed_shape;
    int32_t transition, xdrop, hspthresh, noentropy, max_hits_override, strand;
    uint32_t step, wga_chunk, lastz_interval;
    int32_t sub_mat[64];
    std::vector<char> ref, query;
};

static void die(const char *msg) {
    fprintf(stderr, "oracle_runner: %s\n", msg);
    exit(2);
}

template <typename T>
static T rd(FILE *f) {
    T v;
    if (fread(&v, sizeof(T), 1, f) != 1) die("short read");
    return v;
}

static CaseFile read_case(const char *path) {
    FILE *f = fopen(path, "rb");
    if (!f) die("cannot open case file");
    char magic[8];
    if (fread(magic, 1, 8, f) != 8 || memcmp(magic, "SACASE01", 8) != 0) die("bad case magic");
    CaseFile c;
    uint32_t n = rd<uint32_t>(f);
    c.seed_shape.resize(n);
    if (n && fread(&c.seed_shape[0], 1, n, f) != n) die("short read");
    c.transition = rd<int32_t>(f);
    c.step = rd<uint32_t>(f);
    c.xdrop = rd<int32_t>(f);
    c.hspthresh = rd<int32_t>(f);
    c.noentropy = rd<int32_t>(f);
    c.wga_chunk = rd<uint32_t>(f);
    c.lastz_interval = rd<uint32_t>(f);
    c.max_hits_override = rd<int32_t>(f);
    c.strand = rd<int32_t>(f);
    for (int i = 0; i < 64; i++) c.sub_mat[i] = rd<int32_t>(f);
    uint64_t rl = rd<uint64_t>(f);
    c.ref.resize(rl);
    if (rl && fread(c.ref.data(), 1, rl, f) != rl) die("short read");
    uint64_t ql = rd<uint64_t>(f);
    c.query.resize(ql);
    if (ql && fread(c.query.data(), 1, ql, f) != ql) die("short read");
    fclose(f);
    return c;
}

struct CallRecord {
    uint32_t rev, chunk_start, chunk_end, num_seeds;
    std::vector<segmentPair> out; // element 0 = header
};

int main(int argc, char **argv) {
    if (argc < 3) die("usage: oracle_runner CASE OUT [--dump-table] [--check-seeder] [--repeat N]");
    bool dump_table = false, check_seeder = false;
    int repeat = 1;
    for (int i = 3; i < argc; i++) {
        if (!strcmp(argv[i], "--dump-table")) dump_table = true;
        else if (!strcmp(argv[i], "--check-seeder")) check_seeder = true;
        else if (!strcmp(argv[i], "--repeat") && i + 1 < argc) repeat = atoi(argv[++i]);
    }
    CaseFile c = read_case(argv[1]);

    // main.cpp:159-180
    cfg.seed.transition = c.transition != 0;
    cfg.seed_shape = c.seed_shape;
    if (cfg.seed_shape == "12of19") {
        cfg.seed.shape = "TTT0T00TT00T0T0TTTT";
        cfg.seed.size = 19;
    } else if (cfg.seed_shape == "14of22") {
        cfg.seed.shape = "TTT0T0TT00TT00T0T0TTTT";
        cfg.seed.size = 22;
    } else {
        cfg.seed.shape = cfg.seed_shape;
        for (size_t i = 0; i < cfg.seed_shape.size(); i++)
            cfg.seed.shape[i] = (cfg.seed_shape[i] == '1') ? 'T' : '0';
        cfg.seed.size = (int)cfg.seed_shape.size();
    }
    cfg.seed.kmer_size = GenerateShapePos(cfg.seed.shape);
    cfg.step = c.step;
    cfg.xdrop = c.xdrop;
    cfg.hspthresh = c.hspthresh;
    cfg.noentropy = c.noentropy != 0;
    cfg.wga_chunk_size = c.wga_chunk;
    cfg.lastz_interval_size = c.lastz_interval;
    cfg.strand = c.strand == 0 ? "both" : (c.strand == 1 ? "plus" : "minus");
    for (int i = 0; i < 64; i++) cfg.sub_mat[i] = c.sub_mat[i];

    // main.cpp:297-302
    cfg.num_gpu = g_InitializeInterface(1);
    g_InitializeProcessor(cfg.seed.transition, cfg.wga_chunk_size, cfg.seed.size, cfg.sub_mat,
                          cfg.xdrop, cfg.hspthresh, cfg.noentropy);
#ifndef NEW_BACKEND
    int ref_max_hits = MAX_HITS;
    if (c.max_hits_override > 0) MAX_HITS = c.max_hits_override;
#else
    int ref_max_hits = (int)sa_get_max_hits();
    if (c.max_hits_override > 0) sa_set_max_hits((uint32_t)c.max_hits_override);
    int MAX_HITS = (int)sa_get_max_hits();
#endif
    ref_DRAM = new DRAM;
    query_DRAM = new DRAM;
    query_rc_DRAM = new DRAM;

    uint32_t r_len = (uint32_t)c.ref.size();
    uint32_t q_len = (uint32_t)c.query.size();
    memcpy(ref_DRAM->buffer, c.ref.data(), r_len);
    ref_DRAM->bufferPosition = r_len;
    memcpy(query_DRAM->buffer, c.query.data(), q_len);
    query_DRAM->bufferPosition = q_len;
    RevComp(query_rc_DRAM->buffer, query_DRAM->buffer, 0, 0, q_len); // main.cpp:377/:426
    query_rc_DRAM->bufferPosition = q_len;

    // main.cpp:613-621
    auto t0 = std::chrono::steady_clock::now();
    g_SendRefWriteRequest(ref_DRAM->buffer, 0, r_len);
    cudaDeviceSynchronize();
    auto t1 = std::chrono::steady_clock::now();
    GenerateSeedPosTable(ref_DRAM->buffer, 0, r_len, cfg.step, cfg.seed.size, cfg.seed.kmer_size);
    cudaDeviceSynchronize();
    auto t2 = std::chrono::steady_clock::now();
    // main.cpp:661
    g_SendQueryWriteRequest(0, q_len, 0);
    cudaDeviceSynchronize();
    auto t3 = std::chrono::steady_clock::now();

    // interval list, main.cpp:380-393 / :438-451; q_len passed to the seeder is len - seed.size
    // (main.cpp:714)
    uint32_t q_block_len = q_len - cfg.seed.size;
    uint32_t end_pos = q_len - cfg.seed.size;

    std::vector<CallRecord> calls;
    double seed_gen_s = 0, saf_s = 0;
    uint64_t total_seeds = 0, total_hits = 0, total_hsps = 0;
    for (int rep = 0; rep < repeat; rep++) {
        calls.clear();
        seed_gen_s = saf_s = 0;
        total_seeds = total_hits = total_hsps = 0;
        for (uint32_t curr = 0; curr < end_pos; curr += cfg.lastz_interval_size) {
            uint32_t q_inter_start = curr;
            uint32_t q_inter_end = std::min(end_pos, curr + cfg.lastz_interval_size);
            uint32_t rc_q_inter_start = q_block_len - q_inter_end;   // seeder.cpp:33
            uint32_t rc_q_inter_end = q_block_len - q_inter_start;   // seeder.cpp:34
            for (int rev = 0; rev < 2; rev++) {
                if (rev == 0 && !(cfg.strand == "plus" || cfg.strand == "both")) continue;
                if (rev == 1 && !(cfg.strand == "minus" || cfg.strand == "both")) continue;
                uint32_t lo = rev ? rc_q_inter_start : q_inter_start;
                uint32_t hi = rev ? rc_q_inter_end : q_inter_end;
                char *buf = rev ? query_rc_DRAM->buffer : query_DRAM->buffer;
                for (uint32_t i = lo; i < hi; i += cfg.wga_chunk_size) { // seeder.cpp:48 / :89
                    uint32_t e = std::min(i + cfg.wga_chunk_size, hi);
                    auto s0 = std::chrono::steady_clock::now();
                    std::vector<uint64_t> seed_offset_vector;
                    for (uint32_t j = i; j < e; j++) { // seeder.cpp:57-74
                        uint64_t kmer_index = GetKmerIndexAtPos(buf, 0 + j, cfg.seed.size);
                        if (kmer_index != ((uint32_t)1 << 31)) {
                            uint64_t seed_offset = (kmer_index << 32) + j;
                            seed_offset_vector.push_back(seed_offset);
                            if (cfg.seed.transition) {
                                for (int t = 0; t < cfg.seed.kmer_size; t++) {
                                    if (IsTransitionAtPos(t) == 1) {
                                        uint64_t tr = (kmer_index ^ (TRANSITION_MASK << (2 * t)));
                                        seed_offset_vector.push_back((tr << 32) + j);
                                    }
                                }
                            }
                        }
                    }
                    auto s1 = std::chrono::steady_clock::now();
                    seed_gen_s += std::chrono::duration<double>(s1 - s0).count();
                    if (seed_offset_vector.size() > 0) { // seeder.cpp:76
                        CallRecord cr;
                        cr.rev = rev;
                        cr.chunk_start = i;
                        cr.chunk_end = e;
                        cr.num_seeds = (uint32_t)seed_offset_vector.size();
                        cr.out = g_SeedAndFilter(seed_offset_vector, rev != 0, 0);
                        auto s2 = std::chrono::steady_clock::now();
                        saf_s += std::chrono::duration<double>(s2 - s1).count();
                        total_seeds += cr.num_seeds;
                        total_hits += (uint32_t)cr.out[0].score;
                        total_hsps += cr.out.size() - 1;
                        calls.push_back(std::move(cr));
                    }
                }
            }
        }
    }

    if (check_seeder) {
        // run the reference's own seeder_body (src/seeder.cpp:12-127) over the same intervals
        std::vector<segmentPair> fw_all, rc_all, fw_mine, rc_mine;
        uint32_t ninv = 0;
        for (uint32_t curr = 0; curr < end_pos; curr += cfg.lastz_interval_size) {
            seq_block blk;
            blk.r_index = 1; blk.q_index = 0; blk.r_start = 0; blk.q_start = 0;
            blk.r_len = r_len; blk.q_len = q_block_len;
            seed_interval inter;
            inter.start = curr;
            inter.end = std::min(end_pos, curr + cfg.lastz_interval_size);
            inter.num_invoked = ++ninv; inter.num_intervals = 0; inter.buffer = 0;
            seeder_body body;
            printer_input pi = body(seeder_input(seeder_payload(blk, inter), 0));
            auto &pl = get<0>(pi);
            auto &fw = get<1>(pl);
            auto &rc = get<2>(pl);
            fw_all.insert(fw_all.end(), fw.begin(), fw.end());
            rc_all.insert(rc_all.end(), rc.begin(), rc.end());
        }
        for (auto &cr : calls) {
            auto &dst = cr.rev ? rc_mine : fw_mine;
            dst.insert(dst.end(), cr.out.begin() + 1, cr.out.end());
        }
        bool ok = fw_all.size() == fw_mine.size() && rc_all.size() == rc_mine.size() &&
                  (fw_all.empty() || !memcmp(fw_all.data(), fw_mine.data(), fw_all.size() * 16)) &&
                  (rc_all.empty() || !memcmp(rc_all.data(), rc_mine.data(), rc_all.size() * 16));
        fprintf(stderr, "check-seeder: %s (fw %zu/%zu rc %zu/%zu)\n", ok ? "OK" : "MISMATCH",
                fw_all.size(), fw_mine.size(), rc_all.size(), rc_mine.size());
        if (!ok) return 3;
    }

    FILE *o = fopen(argv[2], "wb");
    if (!o) die("cannot open output");
    fwrite("SAOUT001", 1, 8, o);
    uint32_t ncalls = (uint32_t)calls.size();
    fwrite(&ncalls, 4, 1, o);
    for (auto &cr : calls) {
        uint32_t nseg = (uint32_t)cr.out.size() - 1;
        uint32_t hdr[7] = {cr.rev, cr.chunk_start, cr.chunk_end, cr.num_seeds, nseg,
                           cr.out[0].len, (uint32_t)cr.out[0].score};
        fwrite(hdr, 4, 7, o);
        if (nseg) fwrite(cr.out.data() + 1, 16, nseg, o);
    }
    double times[5] = {std::chrono::duration<double>(t1 - t0).count(),
                       std::chrono::duration<double>(t2 - t1).count(),
                       std::chrono::duration<double>(t3 - t2).count(), seed_gen_s, saf_s};
    fwrite(times, 8, 5, o);
    uint64_t counters[4] = {total_seeds, total_hits, total_hsps, (uint64_t)ref_max_hits};
    fwrite(counters, 8, 4, o);
    uint32_t has_table = dump_table ? 1 : 0;
    fwrite(&has_table, 4, 1, o);
    if (dump_table) {
        uint32_t index_size = (uint32_t)1 << (2 * cfg.seed.kmer_size);
        std::vector<uint32_t> idx(index_size);
#ifndef NEW_BACKEND
        cudaMemcpy(idx.data(), d_index_table[0], (size_t)index_size * 4, cudaMemcpyDeviceToHost);
        uint32_t num_pos = idx[index_size - 1];
        std::vector<uint32_t> pos(num_pos ? num_pos : 1);
        cudaMemcpy(pos.data(), d_pos_table[0], (size_t)num_pos * 4, cudaMemcpyDeviceToHost);
#else
        uint32_t num_pos = 0;
        sa_debug_get_table(&index_size, &num_pos, nullptr, nullptr);
        std::vector<uint32_t> pos(num_pos ? num_pos : 1);
        sa_debug_get_table(&index_size, &num_pos, idx.data(), pos.data());
#endif
        fwrite(&index_size, 4, 1, o);
        fwrite(&num_pos, 4, 1, o);
        fwrite(idx.data(), 4, index_size, o);
        fwrite(pos.data(), 4, num_pos, o);
    }
    fclose(o);
    fprintf(stderr,
            "oracle_runner: calls=%u seeds=%lu hits=%lu hsps=%lu  ref_upload=%.3fs table=%.3fs "
            "query_upload=%.3fs seedgen=%.3fs seed_and_filter=%.3fs  MAX_HITS(ref)=%d used=%d\n",
            ncalls, (unsigned long)total_seeds, (unsigned long)total_hits,
            (unsigned long)total_hsps, times[0], times[1], times[2], times[3], times[4],
            ref_max_hits, MAX_HITS);
    // g_ShutdownProcessor() calls cudaDeviceReset(); results are already on disk
    g_ShutdownProcessor();
    return 0;
}
