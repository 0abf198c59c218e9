// segprint_driver.cpp -- runs the UNMODIFIED reference segment printer on HSP lists read from a file.
//
// TEST INFRASTRUCTURE ONLY (see sa_oracle.h).  Our own code; linked by oracle/Makefile against
// src/segment_printer.cpp compiled where it lies under /root/reference.  It fills the globals that
// src/main.cpp keeps (chromosome tables, cfg), wraps one (block, interval, fw_hsps, rc_hsps) tuple
// into the reference's printer_input, pushes it through a real tbb::flow printer_node holding the
// reference's segment_printer_body (src/segment_printer.cpp:11-173) and lets it write its
// tmp*.segments files into the current directory and its LASTZ command lines to stdout.  CPU only.
//
// usage: segprint_runner INPUT_FILE OUT_DIR      (chdir(OUT_DIR); stdout = command lines)
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <unistd.h>

#include "graph.h"
#include "store.h"

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

static void die(const char *m) { fprintf(stderr, "segprint_runner: %s\n", m); exit(2); }
template <typename T> static T rd(FILE *f) { T v; if (fread(&v, sizeof(T), 1, f) != 1) die("short read"); return v; }
static std::string rds(FILE *f) {
    uint32_t n = rd<uint32_t>(f);
    std::string s(n, 0);
    if (n && fread(&s[0], 1, n, f) != n) die("short read");
    return s;
}
static void rd_table(FILE *f, std::vector<std::string> &names, std::vector<uint32_t> &file_names,
                     std::vector<size_t> &starts, std::vector<uint32_t> &lens) {
    uint32_t n = rd<uint32_t>(f);
    for (uint32_t i = 0; i < n; i++) {
        names.push_back(rds(f));
        starts.push_back((size_t)rd<uint64_t>(f));
        lens.push_back(rd<uint32_t>(f));
        file_names.push_back(i);
    }
}
static hsp_output rd_hsps(FILE *f) {
    uint32_t n = rd<uint32_t>(f);
    hsp_output v(n);
    if (n && fread(v.data(), sizeof(segmentPair), n, f) != n) die("short read");
    return v;
}

int main(int argc, char **argv) {
    if (argc < 3) die("usage: segprint_runner INPUT OUT_DIR");
    FILE *f = fopen(argv[1], "rb");
    if (!f) die("cannot open input");
    char magic[8];
    if (fread(magic, 1, 8, f) != 8 || memcmp(magic, "SASEG001", 8) != 0) die("bad magic");
    cfg.data_folder = rds(f);
    cfg.output_format = rds(f);
    cfg.ambiguous = rds(f);
    cfg.scoring_file = rds(f);
    cfg.gapped = rd<int32_t>(f) != 0;
    cfg.ydrop = rd<int32_t>(f);
    cfg.gappedthresh = rd<int32_t>(f);
    cfg.notrivial = rd<int32_t>(f) != 0;
    rd_table(f, r_chr_name, r_chr_file_name, r_chr_start, r_chr_len);
    rd_table(f, q_chr_name, q_chr_file_name, q_chr_start, q_chr_len);
    rd_table(f, rc_q_chr_name, rc_q_chr_file_name, rc_q_chr_start, rc_q_chr_len);
    seq_block blk;
    blk.r_index = rd<int32_t>(f);   // main.cpp hands over ref block index + 1
    blk.q_index = rd<int32_t>(f);
    blk.r_start = (size_t)rd<uint64_t>(f);
    blk.q_start = (size_t)rd<uint64_t>(f);
    blk.r_len = rd<uint32_t>(f);
    blk.q_len = rd<uint32_t>(f);    // block length - seed size (main.cpp:714)
    seed_interval inter;
    inter.start = rd<uint32_t>(f);
    inter.end = rd<uint32_t>(f);
    inter.num_invoked = rd<uint32_t>(f);
    inter.num_intervals = 0;
    inter.buffer = 0;
    hsp_output fw = rd_hsps(f), rc = rd_hsps(f);
    fclose(f);
    if (chdir(argv[2]) != 0) die("cannot chdir to OUT_DIR");

    tbb::flow::graph g;
    printer_node node(g, tbb::flow::unlimited, segment_printer_body());
    printer_payload payload(seeder_payload(blk, inter), fw, rc);
    node.try_put(printer_input(payload, (size_t)0));
    g.wait_for_all();
    fflush(stdout);
    return 0;
}
