// segalign_main.cpp -- command line front end of the whole-genome driver.
//
// Accepts the options of the reference's `segalign` (src/main.cpp:60-148) that concern the
// seed-filter-extend stage and the LASTZ hand-off, without Boost.program_options:
//   segalign_b200 target.fa query.fa [data_folder] [--strand=..] [--ambiguous=..] [--step=N]
//       [--xdrop=N] [--ydrop=N] [--hspthresh=N] [--gappedthresh=N] [--notransition] [--nogapped]
//       [--notrivial] [--noentropy] [--seed=12of19|14of22|pattern] [--wga_chunk_size=N]
//       [--lastz_interval_size=N] [--seq_block_size=N] [--num_gpu=N] [--num_threads=N]
//       [--format=F] [--scoring=FILE (forwarded to LASTZ only)] [--out_dir=DIR]
//       [--output=FILE] [--markend] [--debug] [--version]
// Option names are the reference's (src/main.cpp:62-110; --wga_chunk / --lastz_interval remain as
// aliases); `--key value` is accepted like `--key=value`.  --output names the file the gapped stage's
// wrapper script collects into (scripts/run_segalign): accepted and unused here, as in the reference
// binary; --markend and --debug likewise (the marker line is written by scripts/run_segalign:50, not by
// the binary).
// Inputs are plain FASTA (the reference also wants ref.2bit/query.2bit in data_folder for LASTZ;
// they only appear in the printed command lines).  Everything else is sa_pipeline_run.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/segalign_b200.h"

int main(int argc, char **argv) {
    sa_pipeline_config cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.transition = 1; cfg.gapped = 1;
    cfg.xdrop = 910; cfg.ydrop = 9430; cfg.hspthresh = 3000; cfg.step = 1; cfg.num_gpu = -1;
    std::vector<std::string> pos;
    std::string strand = "both", ambiguous, seed = "12of19", format = "maf-", scoring, out_dir = ".";
    int i = 1;
    auto val = [&](const char *a, const char *key, std::string &out) {
        const size_t n = strlen(key);
        if (strncmp(a, key, n) != 0) return false;
        if (a[n] == '=') { out = a + n + 1; return true; }
        if (a[n] == 0 && i + 1 < argc) { out = argv[++i]; return true; } // "--key value"
        return false;
    };
    bool markend = false, debug = false;
    std::string output;
    for (; i < argc; i++) {
        const char *a = argv[i];
        std::string v;
        if (a[0] != '-') pos.push_back(a);
        else if (!strcmp(a, "--notransition")) cfg.transition = 0;
        else if (!strcmp(a, "--nogapped")) cfg.gapped = 0;
        else if (!strcmp(a, "--notrivial")) cfg.notrivial = 1;
        else if (!strcmp(a, "--noentropy")) cfg.noentropy = 1;
        else if (!strcmp(a, "--markend")) markend = true;
        else if (!strcmp(a, "--debug")) debug = true;
        else if (!strcmp(a, "--version")) { fprintf(stderr, "SegAlign (B200 backend) %s\n", sa_version()); return 0; }
        else if (val(a, "--output", output)) {}
        else if (val(a, "--strand", strand) || val(a, "--ambiguous", ambiguous) || val(a, "--seed", seed) ||
                 val(a, "--format", format) || val(a, "--scoring", scoring) || val(a, "--out_dir", out_dir)) {}
        else if (val(a, "--step", v)) cfg.step = (uint32_t)atoi(v.c_str());
        else if (val(a, "--xdrop", v)) cfg.xdrop = atoi(v.c_str());
        else if (val(a, "--ydrop", v)) cfg.ydrop = atoi(v.c_str());
        else if (val(a, "--hspthresh", v)) cfg.hspthresh = atoi(v.c_str());
        else if (val(a, "--gappedthresh", v)) cfg.gappedthresh = atoi(v.c_str());
        else if (val(a, "--wga_chunk_size", v) || val(a, "--wga_chunk", v)) cfg.wga_chunk = (uint32_t)atoi(v.c_str());
        else if (val(a, "--lastz_interval_size", v) || val(a, "--lastz_interval", v)) cfg.lastz_interval = (uint32_t)atoi(v.c_str());
        else if (val(a, "--seq_block_size", v)) cfg.seq_block_size = strtoull(v.c_str(), nullptr, 10);
        else if (val(a, "--num_gpu", v)) cfg.num_gpu = atoi(v.c_str());
        else if (val(a, "--num_threads", v)) cfg.num_threads = atoi(v.c_str());
        else { fprintf(stderr, "unknown option %s\n", a); return 1; }
    }
    if (pos.size() < 2) {
        fprintf(stderr, "usage: %s target.fa query.fa [data_folder] [options]  (see the header of segalign_main.cpp)\n", argv[0]);
        return 1;
    }
    std::string data_folder = pos.size() > 2 ? pos[2] : "";
    if (!data_folder.empty() && data_folder.back() != '/') data_folder += '/';
    cfg.ref_fasta = pos[0].c_str(); cfg.query_fasta = pos[1].c_str(); cfg.out_dir = out_dir.c_str();
    cfg.data_folder = data_folder.c_str(); cfg.seed_shape = seed.c_str(); cfg.strand = strand.c_str();
    cfg.ambiguous = ambiguous.c_str(); cfg.output_format = format.c_str(); cfg.scoring_file = scoring.c_str();
    sa_pipeline_report rep;
    const int rc = sa_pipeline_run(&cfg, &rep);
    if (rc < 0) {
        fprintf(stderr, "%s\n", sa_last_error());
        return -rc; // the reference's exit codes (scripts/run_segalign:3-13)
    }
    fprintf(stderr, "ref blocks %llu, query blocks %llu, intervals %llu, SeedAndFilter calls %llu, seeds %llu, hits %llu, "
                    "HSPs %llu, segment files %llu, %.2f s (reference blocks: upload + encode %.1f ms, seed position tables %.1f ms; "
                    "query blocks: upload + encode %.1f ms); reading the inputs %.2f s, device set-up %.2f s, alignment %.2f s\n",
            (unsigned long long)rep.ref_blocks, (unsigned long long)rep.query_blocks, (unsigned long long)rep.intervals,
            (unsigned long long)rep.calls, (unsigned long long)rep.seeds, (unsigned long long)rep.hits,
            (unsigned long long)rep.hsps, (unsigned long long)rep.segment_files, rep.seconds, rep.ms_ref_upload,
            rep.ms_table_build, rep.ms_query_upload, rep.seconds_read_input, rep.seconds_device_init, rep.seconds_align);
    (void)debug; (void)markend;
    return 0;
}
