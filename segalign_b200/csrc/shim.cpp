// shim.cpp -- re-exports the reference backend's own C++ symbols on top of the C ABI.
//
// Linking this TU + libsegalign_b200.so instead of the reference's seed_filter_interface.cu,
// seed_pos_table.cu and seed_filter.cu gives the unchanged host pipeline (src/main.cpp,
// src/seeder.cpp, src/segment_printer.cpp, common/ntcoding.cpp, common/DRAM.cpp) the B200
// backend.  Everything here is declared exactly as the reference declares it:
//   common/seed_filter_interface.h:3-11   g_InitializeInterface, g_SendRefWriteRequest,
//                                         g_ClearRef, g_ShutdownProcessor
//   src/seed_filter.h:4-14                g_InitializeProcessor, g_SendQueryWriteRequest,
//                                         g_SeedAndFilter, g_ClearQuery
//   common/ntcoding.h:9                   GenerateSeedPosTable
// and it imports what the reference backend imports from the host side:
//   src/store.h:7                         extern DRAM* query_DRAM   (seed_filter.cu:910)
//   common/ntcoding.cpp:6-8               shape_pos / shape_size / transition_pos (the seed
//                                         shape set by GenerateShapePos, main.cpp:180)
// The mirror declarations below keep this file compilable without the reference tree; they
// must stay layout-identical to src/graph.h:25-30 (segmentPair) and common/DRAM.h:4-12 (DRAM).
// Compile with the same libstdc++ ABI as the host code (std::vector crosses the boundary).
//
// -DSA_SHIM_REPEAT_MASKER builds the flavour segalign_repeat_masker links: the three
// symbols whose signatures differ there (repeat_masker_src/seed_filter.h:5-14) --
// g_SendQueryWriteRequest(), g_SeedAndFilter(seeds, rev, ref_start, ref_end), g_ClearQuery() -- and no
// query_DRAM import (the block is aligned against itself).
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../include/segalign_b200.h"

struct segmentPair { // src/graph.h:25-30
    uint32_t ref_start;
    uint32_t query_start;
    uint32_t len;
    int score;
};
static_assert(sizeof(segmentPair) == sizeof(sa_segment), "segmentPair must stay 16 bytes");

class DRAM { // common/DRAM.h:4-12
public:
    char *buffer;
    std::size_t size;
    std::size_t seqSize;
    std::size_t bufferPosition;
    DRAM();
    ~DRAM();
};

#ifndef SA_SHIM_REPEAT_MASKER
extern DRAM *query_DRAM;        // src/store.h:7, defined by the host (main.cpp:30)
#endif
extern int shape_pos[32];       // common/ntcoding.cpp:6
extern int shape_size;          // common/ntcoding.cpp:7 (number of care positions)
extern int transition_pos[32];  // common/ntcoding.cpp:8

typedef int (*InitializeInterface_ptr)(int num_gpu);
typedef void (*SendRefWriteRequest_ptr)(char *seq, size_t addr, uint32_t len);
typedef void (*ClearRef_ptr)();
typedef void (*ShutdownProcessor_ptr)();
typedef void (*InitializeProcessor_ptr)(bool transition, uint32_t WGA_CHUNK, uint32_t input_seed_size,
                                        int *sub_mat, int input_xdrop, int input_hspthresh,
                                        bool input_noentropy);
#ifndef SA_SHIM_REPEAT_MASKER
typedef void (*SendQueryWriteRequest_ptr)(size_t addr, uint32_t len, uint32_t buffer);
typedef std::vector<segmentPair> (*SeedAndFilter_ptr)(std::vector<uint64_t> seed_offset_vector,
                                                      bool rev, uint32_t buffer);
typedef void (*ClearQuery_ptr)(uint32_t buffer);
#else // repeat_masker_src/seed_filter.h:5-8
typedef void (*SendQueryWriteRequest_ptr)();
typedef std::vector<segmentPair> (*SeedAndFilter_ptr)(std::vector<uint64_t> seed_offset_vector, bool rev,
                                                      uint32_t ref_start, uint32_t ref_end);
typedef void (*ClearQuery_ptr)();
#endif

namespace {

// Reference error convention: print to stderr and exit(code) (common/cuda_utils.h:4-37,
// seed_filter_interface.cu:54-69; list in scripts/run_segalign:3-13).
[[noreturn]] void die(int rc) {
    fprintf(stderr, "%s\n", sa_last_error());
    switch (rc) {
        case SA_ERR_NO_GPU: exit(1);
        case SA_ERR_TOO_MANY_GPUS: exit(10);
        case SA_ERR_SET_DEVICE: exit(11);
        case SA_ERR_MALLOC: exit(12);
        case SA_ERR_MEMCPY: exit(13);
        case SA_ERR_FREE: exit(14);
        case SA_ERR_MAX_SEEDS: abort(); // the reference assert()s (seed_filter.cu:688-692)
        default: exit(15);              // no reference equivalent: state / argument / launch error
    }
}
inline void check(int rc) { if (rc < 0) die(rc); }

uint32_t g_seed_span = 0;

int InitializeInterface(int num_gpu) {
    int n = sa_initialize_interface(num_gpu);
    check(n);
    return n;
}

void InitializeProcessor(bool transition, uint32_t WGA_CHUNK, uint32_t input_seed_size, int *sub_mat,
                         int input_xdrop, int input_hspthresh, bool input_noentropy) {
    g_seed_span = input_seed_size;
    check(sa_initialize_processor(transition, WGA_CHUNK, input_seed_size, sub_mat, input_xdrop,
                                  input_hspthresh, input_noentropy));
}

void SendRefWriteRequest(char *seq, size_t addr, uint32_t len) { check(sa_send_ref(seq, addr, len)); }
void ClearRef() { check(sa_clear_ref()); }
void ShutdownProcessor() { check(sa_shutdown_processor()); }
#ifndef SA_SHIM_REPEAT_MASKER
void SendQueryWriteRequest(size_t addr, uint32_t len, uint32_t buffer) {
    check(sa_send_query(query_DRAM->buffer, addr, len, buffer)); // seed_filter.cu:910
}
void ClearQuery(uint32_t buffer) { check(sa_clear_query(buffer)); }

std::vector<segmentPair> SeedAndFilter(std::vector<uint64_t> seed_offset_vector, bool rev, uint32_t buffer) {
    sa_segment *out = nullptr;
    uint32_t n = 0;
    check(sa_seed_and_filter(seed_offset_vector.data(), (uint32_t)seed_offset_vector.size(), rev, buffer,
                             &out, &n));
#else
void SendQueryWriteRequest() { check(sa_rm_send_query()); }   // repeat_masker_src/seed_filter.cu:951-961
void ClearQuery() { check(sa_rm_clear_query()); }             // :963-971

std::vector<segmentPair> SeedAndFilter(std::vector<uint64_t> seed_offset_vector, bool rev, uint32_t ref_start,
                                       uint32_t ref_end) {                                   // :724
    sa_segment *out = nullptr;
    uint32_t n = 0;
    check(sa_rm_seed_and_filter(seed_offset_vector.data(), (uint32_t)seed_offset_vector.size(), rev, ref_start,
                                ref_end, &out, &n));
#endif
    const segmentPair *p = reinterpret_cast<const segmentPair *>(out);
    std::vector<segmentPair> result(p, p + n); // element 0 = header (hit and anchor totals)
    sa_release_result(out);
    return result;
}

} // namespace

InitializeInterface_ptr g_InitializeInterface = InitializeInterface;
SendRefWriteRequest_ptr g_SendRefWriteRequest = SendRefWriteRequest;
ClearRef_ptr g_ClearRef = ClearRef;
ShutdownProcessor_ptr g_ShutdownProcessor = ShutdownProcessor;
InitializeProcessor_ptr g_InitializeProcessor = InitializeProcessor;
SendQueryWriteRequest_ptr g_SendQueryWriteRequest = SendQueryWriteRequest;
SeedAndFilter_ptr g_SeedAndFilter = SeedAndFilter;
ClearQuery_ptr g_ClearQuery = ClearQuery;

// common/ntcoding.h:9 -- the reference builds the table on the host from ref_str; here the
// table is built on the GPU from the block SendRefWriteRequest already encoded (main.cpp:615
// precedes :621).  The seed shape is read from ntcoding.cpp's globals, as the reference's
// GetKmerIndexAtPos does.
void GenerateSeedPosTable(char *ref_str, size_t start_addr, uint32_t ref_length, uint32_t step,
                          int shape_size_arg, int kmer_size) {
    char pattern[64];
    int span = shape_size_arg;
    if (span <= 0 || span > 32) { fprintf(stderr, "seed span %d unsupported (1..32)\n", span); exit(15); }
    for (int i = 0; i < span; i++) pattern[i] = '0';
    pattern[span] = 0;
    for (int t = 0; t < shape_size; t++) pattern[shape_pos[t]] = transition_pos[t] ? 'T' : '1';
    check(sa_set_seed_shape(pattern));
    check(sa_generate_seed_pos_table(ref_str, start_addr, ref_length, step, shape_size_arg, kmer_size));
}
