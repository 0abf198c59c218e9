// sa_backend.cu -- host side of the B200 backend + the C ABI of include/segalign_b200.h.
//
// Mirrors the reference backend's host logic (common/seed_filter_interface.cu:49-113,
// common/seed_pos_table.cu:33-109, src/seed_filter.cu:682-940) with a different execution
// model: per-GPU contexts, several independent workspaces (stream + buffers) per GPU so that
// concurrent SeedAndFilter calls from the host pipeline overlap, table build on the device,
// and buffers that grow on demand instead of a 36-byte-per-MAX_HITS up-front allocation.
// There is NO CPU fallback: every entry point fails with an SA_ERR_* code if CUDA fails.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <fstream>
#include <thread>
#include <condition_variable>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <mutex>
#include <string>
#include <vector>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "kernels_encode.cuh"
#include "kernels_extend.cuh"
#include "kernels_extend_wide.cuh"
#include "kernels_filter.cuh"
#include "kernels_lookup.cuh"
#include "kernels_merge.cuh"
#include "kernels_screen.cuh"
#include "kernels_sort.cuh"
#include "sa_common.cuh"

using namespace sa;

namespace {

thread_local std::string g_err;

int fail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CU(expr, code)                                                                      \
    do {                                                                                    \
        cudaError_t _e = (expr);                                                            \
        if (_e != cudaSuccess)                                                              \
            return fail((code), "%s failed with error \" %s \" (%s:%d)", #expr,             \
                        cudaGetErrorString(_e), __FILE__, __LINE__);                        \
    } while (0)

#define TRY(expr)                  \
    do {                           \
        int _r = (expr);           \
        if (_r != SA_OK) return _r; \
    } while (0)

int g_sm_count = 148; // multiProcessorCount of device 0, read in sa_initialize_interface_at
inline int grid_for(size_t n, int block, int per_sm = 8) {
    size_t want = (n + block - 1) / block;
    size_t cap = (size_t)g_sm_count * per_sm;
    return (int)std::max<size_t>(1, std::min(want, cap));
}

template <typename T>
int ensure(T *&ptr, size_t &cap, size_t need, const char *tag, size_t slack_num = 5,
           size_t slack_den = 4) {
    if (need <= cap && ptr) return SA_OK;
    if (ptr) CU(cudaFree(ptr), SA_ERR_FREE);
    ptr = nullptr;
    size_t n = std::max<size_t>(need * slack_num / slack_den, 1024);
    cudaError_t e = cudaMalloc((void **)&ptr, n * sizeof(T));
    if (e != cudaSuccess) {
        cap = 0;
        return fail(SA_ERR_MALLOC, "cudaMalloc of %zu bytes for %s failed with error \" %s \"",
                    n * sizeof(T), tag, cudaGetErrorString(e));
    }
    cap = n;
    return SA_OK;
}

constexpr uint32_t kDedupSlots = 1u << 16; // exact-duplicate table of the exact stage (kernels_extend.cuh)

struct Workspace {
    int gpu = 0; // index into G.gpus
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[10] = {}; // indexed by Phase
    uint64_t *d_seeds = nullptr; size_t seeds_cap = 0;
    uint32_t *d_prefix = nullptr; size_t prefix_cap = 0;
    uint32_t *d_limit_pos = nullptr; size_t limit_cap = 0;
    uint32_t *d_hit_bound = nullptr; size_t bound_cap = 0;
    uint32_t *d_plan = nullptr;      // [0]=num_iter [1]=num_hits
    uint32_t *d_counters = nullptr;  // CTR_WORDS counters, see the CTR_* enum (kernels_filter.cuh)
    uint2 *d_hits = nullptr; size_t hits_cap = 0;
    SurvRec *d_surv = nullptr; size_t surv_cap = 0;  // filter survivors (anchor pair + key)
    SurvRec *d_surv2 = nullptr; size_t surv2_cap = 0; // survivors k_extend_wide hands on to k_extend_hits
    SurvRec *d_surv_m = nullptr; size_t surv_m_cap = 0; // merge pass: the representatives (kernels_merge.cuh)
    unsigned long long *d_mkeys = nullptr; size_t mkeys_cap = 0; // merge pass: sort keys, double buffer
    uint32_t *d_midx = nullptr; size_t midx_cap = 0;             // merge pass: survivor indices, double buffer
    unsigned long long *d_dedup = nullptr;           // exact-duplicate table: k0[slots], k1[slots], tagbits[slots]
    Anchor *d_anchors_a = nullptr; size_t anchors_a_cap = 0;
    Anchor *d_anchors_b = nullptr; size_t anchors_b_cap = 0;
    sa_segment *d_out = nullptr; size_t out_cap = 0;
    uint8_t *d_temp = nullptr; size_t temp_cap = 0;
    uint32_t *d_flags = nullptr; size_t flags_cap = 0;
    uint32_t *d_excl = nullptr; size_t excl_cap = 0;
    uint64_t *h_bases = nullptr; size_t h_bases_cap = 0; // pinned: base seed words of a canonical seed vector
    uint64_t *d_bases = nullptr; size_t d_bases_cap = 0;
    uint32_t *h_small = nullptr; // pinned, 64 words: [0..15] counters, [16..19] plan, [20] seed count staging
    sa_segment *h_out = nullptr; // pinned staging for the first FINALIZE_CAP result records
};

struct GpuCtx {
    int device = 0;
    cudaStream_t ctrl = nullptr;
    SeqPlanes ref;
    SeqPlanes q_fwd[SA_BUFFER_DEPTH], q_rc[SA_BUFFER_DEPTH];
    uint32_t *d_index = nullptr;
    uint32_t *d_pos = nullptr;
    uint32_t index_size = 0, num_pos = 0;
    int *d_sub_mat = nullptr;
    uint8_t *d_ascii = nullptr; size_t ascii_cap = 0; // ASCII staging buffer of block uploads
    cudaEvent_t ev_up = nullptr;                      // upload / peer copy of the current block has landed
    uint64_t calls = 0;                               // SeedAndFilter calls served by this GPU (under stats_mu)
    std::vector<Workspace *> ws;
};

struct Global {
    bool interface_ready = false, processor_ready = false;
    std::vector<GpuCtx> gpus;
    // InitializeProcessor scalars (seed_filter.cu:20-27)
    uint32_t max_seeds = 0, max_hits = 0, max_hits_device = 0, seed_size = 0;
    int sub_mat[64] = {};
    int xdrop = 0, hspthresh = 0, noentropy = 0, diag_all_positive = 0, transition = 0;
    uint32_t term_codes = 0;   // non-ACGT codes the filter stage treats as X-drop terminators (screen_terminator_codes)
    uint32_t strict_term_codes = 0; // ... those that trip the X-drop rule against every code
    uint32_t zero_flat = 0, zero_partners = 0; // code sets of the zero-run planes (screen_bound.h: zero_run_codes); 0 = none
    bool filter_ok = false;    // ACGT x ACGT scores fit int8: the filter stage is usable
    bool use_filter = true;    // SEGALIGN_B200_FILTER=0 routes every hit to the exact kernel
    bool use_dedup = true;     // SEGALIGN_B200_DEDUP=0 appends every passing record (no duplicate table)
    bool use_fused = true;     // SEGALIGN_B200_FUSED=0 always takes the general (materialised hit list) path
    int filter2_grid = 0;      // tile-walk kernel (k_filter_hits2)
    int filter3_grid = 0;      // popcount screen + tile walk (k_filter_hits3), the default on the fused path
    int filter3_grid_alone = 0; // the same when no other call is in flight: every block slot of every SM
    std::atomic<int> calls_in_flight[64] = {}; // per GPU of the pool
    int filter_kernel = 3;     // SEGALIGN_B200_FILTER_KERNEL=2 selects the tile-walk-only kernel on the fused path
    ScreenConsts screen = {};  // class scores of the popcount screen (screen_bound.h)
    int extend_grid = 0;
    int wide_grid = 0;         // k_extend_wide (warp per hit), SEGALIGN_B200_WIDE=0 sends all survivors to k_extend_hits
    bool use_wide = true;
    uint32_t finalize_cap = FINALIZE_CAP; // anchors one block sorts in shared memory; more take the device-wide radix path (SEGALIGN_B200_FINALIZE_CAP lowers it: tests)
    uint32_t merge_min = 65536; // calls with more filter survivors than this take the merge pass (SEGALIGN_B200_MERGE_MIN, 0 = never)
    bool use_compact = true;   // SEGALIGN_B200_COMPACT_SEEDS=0: always copy seed vectors as they are
    uint32_t ref_len = 0;
    bool ref_loaded = false, table_ready = false;
    bool ascii_holds_ref = false; // the per-GPU ASCII staging buffers still hold the reference block (sa_rm_send_query)
    bool rm_query = false;        // query slot 0 = the reference block itself + its reverse complement (repeat-masker variant)
    uint32_t query_len[SA_BUFFER_DEPTH] = {};
    bool query_loaded[SA_BUFFER_DEPTH] = {};
    ShapeDesc shape = {};
    bool shape_set = false;
    // workspace pool (the reference's mu/cv/available_gpus, store_gpu.h:4-6)
    std::mutex mu;
    std::condition_variable cv;
    std::vector<Workspace *> free_ws;
    int ws_per_gpu = 3;
    // stats
    std::mutex stats_mu;
    sa_stats stats = {};
    bool profiling = false; // sa_set_profiling: per-phase CUDA events (adds a stream sync per call); off by default
};

Global G;

void add_launches(uint64_t n) {
    std::lock_guard<std::mutex> l(G.stats_mu);
    G.stats.launches += n;
}

// ------------------------------------------------------------------ sequence planes
// Sequence planes come from the device's stream-ordered memory pool (release threshold = keep
// everything): ClearQuery + SendQueryWriteRequest of the next block then reuse the previous block's
// memory without a cudaFree / cudaMalloc pair -- those synchronise the device and, with several
// processes on one box, cost tens of milliseconds per block.
int free_planes(SeqPlanes &p, cudaStream_t st) {
    if (p.b8) CU(cudaFreeAsync(p.b8, st), SA_ERR_FREE);
    if (p.p2_base) CU(cudaFreeAsync(p.p2_base, st), SA_ERR_FREE);
    if (p.m1) CU(cudaFreeAsync(p.m1, st), SA_ERR_FREE);
    if (p.softmap) CU(cudaFreeAsync(p.softmap, st), SA_ERR_FREE);
    if (p.rec_base) CU(cudaFreeAsync(p.rec_base, st), SA_ERR_FREE);
    if (p.zr) CU(cudaFreeAsync(p.zr, st), SA_ERR_FREE);
    p = SeqPlanes();
    return SA_OK;
}

int alloc_planes(SeqPlanes &p, uint32_t len, const char *tag, cudaStream_t st) {
    p.len = len;
    p.words = ((size_t)len + 31) / 32 + PAD_WORDS;
    cudaError_t e = cudaMallocAsync((void **)&p.b8, (size_t)len + 64, st);
    if (e == cudaSuccess) e = cudaMallocAsync((void **)&p.p2_base, (p.words + REC_FRONT) * sizeof(uint64_t), st);
    if (e == cudaSuccess) e = cudaMemsetAsync(p.p2_base, 0, REC_FRONT * sizeof(uint64_t), st);
    p.p2 = p.p2_base ? p.p2_base + REC_FRONT : nullptr;
    p.softmap_words = (uint32_t)((p.words + REC_FRONT) / 32 + 2);
    if (e == cudaSuccess) e = cudaMallocAsync((void **)&p.softmap, ((size_t)p.softmap_words + 1) * sizeof(uint32_t), st);
    if (e == cudaSuccess) e = cudaMallocAsync((void **)&p.m1, p.words * sizeof(uint32_t), st);
    if (e == cudaSuccess) e = cudaMallocAsync((void **)&p.rec_base, (p.words + REC_FRONT + 1) * sizeof(uint4), st);
    p.rec = p.rec_base ? p.rec_base + REC_FRONT : nullptr;
    if (e != cudaSuccess)
        return fail(SA_ERR_MALLOC, "cudaMalloc of %lu bytes for %s failed with error \" %s \"",
                    (unsigned long)len, tag, cudaGetErrorString(e));
    return SA_OK;
}

// filter records of one block under the current terminator set, and its zero-run planes under the current flat /
// partner code sets (async on the control stream)
int build_records(GpuCtx &g, SeqPlanes &p) {
    cudaMemsetAsync(p.softmap, 0, ((size_t)p.softmap_words + 1) * sizeof(uint32_t), g.ctrl);
    k_pack_records<<<grid_for(p.words + REC_FRONT, 256), 256, 0, g.ctrl>>>(p.b8, p.len, p.rec, REC_FRONT,
                                                                            (uint32_t)p.words, G.term_codes, p.softmap, p.softmap_words);
    p.term_codes = G.term_codes;
    p.zero_codes = G.zero_flat | (G.zero_partners << 8);
    if (G.zero_flat) {
        p.coarse_words = (uint32_t)(p.words / 1024 + 3);
        const size_t zw = (p.words + 3) & ~(size_t)3; // plane stride: keeps g1 16-byte aligned
        const size_t total = 2 * zw + 2 * (size_t)p.coarse_words + 1;
        if (!p.zr) {
            cudaError_t e = cudaMallocAsync((void **)&p.zr, total * sizeof(uint32_t), g.ctrl);
            if (e != cudaSuccess)
                return fail(SA_ERR_MALLOC, "cudaMalloc of %zu bytes for the zero-run planes failed with error \" %s \"",
                            total * sizeof(uint32_t), cudaGetErrorString(e));
        }
        p.f1 = p.zr; p.g1 = p.f1 + zw; p.F1k = p.g1 + zw; p.G1k = p.F1k + p.coarse_words;
        uint32_t *counter = p.G1k + p.coarse_words;
        cudaMemsetAsync(p.F1k, 0, (2 * (size_t)p.coarse_words + 1) * sizeof(uint32_t), g.ctrl);
        k_pack_zero_planes<<<grid_for(p.words, 256), 256, 0, g.ctrl>>>(p.b8, p.len, p.f1, p.g1, (uint32_t)p.words, G.zero_flat,
                                                                     G.zero_partners, counter);
        k_coarse_zero_planes<<<grid_for((size_t)p.coarse_words * 32, 256), 256, 0, g.ctrl>>>(p.f1, p.g1, (uint32_t)p.words, p.F1k, p.G1k,
                                                                                          p.coarse_words);
        add_launches(2);
    } else if (p.zr) {
        cudaFreeAsync(p.zr, g.ctrl);
        p.zr = p.f1 = p.g1 = p.F1k = p.G1k = nullptr;
    }
    return SA_OK;
}
// host copies of the block's soft-record and flat-cell counts (after the control stream has been awaited)
int read_soft_flag(SeqPlanes &p) {
    p.has_soft = 0;
    if (p.softmap) CU(cudaMemcpy(&p.has_soft, p.softmap + p.softmap_words, sizeof(uint32_t), cudaMemcpyDeviceToHost), SA_ERR_MEMCPY);
    p.has_flat = 0;
    if (p.zr) CU(cudaMemcpy(&p.has_flat, p.G1k + p.coarse_words, sizeof(uint32_t), cudaMemcpyDeviceToHost), SA_ERR_MEMCPY);
    return SA_OK;
}

// ASCII staging buffer of one GPU (plain cudaMalloc, kept and grown: peer copies read it)
int ensure_ascii(GpuCtx &g, size_t len) {
    if (g.d_ascii && g.ascii_cap >= len + 64) return SA_OK;
    if (g.d_ascii) CU(cudaFree(g.d_ascii), SA_ERR_FREE);
    g.d_ascii = nullptr; g.ascii_cap = 0;
    const size_t cap = len + len / 8 + 4096;
    cudaError_t e = cudaMalloc((void **)&g.d_ascii, cap);
    if (e != cudaSuccess)
        return fail(SA_ERR_MALLOC, "cudaMalloc of %zu bytes for the ASCII staging buffer failed with error \" %s \"",
                    cap, cudaGetErrorString(e));
    g.ascii_cap = cap;
    return SA_OK;
}

// encode the ASCII block in g.d_ascii into the planes (async on the control stream); fwd always, rc if rc != nullptr
int enqueue_encode(GpuCtx &g, uint32_t len, SeqPlanes &fwd, SeqPlanes *rc, const char *tag) {
    TRY(alloc_planes(fwd, len, tag, g.ctrl));
    if (rc) TRY(alloc_planes(*rc, len, tag, g.ctrl));
    if (len > 0) {
        int grid = grid_for(((size_t)len + 15) / 16, 256);
        k_encode_b8<<<grid, 256, 0, g.ctrl>>>(g.d_ascii, len, fwd.b8, rc ? rc->b8 : nullptr);
    }
    k_pack_planes<<<grid_for(fwd.words, 256), 256, 0, g.ctrl>>>(fwd.b8, len, fwd.p2, fwd.m1,
                                                                 (uint32_t)fwd.words);
    if (rc)
        k_pack_planes<<<grid_for(rc->words, 256), 256, 0, g.ctrl>>>(rc->b8, len, rc->p2, rc->m1,
                                                                     (uint32_t)rc->words);
    TRY(build_records(g, fwd));
    if (rc) TRY(build_records(g, *rc));
    add_launches(rc ? 5 : 3);
    CU(cudaGetLastError(), SA_ERR_KERNEL);
    return SA_OK;
}

// One block (reference, or query slot `slot` when slot >= 0) onto EVERY GPU of the pool: the ASCII bytes
// cross PCIe once (to GPU 0) and reach the other GPUs by peer copies over NVLink; each GPU then encodes
// its own planes.  Everything is enqueued on the per-GPU control streams first and awaited at the end,
// so the GPUs work concurrently (the reference uploads and encodes GPU after GPU with blocking copies,
// common/seed_filter_interface.cu:82-101, src/seed_filter.cu:899-919).
int upload_block_all_gpus(const char *src, uint32_t len, int slot, const char *tag) {
    const size_t n = G.gpus.size();
    GpuCtx &g0 = G.gpus[0];
    CU(cudaSetDevice(g0.device), SA_ERR_SET_DEVICE);
    TRY(ensure_ascii(g0, len));
    if (len) {
        cudaError_t e = cudaMemcpyAsync(g0.d_ascii, src, len, cudaMemcpyHostToDevice, g0.ctrl);
        if (e != cudaSuccess)
            return fail(SA_ERR_MEMCPY, "cudaMemcpy of %lu bytes for %s failed with error \" %s \"",
                        (unsigned long)len, tag, cudaGetErrorString(e));
    }
    CU(cudaEventRecord(g0.ev_up, g0.ctrl), SA_ERR_KERNEL);
    for (size_t i = 1; i < n; i++) {
        GpuCtx &g = G.gpus[i];
        CU(cudaSetDevice(g.device), SA_ERR_SET_DEVICE);
        TRY(ensure_ascii(g, len));
        CU(cudaStreamWaitEvent(g.ctrl, g0.ev_up, 0), SA_ERR_KERNEL);
        if (len) CU(cudaMemcpyPeerAsync(g.d_ascii, g.device, g0.d_ascii, g0.device, len, g.ctrl), SA_ERR_MEMCPY);
        CU(cudaEventRecord(g.ev_up, g.ctrl), SA_ERR_KERNEL);
    }
    for (size_t i = 0; i < n; i++) {
        GpuCtx &g = G.gpus[i];
        CU(cudaSetDevice(g.device), SA_ERR_SET_DEVICE);
        if (slot < 0) TRY(enqueue_encode(g, len, g.ref, nullptr, tag));
        else TRY(enqueue_encode(g, len, g.q_fwd[slot], &g.q_rc[slot], tag));
    }
    // GPU 0's staging buffer may be overwritten by the next upload only after every peer has read it
    CU(cudaSetDevice(g0.device), SA_ERR_SET_DEVICE);
    for (size_t i = 1; i < n; i++) CU(cudaStreamWaitEvent(g0.ctrl, G.gpus[i].ev_up, 0), SA_ERR_KERNEL);
    for (size_t i = 0; i < n; i++) {
        CU(cudaSetDevice(G.gpus[i].device), SA_ERR_SET_DEVICE);
        CU(cudaStreamSynchronize(G.gpus[i].ctrl), SA_ERR_KERNEL);
        if (slot < 0) TRY(read_soft_flag(G.gpus[i].ref));
        else { TRY(read_soft_flag(G.gpus[i].q_fwd[slot])); G.gpus[i].q_rc[slot].has_flat = G.gpus[i].q_fwd[slot].has_flat; }
    }
    return SA_OK;
}

// ------------------------------------------------------------------ workspace pool
Workspace *acquire_ws() {
    std::unique_lock<std::mutex> lk(G.mu);
    G.cv.wait(lk, [] { return !G.free_ws.empty(); });
    Workspace *w = G.free_ws.back();
    G.free_ws.pop_back();
    return w;
}
void release_ws(Workspace *w) {
    {
        std::lock_guard<std::mutex> lk(G.mu);
        G.free_ws.push_back(w);
    }
    G.cv.notify_one();
}
struct WsGuard {
    Workspace *w;
    explicit WsGuard(Workspace *w_) : w(w_) {}
    ~WsGuard() { release_ws(w); }
};

int make_workspace(int gpu_index, Workspace *&out) {
    Workspace *w = new Workspace();
    w->gpu = gpu_index;
    CU(cudaStreamCreateWithFlags(&w->stream, cudaStreamNonBlocking), SA_ERR_KERNEL);
    for (auto &e : w->ev) CU(cudaEventCreate(&e), SA_ERR_KERNEL);
    CU(cudaMalloc((void **)&w->d_plan, 4 * sizeof(uint32_t)), SA_ERR_MALLOC);
    CU(cudaMalloc((void **)&w->d_counters, CTR_WORDS * sizeof(uint32_t)), SA_ERR_MALLOC);
    CU(cudaMallocHost((void **)&w->h_small, 64 * sizeof(uint32_t)), SA_ERR_MALLOC);
    CU(cudaMallocHost((void **)&w->h_out, FINALIZE_CAP * sizeof(sa_segment)), SA_ERR_MALLOC);
    TRY(ensure(w->d_out, w->out_cap, FINALIZE_CAP, "hsp_out", 1, 1));
    CU(cudaMalloc((void **)&w->d_dedup, (size_t)kDedupSlots * 20), SA_ERR_MALLOC);
    TRY(ensure(w->d_seeds, w->seeds_cap, std::max<size_t>(G.max_seeds, 1024), "seed_offsets", 1, 1));
    TRY(ensure(w->d_prefix, w->prefix_cap, std::max<size_t>(G.max_seeds, 1024), "hit_num", 1, 1));
    TRY(ensure(w->d_limit_pos, w->limit_cap, 64, "limit_pos", 1, 1));
    TRY(ensure(w->d_hit_bound, w->bound_cap, 64, "hit_bound", 1, 1));
    out = w;
    return SA_OK;
}

void destroy_workspace(Workspace *w) {
    cudaFree(w->d_seeds); cudaFree(w->d_prefix); cudaFree(w->d_limit_pos);
    cudaFree(w->d_hit_bound); cudaFree(w->d_plan); cudaFree(w->d_counters);
    cudaFree(w->d_surv_m); cudaFree(w->d_mkeys); cudaFree(w->d_midx);
    cudaFree(w->d_hits); cudaFree(w->d_surv); cudaFree(w->d_surv2); cudaFree(w->d_dedup); cudaFree(w->d_anchors_a); cudaFree(w->d_anchors_b);
    cudaFree(w->d_out); cudaFree(w->d_temp); cudaFree(w->d_flags); cudaFree(w->d_excl);
    cudaFreeHost(w->h_small);
    if (w->h_bases) cudaFreeHost(w->h_bases);
    cudaFree(w->d_bases);
    cudaFreeHost(w->h_out);
    for (auto &e : w->ev) if (e) cudaEventDestroy(e);
    if (w->stream) cudaStreamDestroy(w->stream);
    delete w;
}

// ------------------------------------------------------------------ the per-call pipeline
enum Phase { PH_START = 0, PH_SEEDS, PH_PLAN, PH_LOOKUP, PH_FILTER, PH_EXTEND, PH_SORT, PH_D2H, PH_COUNT };
struct PhaseTimer {
    Workspace *w;
    bool on;
    bool have[PH_COUNT] = {};
    explicit PhaseTimer(Workspace *w_) : w(w_), on(G.profiling) {}
    void mark(Phase p) { if (on) { cudaEventRecord(w->ev[p], w->stream); have[p] = true; } }
    // time between the latest recorded phase before b and b
    float ms(Phase b) const {
        if (!on || !have[b]) return 0.f;
        int a = (int)b - 1;
        while (a >= 0 && !have[a]) a--;
        float t = 0;
        if (a >= 0) cudaEventElapsedTime(&t, w->ev[a], w->ev[b]);
        return t;
    }
};

// Device-wide sort of n anchors under one of the orders of kernels_sort.cuh: three stable 64-bit radix passes
// (least significant key word first) over a permutation, then one gather.  in -> out (distinct buffers).
int radix_sort_anchors(Workspace *w, const Anchor *in, Anchor *out, uint32_t n, int order) {
    cudaStream_t st = w->stream;
    if (n == 0) return SA_OK;
    TRY(ensure(w->d_mkeys, w->mkeys_cap, 2 * (size_t)n, "sort_keys"));
    TRY(ensure(w->d_midx, w->midx_cap, 2 * (size_t)n, "sort_index"));
    cub::DoubleBuffer<unsigned long long> dk(w->d_mkeys, w->d_mkeys + n);
    cub::DoubleBuffer<uint32_t> dv(w->d_midx, w->d_midx + n);
    size_t bytes = 0;
    CU(cub::DeviceRadixSort::SortPairs(nullptr, bytes, dk, dv, (int)n, 0, 64, st), SA_ERR_KERNEL);
    TRY(ensure(w->d_temp, w->temp_cap, bytes, "sort_temp"));
    for (int word = 2; word >= 0; word--) {
        k_anchor_keys<<<grid_for(n, 256), 256, 0, st>>>(in, word == 2 ? nullptr : dv.Current(), n, order, word, dk.Current(), dv.Current());
        CU(cub::DeviceRadixSort::SortPairs(w->d_temp, bytes, dk, dv, (int)n, 0, 64, st), SA_ERR_KERNEL);
    }
    k_anchor_gather<<<grid_for(n, 256), 256, 0, st>>>(in, dv.Current(), n, out);
    CU(cudaGetLastError(), SA_ERR_KERNEL);
    return SA_OK;
}

// What one SeedAndFilter call works on.
struct CallInput {
    int src;             // SRC_SEEDS: seed words already in w->d_seeds, count in d_plan[2];  SRC_RANGE: query range
    uint32_t max_items;  // seed words (SRC_SEEDS) or positions x words per position (SRC_RANGE)
    uint32_t q_start, q_end, per; // SRC_RANGE
    int transition;
    // repeat-masker variant (repeat_masker_src/seed_filter.cu:724): the block against itself / its reverse
    // complement, hits outside [win_lo, win_hi] counted but not extended, its own sort/unique chain and header
    bool rm = false;
    uint32_t win_lo = 0, win_hi = 0xFFFFFFFFu;
};

// SRC_RANGE, general path only: seed words of src/seeder.cpp:57-74 on the device; their count
// stays on the device (d_plan[2]).
int enqueue_range_seeding(Workspace *w, const SeqPlanes &q, const CallInput &in) {
    cudaStream_t st = w->stream;
    const uint32_t n = in.q_end - in.q_start;
    TRY(ensure(w->d_flags, w->flags_cap, n, "seed_flags"));
    TRY(ensure(w->d_excl, w->excl_cap, (size_t)n + 1, "seed_excl"));
    TRY(ensure(w->d_seeds, w->seeds_cap, in.max_items, "seed_offsets"));
    size_t bytes = 0;
    CU(cub::DeviceScan::ExclusiveSum(nullptr, bytes, w->d_flags, w->d_excl, (int)n, st), SA_ERR_KERNEL);
    TRY(ensure(w->d_temp, w->temp_cap, bytes, "scan_temp"));
    k_seed_flags<<<grid_for(n, 256), 256, 0, st>>>(q.m1, G.shape.span, in.q_start, in.q_end, w->d_flags);
    CU(cub::DeviceScan::ExclusiveSum(w->d_temp, bytes, w->d_flags, w->d_excl, (int)n, st), SA_ERR_KERNEL);
    k_seed_emit<<<grid_for(n, 256), 256, 0, st>>>(q.p2, w->d_flags, w->d_excl, G.shape, in.transition, in.q_start, in.q_end,
                                                  w->d_seeds, w->d_plan + 2);
    add_launches(4);
    return SA_OK;
}

// Merge pass (kernels_merge.cuh): the n filter survivors in w->d_surv sorted by (diagonal, anchor), those
// connected to their predecessor by an all-match stretch dropped, the rest in w->d_surv_m (count in
// counters[CTR_MERGED]).  Enqueued on the call's stream; n is known because the first attempt's stage B
// declined the call and the host has seen its counters.
int enqueue_merge(Workspace *w, const ExtendParams &P, uint32_t n) {
    cudaStream_t st = w->stream;
    TRY(ensure(w->d_surv_m, w->surv_m_cap, n, "merged_survivors"));
    TRY(ensure(w->d_mkeys, w->mkeys_cap, 2 * (size_t)n, "merge_keys"));
    TRY(ensure(w->d_midx, w->midx_cap, 2 * (size_t)n, "merge_index"));
    cub::DoubleBuffer<unsigned long long> dk(w->d_mkeys, w->d_mkeys + n);
    cub::DoubleBuffer<uint32_t> dv(w->d_midx, w->d_midx + n);
    size_t bytes = 0;
    CU(cub::DeviceRadixSort::SortPairs(nullptr, bytes, dk, dv, (int)n, 0, 64, st), SA_ERR_KERNEL);
    TRY(ensure(w->d_temp, w->temp_cap, bytes, "sort_temp"));
    CU(cudaMemsetAsync(w->d_counters + CTR_MERGED, 0, sizeof(uint32_t), st), SA_ERR_MEMCPY);
    k_merge_keys<<<grid_for(n, 256), 256, 0, st>>>(w->d_surv, n, dk.Current(), dv.Current());
    CU(cub::DeviceRadixSort::SortPairs(w->d_temp, bytes, dk, dv, (int)n, 0, 64, st), SA_ERR_KERNEL);
    k_merge_mark<<<grid_for(n, 256), 256, 0, st>>>(P, w->d_surv, dv.Current(), n, w->d_surv_m, w->d_counters);
    CU(cudaGetLastError(), SA_ERR_KERNEL);
    return SA_OK;
}

// One SeedAndFilter call.  Produces the malloc'd result (header + HSPs).
//
// Fast path (the filter stage is usable): ONE kernel does seeding (SRC_RANGE) / seed-word reading
// (SRC_SEEDS), seed-position-table lookup, bucket expansion and the score filter; it is followed by
// the exact extension of the survivors and a one-block sort/dedupe/sort, and by a single
// synchronisation that brings back the counters and the records.  Hit counts, prefix sums and the
// hit list are never written to HBM.  It is valid when the call is one reference iteration pair
// (num_hits < MAX_HITS), which is known from the hit total after the fact; otherwise --
// or when the filter is disabled -- the call is (re)played on the general path: hit counts -> scan
// -> iteration plan -> materialised hit list -> filter -> exact, any number of iterations.
// Either way nothing between the seeds and the result needs a host round trip: sizes the host does
// not know yet are read by the kernels from device memory, buffers are sized from what earlier
// calls needed, and the synchronisation at the end tells the host whether a buffer was too small
// (then it grows the buffer and replays -- rare after the first calls of a block).
int run_pipeline(Workspace *w, const CallInput &in, int rev, uint32_t buffer, sa_segment **out,
                 uint32_t *out_count, uint32_t *out_num_seeds, PhaseTimer &pt) {
    GpuCtx &g = G.gpus[w->gpu];
    cudaStream_t st = w->stream;
    struct InFlight { std::atomic<int> &c;
                      explicit InFlight(std::atomic<int> &c_) : c(c_) { c.fetch_add(1, std::memory_order_relaxed); }
                      ~InFlight() { c.fetch_sub(1, std::memory_order_relaxed); } } in_flight(G.calls_in_flight[w->gpu & 63]);
    uint64_t launches = 0;
    uint32_t num_hits = 0, num_iter = 0, num_seeds = 0, n_pre = 0, n_final = 0, n_surv = 0, n_walked = 0;
    unsigned long long ext_cells = 0, num_hits64 = 0;
    const SeqPlanes &q = rev ? g.q_rc[buffer] : g.q_fwd[buffer];
    const bool filter = G.filter_ok && G.use_filter;
    bool fused = filter && G.use_fused;
    bool seeds_ready = in.src == SRC_SEEDS; // seed words in d_seeds, count in d_plan[2]
    bool staged = false;                    // result records already in w->h_out
    bool merged = false;                    // stage B replays on the representatives of the merge pass
    const uint32_t max_items = in.max_items;

    if (!w->d_surv) TRY(ensure(w->d_surv, w->surv_cap, (size_t)1 << 20, "survivors", 1, 1));
    if (!w->d_anchors_a) TRY(ensure(w->d_anchors_a, w->anchors_a_cap, (size_t)1 << 20, "hsp_reduced", 1, 1));

    ExtendParams P;
    P.rb8 = g.ref.b8; P.rp2 = g.ref.p2; P.rm1 = g.ref.m1; P.ref_len = g.ref.len;
    P.qb8 = q.b8; P.qp2 = q.p2; P.qm1 = q.m1; P.query_len = q.len;
    P.xdrop = G.xdrop; P.hspthresh = G.hspthresh; P.noentropy = G.noentropy;
    P.diag_all_positive = G.diag_all_positive;
    P.scores_fit_int8 = G.filter_ok;
    P.soft_runs = (((G.strict_term_codes >> L_NT) & 1u) == 0 || ((G.strict_term_codes >> N_NT) & 1u) == 0) ? 1 : 0;
    P.win_lo = in.win_lo; P.win_hi = in.win_hi;
    P.zskip = (g.ref.zr && q.zr && (g.ref.has_flat || q.has_flat)) ? 1 : 0;
    P.rz = ZeroPlanes{g.ref.f1, g.ref.g1, g.ref.F1k, g.ref.G1k};
    P.qz = ZeroPlanes{q.f1, q.g1, q.F1k, q.G1k};
    FilterParams F;
    F.rrec = g.ref.rec; F.qrec = q.rec;
    F.rp2 = g.ref.p2; F.rsoft = g.ref.softmap; F.ref_has_soft = g.ref.has_soft ? 1 : 0;
    F.xdrop = G.xdrop; F.hspthresh = G.hspthresh; F.diag_all_positive = G.diag_all_positive;
    F.k_mul = 4u | (64u << 8); F.k_m4 = 0x01010101u;
    HitSource H = {};
    H.plan = w->d_plan;
    H.num_items = max_items;
    H.index_table = g.d_index; H.pos_table = g.d_pos; H.seed_size = G.seed_size;
    H.index_size = g.index_size; H.query_len = q.len;
    H.win_lo = in.win_lo; H.win_hi = in.win_hi;
    const SeedBounds SB = {g.index_size, q.len, G.seed_size};
    H.j0 = in.q_start; H.per = in.per; H.shape = G.shape;
    DedupTable D;
    D.k0 = w->d_dedup; D.k1 = w->d_dedup + kDedupSlots;
    D.tagbits = reinterpret_cast<uint32_t *>(w->d_dedup + 2 * (size_t)kDedupSlots);
    D.mask = (G.use_dedup && G.hspthresh > 0) ? kDedupSlots - 1 : 0;
    const size_t lut_bytes = FILTER_LUT_WORDS * sizeof(uint32_t);

    for (int attempt = 0;; attempt++) {
        if (attempt > 12) return fail(SA_ERR_KERNEL, "SeedAndFilter did not converge on buffer sizes");
        const uint32_t surv_cap = (uint32_t)std::min<size_t>(w->surv_cap, 0xFFFFFFFFu);
        const uint32_t anchor_cap = (uint32_t)std::min<size_t>(w->anchors_a_cap, 0xFFFFFFFFu);
        uint32_t hits_cap = 0;
        if (!merged) CU(cudaMemsetAsync(w->d_counters, 0, CTR_WORDS * sizeof(uint32_t), st), SA_ERR_MEMCPY);
        else { // stage A's counters stay; only what stage B and the finalisation write starts over
            CU(cudaMemsetAsync(w->d_counters + CTR_ANCHORS, 0, 2 * sizeof(uint32_t), st), SA_ERR_MEMCPY); // + CTR_DEDUPE
            CU(cudaMemsetAsync(w->d_counters + CTR_OUT, 0, sizeof(uint32_t), st), SA_ERR_MEMCPY);
            CU(cudaMemsetAsync(w->d_counters + CTR_SURV2, 0, sizeof(uint32_t), st), SA_ERR_MEMCPY);
        }
        if (merged) {
            // stage A already ran: its survivors were reduced to one representative per all-match chain
        } else if (fused) {
            H.seeds = w->d_seeds;
            if (G.filter_kernel == 3 && G.screen.enabled) {
                FilterParams F3 = F;
                F3.k_mul = SCR_K_MUL;
                // a launch normally leaves one block slot per SM to the next call's kernel (their head and tail
                // overlap); a call that is alone on the device takes them all
                const int grid3 = G.calls_in_flight[w->gpu & 63].load(std::memory_order_relaxed) <= 1 ? G.filter3_grid_alone : G.filter3_grid;
                if (in.src == SRC_SEEDS)
                    k_filter_hits3<SRC_SEEDS><<<grid3, SCR_THREADS, SCR_SMEM_BYTES, st>>>(F3, G.screen, H, g.d_sub_mat, w->d_surv, surv_cap, w->d_counters);
                else
                    k_filter_hits3<SRC_RANGE><<<grid3, SCR_THREADS, SCR_SMEM_BYTES, st>>>(F3, G.screen, H, g.d_sub_mat, w->d_surv, surv_cap, w->d_counters);
            } else {
                if (in.src == SRC_SEEDS)
                    k_filter_hits2<SRC_SEEDS><<<G.filter2_grid, FILTER_THREADS, lut_bytes, st>>>(F, H, g.d_sub_mat, w->d_surv, surv_cap, w->d_counters);
                else
                    k_filter_hits2<SRC_RANGE><<<G.filter2_grid, FILTER_THREADS, lut_bytes, st>>>(F, H, g.d_sub_mat, w->d_surv, surv_cap, w->d_counters);
            }
            launches++;
            pt.mark(PH_FILTER);
        } else {
            if (!seeds_ready) {
                TRY(enqueue_range_seeding(w, q, in));
                seeds_ready = true;
                pt.mark(PH_SEEDS);
            }
            if (!w->d_hits) TRY(ensure(w->d_hits, w->hits_cap, (size_t)1 << 25, "hsp", 1, 1));
            hits_cap = (uint32_t)std::min<size_t>(w->hits_cap, 0xFFFFFFFFu);
            TRY(ensure(w->d_prefix, w->prefix_cap, max_items, "hit_num"));
            size_t bytes = 0;
            CU(cub::DeviceScan::InclusiveSum(nullptr, bytes, w->d_prefix, w->d_prefix, (int)max_items, st), SA_ERR_KERNEL);
            TRY(ensure(w->d_temp, w->temp_cap, bytes, "scan_temp"));
            // 1. bucket sizes + inclusive scan (seed_filter.cu:712-714)
            k_count_hits<<<grid_for(max_items, 256), 256, 0, st>>>(w->d_seeds, max_items, w->d_plan + 2, g.d_index, SB, w->d_prefix,
                                                                        reinterpret_cast<unsigned long long *>(w->d_counters + CTR_NHITS64));
            CU(cub::DeviceScan::InclusiveSum(w->d_temp, bytes, w->d_prefix, w->d_prefix, (int)max_items, st), SA_ERR_KERNEL);
            // 2. iteration plan on the device (seed_filter.cu:718-745)
            k_plan_iterations<<<1, 32, 0, st>>>(w->d_prefix, max_items, G.max_hits, (uint32_t)std::min(w->limit_cap, w->bound_cap),
                                                w->d_limit_pos, w->d_hit_bound, w->d_plan);
            pt.mark(PH_PLAN);
            // 3. flat hit expansion (seed_filter.cu:760)
            k_expand_hits<<<grid_for(((size_t)max_items + 31) / 32 * 32, 256), 256, 0, st>>>(
                w->d_seeds, max_items, w->d_plan + 2, g.d_index, g.d_pos, w->d_prefix, G.seed_size, SB, w->d_hits, hits_cap);
            pt.mark(PH_LOOKUP);
            launches += 5;
            // 4a. stage A: conservative score bound over all hits -> survivor records
            if (filter) {
                H.hits = w->d_hits; H.hits_cap = hits_cap;
                k_filter_hits2<SRC_HITS><<<G.filter2_grid, FILTER_THREADS, lut_bytes, st>>>(F, H, g.d_sub_mat, w->d_surv, surv_cap, w->d_counters);
                launches++;
                pt.mark(PH_FILTER);
            }
        }
        // 4b. stage B: exact extension of the survivors (seed_filter.cu:762-774)
        if (D.mask) { // the duplicate table of stage B (cleared here, behind the filter kernel: its phase timer is the kernel's)
            CU(cudaMemsetAsync(w->d_dedup, 0xFF, (size_t)kDedupSlots * 16, st), SA_ERR_MEMCPY);
            CU(cudaMemsetAsync(D.tagbits, 0, (size_t)kDedupSlots * 4, st), SA_ERR_MEMCPY);
        }
        // (a call with more than merge_min survivors is declined here and replayed on the representatives
        // of the merge pass, see below; only the fused path, whose calls are one dedupe scope pair)
        const SurvRec *sv = merged ? w->d_surv_m : w->d_surv;
        const int sv_ctr = merged ? (int)CTR_MERGED : (int)CTR_SURV;
        const uint32_t decline = (fused && filter && !merged && P.diag_all_positive) ? G.merge_min : 0u;
        if (filter && G.use_wide) {
            // warp-per-hit pass over all survivors; what needs the entropy counters goes on to the lane-pair kernel
            if (w->surv2_cap < w->surv_cap) TRY(ensure(w->d_surv2, w->surv2_cap, w->surv_cap, "survivors2", 1, 1));
            k_extend_wide<<<G.wide_grid, WIDE_THREADS, 0, st>>>(P, g.d_sub_mat, sv, surv_cap, sv_ctr, decline, w->d_surv2, fused ? 1 : 0,
                                                                w->d_hit_bound, w->d_plan, w->d_anchors_a, anchor_cap, w->d_counters, D);
            k_extend_hits<<<G.extend_grid, EXTEND_THREADS, 0, st>>>(
                P, g.d_sub_mat, w->d_hits, hits_cap, w->d_surv2, surv_cap, (int)CTR_SURV2, 0u, fused ? 1 : 0, w->d_hit_bound, w->d_plan,
                w->d_anchors_a, anchor_cap, w->d_counters, D);
            launches++;
        } else {
            k_extend_hits<<<G.extend_grid, EXTEND_THREADS, 0, st>>>(
                P, g.d_sub_mat, w->d_hits, hits_cap, filter ? sv : nullptr, surv_cap, sv_ctr, decline, fused ? 1 : 0, w->d_hit_bound, w->d_plan,
                w->d_anchors_a, anchor_cap, w->d_counters, D);
        }
        pt.mark(PH_EXTEND);
        // 5. diagonal sort, dedupe, final order (seed_filter.cu:776-782): one block when the anchors
        //    fit; the counters, the plan and the first FINALIZE_CAP records come back together
        if (in.rm) k_finalize_small_rm<<<1, FINALIZE_THREADS, 0, st>>>(w->d_anchors_a, anchor_cap, w->d_out, w->d_counters, G.finalize_cap, rev, g.ref.len);
        else k_finalize_small<<<1, FINALIZE_THREADS, 0, st>>>(w->d_anchors_a, anchor_cap, w->d_out, w->d_counters, G.finalize_cap);
        pt.mark(PH_SORT);
        launches += 2;
        CU(cudaMemcpyAsync(w->h_small, w->d_counters, CTR_WORDS * sizeof(uint32_t), cudaMemcpyDeviceToHost, st), SA_ERR_MEMCPY);
        CU(cudaMemcpyAsync(w->h_small + 16, w->d_plan, 4 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st), SA_ERR_MEMCPY);
        CU(cudaMemcpyAsync(w->h_out, w->d_out, FINALIZE_CAP * sizeof(sa_segment), cudaMemcpyDeviceToHost, st), SA_ERR_MEMCPY);
        CU(cudaStreamSynchronize(st), SA_ERR_KERNEL);
        n_pre = w->h_small[CTR_ANCHORS];
        memcpy(&ext_cells, &w->h_small[CTR_EXT_LO], 8);
        memcpy(&num_hits64, &w->h_small[CTR_NHITS64], 8);
        if (in.rm && num_hits64 > 0xFFFFFFFFull)
            return fail(SA_ERR_STATE, "repeat-masker call with %llu seed hits: more than 2^32 hits per call are not supported",
                        (unsigned long long)num_hits64);
        if (fused) {
            num_hits = w->h_small[CTR_NHITS];
            num_seeds = w->h_small[CTR_NSEEDS];
            n_surv = w->h_small[CTR_SURV];
            n_walked = w->h_small[CTR_WALKED];
            if (num_hits >= G.max_hits) { // more than one iteration pair: the general path decides
                fused = false;
                continue;
            }
        } else {
            num_iter = w->h_small[16];
            num_hits = w->h_small[17];
            num_seeds = std::min(w->h_small[18], max_items);
            n_surv = filter ? w->h_small[CTR_SURV] : num_hits;
            if (num_iter == 0xFFFFFFFFu) { // more iterations than the plan arrays hold
                size_t need = (size_t)num_hits / G.max_hits + 2;
                TRY(ensure(w->d_limit_pos, w->limit_cap, need, "limit_pos"));
                TRY(ensure(w->d_hit_bound, w->bound_cap, need, "hit_bound"));
                continue;
            }
            if (num_hits > hits_cap) { // the hit list did not fit: grow and replay
                TRY(ensure(w->d_hits, w->hits_cap, num_hits, "hsp"));
                continue;
            }
        }
        if (filter && n_surv > surv_cap) { // the survivor list did not fit
            TRY(ensure(w->d_surv, w->surv_cap, n_surv, "survivors"));
            continue;
        }
        if (n_pre > anchor_cap) { // the anchor list did not fit
            TRY(ensure(w->d_anchors_a, w->anchors_a_cap, n_pre, "hsp_reduced"));
            continue;
        }
        if (decline && n_surv > decline) { // stage B declined the call: drop the provable copies, replay stage B
            TRY(enqueue_merge(w, P, n_surv));
            launches += 2 + 2 * 8;
            merged = true;
            continue;
        }
        break;
    }
    if (num_hits > 0) {
        if (w->h_small[CTR_OUT] != 0xFFFFFFFFu) {
            n_final = w->h_small[CTR_OUT];
            staged = true;
        } else if (in.rm) {
            // repeat_masker_src/seed_filter.cu:819-835 device-wide (kernels_sort.cuh: the three total orders)
            auto count_of = [&](uint32_t &n) -> int {
                CU(cudaMemcpyAsync(w->h_small, w->d_counters + CTR_DEDUPE, sizeof(uint32_t), cudaMemcpyDeviceToHost, st), SA_ERR_MEMCPY);
                CU(cudaStreamSynchronize(st), SA_ERR_KERNEL);
                n = w->h_small[0];
                return SA_OK;
            };
            if (rev) k_rm_to_forward<<<grid_for(n_pre, 256), 256, 0, st>>>(w->d_anchors_a, n_pre, g.ref.len);
            TRY(ensure(w->d_anchors_b, w->anchors_b_cap, n_pre, "hsp_unique"));
            TRY(radix_sort_anchors(w, w->d_anchors_a, w->d_anchors_b, n_pre, ORD_RM_FIRST));
            CU(cudaMemsetAsync(w->d_counters + CTR_DEDUPE, 0, sizeof(uint32_t), st), SA_ERR_MEMCPY);
            k_dedupe_exact<<<grid_for(n_pre, 256), 256, 0, st>>>(w->d_anchors_b, n_pre, w->d_anchors_a, w->d_counters + CTR_DEDUPE);
            uint32_t n1 = 0;
            TRY(count_of(n1));
            TRY(radix_sort_anchors(w, w->d_anchors_a, w->d_anchors_b, n1, ORD_RM_DIAG));
            CU(cudaMemsetAsync(w->d_counters + CTR_DEDUPE, 0, sizeof(uint32_t), st), SA_ERR_MEMCPY);
            k_dedupe<<<grid_for(n1, 256), 256, 0, st>>>(w->d_anchors_b, n1, w->d_anchors_a, w->d_counters + CTR_DEDUPE);
            TRY(count_of(n_final));
            TRY(radix_sort_anchors(w, w->d_anchors_a, w->d_anchors_b, n_final, ORD_RM_FINAL));
            TRY(ensure(w->d_out, w->out_cap, n_final, "hsp_out"));
            k_strip_tags<<<grid_for(n_final, 256), 256, 0, st>>>(w->d_anchors_b, n_final, w->d_out);
            launches += 3 * 32 + 5;
            pt.mark(PH_SORT);
        } else {
            // many distinct anchors (repeat families, forced small iterations): device-wide radix sorts
            TRY(ensure(w->d_anchors_b, w->anchors_b_cap, n_pre, "hsp_unique"));
            TRY(radix_sort_anchors(w, w->d_anchors_a, w->d_anchors_b, n_pre, ORD_DIAG));
            CU(cudaMemsetAsync(w->d_counters + CTR_DEDUPE, 0, sizeof(uint32_t), st), SA_ERR_MEMCPY);
            k_dedupe<<<grid_for(n_pre, 256), 256, 0, st>>>(w->d_anchors_b, n_pre, w->d_anchors_a, w->d_counters + CTR_DEDUPE);
            CU(cudaMemcpyAsync(w->h_small, w->d_counters + CTR_DEDUPE, sizeof(uint32_t), cudaMemcpyDeviceToHost, st), SA_ERR_MEMCPY);
            CU(cudaStreamSynchronize(st), SA_ERR_KERNEL);
            n_final = w->h_small[0];
            TRY(radix_sort_anchors(w, w->d_anchors_a, w->d_anchors_b, n_final, ORD_LASTZ));
            TRY(ensure(w->d_out, w->out_cap, n_final, "hsp_out"));
            k_strip_tags<<<grid_for(n_final, 256), 256, 0, st>>>(w->d_anchors_b, n_final, w->d_out);
            launches += 2 * 32 + 4;
            pt.mark(PH_SORT);
        }
    }
    // 6. result (seed_filter.cu:786-788, :804-822)
    sa_segment *res = (sa_segment *)malloc(((size_t)n_final + 1) * sizeof(sa_segment));
    if (!res) return fail(SA_ERR_MALLOC, "malloc of result failed");
    if (staged) {
        memcpy(res + 1, w->h_out, (size_t)n_final * sizeof(sa_segment));
    } else if (n_final > 0) {
        cudaError_t e = cudaMemcpyAsync(res + 1, w->d_out, (size_t)n_final * sizeof(sa_segment), cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) {
            free(res);
            return fail(SA_ERR_MEMCPY, "cudaMemcpy of %lu bytes for hsp_output failed with error \" %s \"",
                        (unsigned long)n_final * sizeof(sa_segment), cudaGetErrorString(e));
        }
    }
    if (in.rm) { // repeat_masker_src/seed_filter.cu:856-861: 64-bit totals, low word first
        res[0].ref_start = (uint32_t)(num_hits64 & 0xFFFFFFFFull);
        res[0].query_start = (uint32_t)(num_hits64 >> 32);
        res[0].len = n_final;
        res[0].score = 0;
    } else {
        res[0].ref_start = 0;
        res[0].query_start = 0;
        res[0].len = n_final;
        res[0].score = (int32_t)num_hits;
    }
    pt.mark(PH_D2H);
    if (pt.on) CU(cudaStreamSynchronize(st), SA_ERR_KERNEL);
    CU(cudaGetLastError(), SA_ERR_KERNEL);
    *out = res;
    *out_count = n_final + 1;
    if (out_num_seeds) *out_num_seeds = num_seeds;
    float ph_ms[PH_COUNT] = {};
    if (pt.on) // event queries stay outside the stats lock: concurrent callers must not serialise on them
        for (int p = PH_SEEDS; p < PH_COUNT; p++) ph_ms[p] = pt.ms((Phase)p);
    {
        std::lock_guard<std::mutex> l(G.stats_mu);
        sa_stats &s = G.stats;
        g.calls++;
        s.calls++; s.seeds += num_seeds; s.hits += num_hits; s.survivors += n_surv;
        s.anchors_pre_dedupe += n_pre; s.hsps += n_final; s.ext_cells += ext_cells;
        s.launches += launches;
        s.walked += n_walked;
        if (merged) { s.merge_calls++; s.merge_dropped += n_surv - std::min(n_surv, w->h_small[CTR_MERGED]); }
        if (pt.on) {
            s.ms_h2d += ph_ms[PH_SEEDS];
            s.ms_count_scan += ph_ms[PH_PLAN];
            s.ms_lookup += ph_ms[PH_LOOKUP];
            s.ms_prefilter += ph_ms[PH_FILTER];
            s.ms_extend += ph_ms[PH_EXTEND];
            s.ms_sort += ph_ms[PH_SORT];
            s.ms_d2h += ph_ms[PH_D2H];
        }
    }
    return SA_OK;
}

int check_call_state(uint32_t buffer) {
    if (!G.processor_ready) return fail(SA_ERR_STATE, "InitializeProcessor has not been called");
    if (!G.ref_loaded || !G.table_ready) return fail(SA_ERR_STATE, "no reference block / seed position table on the GPU");
    if (buffer >= SA_BUFFER_DEPTH || !G.query_loaded[buffer]) return fail(SA_ERR_STATE, "no query block in buffer %u", buffer);
    return SA_OK;
}

} // namespace

// ====================================================================== C ABI
extern "C" {

const char *sa_last_error(void) { return g_err.c_str(); }
const char *sa_version(void) { return "segalign_b200 0.1 (sm_100a)"; }

int sa_initialize_interface_at(int first_device, int num_gpu) {
    int n = 0;
    cudaError_t err = cudaGetDeviceCount(&n);
    if (err != cudaSuccess || n <= 0) return fail(SA_ERR_NO_GPU, "Error: No GPU device found!");
    if (first_device < 0 || first_device >= n) return fail(SA_ERR_ARG, "first_device %d out of range", first_device);
    int avail = n - first_device;
    int use = num_gpu == -1 ? avail : num_gpu;
    if (use > avail || use <= 0) return fail(SA_ERR_TOO_MANY_GPUS, "Requested GPUs greater than available GPUs");
    G.gpus.clear();
    G.gpus.resize(use);
    for (int i = 0; i < use; i++) {
        G.gpus[i].device = first_device + i;
        CU(cudaSetDevice(G.gpus[i].device), SA_ERR_SET_DEVICE);
        CU(cudaStreamCreateWithFlags(&G.gpus[i].ctrl, cudaStreamNonBlocking), SA_ERR_KERNEL);
        CU(cudaEventCreateWithFlags(&G.gpus[i].ev_up, cudaEventDisableTiming), SA_ERR_KERNEL);
        if (i == 0) CU(cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, G.gpus[i].device), SA_ERR_KERNEL);
        if (i > 0) { // block uploads reach this GPU by a peer copy from the first one: direct over NVLink when possible
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, G.gpus[i].device, G.gpus[0].device) == cudaSuccess && can) {
                cudaError_t pe = cudaDeviceEnablePeerAccess(G.gpus[0].device, 0);
                if (pe != cudaSuccess) (void)cudaGetLastError(); // already enabled / unsupported: cudaMemcpyPeerAsync stages instead
            }
        }
        {   // keep freed block memory in the pool (see alloc_planes)
            cudaMemPool_t pool;
            unsigned long long keep = ~0ull;
            CU(cudaDeviceGetDefaultMemPool(&pool, G.gpus[i].device), SA_ERR_MALLOC);
            CU(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep), SA_ERR_MALLOC);
        }
    }
    fprintf(stderr, "Using %d GPU(s)\n", use);
    G.interface_ready = true;
    return use;
}

int sa_initialize_interface(int num_gpu) { return sa_initialize_interface_at(0, num_gpu); }

int sa_initialize_processor(int transition, uint32_t wga_chunk, uint32_t seed_size,
                            const int *sub_mat, int xdrop, int hspthresh, int noentropy) {
    if (!G.interface_ready) return fail(SA_ERR_STATE, "InitializeInterface has not been called");
    if (!sub_mat) return fail(SA_ERR_ARG, "sub_mat is NULL");
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, G.gpus[0].device), SA_ERR_SET_DEVICE);
    // seed_filter.cu:834-841, same float arithmetic
    float global_mem_gb = static_cast<float>(prop.totalGlobalMem / 1073741824.0f);
    G.max_seeds = transition ? 13u * wga_chunk : wga_chunk;
    G.max_hits_device = (uint32_t)(int)(4194304 * global_mem_gb);
    G.max_hits = G.max_hits_device;
    G.transition = transition;
    G.seed_size = seed_size;
    G.xdrop = xdrop;
    G.hspthresh = hspthresh;
    G.noentropy = noentropy;
    memcpy(G.sub_mat, sub_mat, sizeof(G.sub_mat));
    G.diag_all_positive = sub_mat[0] > 0 && sub_mat[9] > 0 && sub_mat[18] > 0 && sub_mat[27] > 0;
    // kernels_filter.cuh preconditions: int8 ACGT block; terminator / soft classes of the non-ACGT codes
    // (screen_bound.h: screen_terminator_codes)
    G.filter_ok = true;
    for (int a = 0; a < 4; a++)
        for (int b = 0; b < 4; b++)
            if (sub_mat[a * 8 + b] < -128 || sub_mat[a * 8 + b] > 127) G.filter_ok = false;
    G.screen = screen_consts_from_matrix(sub_mat, xdrop, hspthresh);
    if (!G.filter_ok) G.screen.enabled = 0;
    G.term_codes = screen_terminator_codes(sub_mat, xdrop, &G.strict_term_codes);
    zero_run_codes(sub_mat, &G.zero_flat, &G.zero_partners);
    if (const char *e = getenv("SEGALIGN_B200_ZERO_RUNS")) if (atoi(e) == 0) G.zero_flat = G.zero_partners = 0;
    const char *fenv = getenv("SEGALIGN_B200_FILTER");
    G.use_filter = !(fenv && atoi(fenv) == 0);
    const char *denv = getenv("SEGALIGN_B200_DEDUP");
    G.use_dedup = !(denv && atoi(denv) == 0);
    const char *uenv = getenv("SEGALIGN_B200_FUSED");
    G.use_fused = !(uenv && atoi(uenv) == 0);
    const char *env = getenv("SEGALIGN_B200_STREAMS");
    if (env && atoi(env) > 0) G.ws_per_gpu = atoi(env);
    for (size_t i = 0; i < G.gpus.size(); i++) {
        GpuCtx &g = G.gpus[i];
        CU(cudaSetDevice(g.device), SA_ERR_SET_DEVICE);
        CU(cudaMalloc((void **)&g.d_sub_mat, 64 * sizeof(int)), SA_ERR_MALLOC);
        CU(cudaMemcpy(g.d_sub_mat, sub_mat, 64 * sizeof(int), cudaMemcpyHostToDevice), SA_ERR_MEMCPY);
        int sms = 0;
        CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, g.device), SA_ERR_KERNEL);
        int per_sm2 = 0;
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm2, k_filter_hits2<SRC_RANGE>, FILTER_THREADS,
                                                         FILTER_LUT_WORDS * sizeof(uint32_t)), SA_ERR_KERNEL);
        if (const char *e = getenv("SEGALIGN_B200_FILTER_CTAS")) if (atoi(e) > 0) per_sm2 = std::min(per_sm2, atoi(e));
        G.filter2_grid = std::max(1, per_sm2) * std::max(1, sms);
        CU(cudaFuncSetAttribute(k_filter_hits3<SRC_RANGE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SCR_SMEM_BYTES), SA_ERR_KERNEL);
        CU(cudaFuncSetAttribute(k_filter_hits3<SRC_SEEDS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SCR_SMEM_BYTES), SA_ERR_KERNEL);
        int per_sm3 = 0;
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm3, k_filter_hits3<SRC_RANGE>, SCR_THREADS, SCR_SMEM_BYTES), SA_ERR_KERNEL);
        // All but one of the block slots per SM by default (2 of 3): the last is left to the next call's kernel
        // (another stream), whose head overlaps this one's tail -- 8 % more calls per second with 16
        // callers than three blocks per SM, at 6 % more time for a launch that runs alone.
        G.filter3_grid_alone = std::max(1, per_sm3) * std::max(1, sms);
        per_sm3 = std::max(1, per_sm3 - 1);
        if (const char *e = getenv("SEGALIGN_B200_FILTER_CTAS")) if (atoi(e) > 0) { per_sm3 = atoi(e); G.filter3_grid_alone = per_sm3 * std::max(1, sms); }
        G.filter3_grid = std::max(1, per_sm3) * std::max(1, sms);
        G.filter_kernel = 3;
        if (const char *e = getenv("SEGALIGN_B200_FILTER_KERNEL")) { int v = atoi(e); if (v == 2 || v == 3) G.filter_kernel = v; }
        G.extend_grid = 8 * std::max(1, sms); // one-warp blocks, persistent over the work list
        if (const char *e = getenv("SEGALIGN_B200_EXTEND_CTAS")) if (atoi(e) > 0) G.extend_grid = atoi(e) * std::max(1, sms);
        G.wide_grid = 4 * std::max(1, sms);   // four-warp blocks, one warp per hit
        if (const char *e = getenv("SEGALIGN_B200_WIDE_CTAS")) if (atoi(e) > 0) G.wide_grid = atoi(e) * std::max(1, sms);
        { const char *e = getenv("SEGALIGN_B200_WIDE"); G.use_wide = !(e && atoi(e) == 0); }
        { const char *e = getenv("SEGALIGN_B200_FINALIZE_CAP"); G.finalize_cap = e ? std::min<uint32_t>((uint32_t)FINALIZE_CAP, (uint32_t)strtoul(e, nullptr, 10)) : (uint32_t)FINALIZE_CAP; }
        { const char *e = getenv("SEGALIGN_B200_MERGE_MIN"); G.merge_min = e ? (uint32_t)strtoul(e, nullptr, 10) : 65536u; }
        { const char *e = getenv("SEGALIGN_B200_COMPACT_SEEDS"); G.use_compact = !(e && atoi(e) == 0); }
        // blocks uploaded before the matrix was known carry records built for another terminator set
        SeqPlanes *all[] = {&g.ref, &g.q_fwd[0], &g.q_rc[0], &g.q_fwd[1], &g.q_rc[1]};
        for (SeqPlanes *p : all)
            if (p->rec && (p->term_codes != G.term_codes || p->zero_codes != (G.zero_flat | (G.zero_partners << 8)))) TRY(build_records(g, *p));
        CU(cudaStreamSynchronize(g.ctrl), SA_ERR_KERNEL);
        for (SeqPlanes *p : all)
            if (p->rec) TRY(read_soft_flag(*p));
        for (int k = 0; k < G.ws_per_gpu; k++) {
            Workspace *w = nullptr;
            TRY(make_workspace((int)i, w));
            g.ws.push_back(w);
        }
    }
    {
        std::lock_guard<std::mutex> lk(G.mu);
        G.free_ws.clear();
        // interleave so consecutive calls spread over the GPUs first
        for (int k = 0; k < G.ws_per_gpu; k++)
            for (size_t i = 0; i < G.gpus.size(); i++) G.free_ws.push_back(G.gpus[i].ws[k]);
        std::reverse(G.free_ws.begin(), G.free_ws.end());
    }
    G.processor_ready = true;
    return SA_OK;
}

int sa_set_max_hits(uint32_t max_hits) {
    G.max_hits = max_hits ? max_hits : G.max_hits_device;
    return SA_OK;
}
uint32_t sa_get_max_hits(void) { return G.max_hits; }
int sa_set_filter_kernel(int kernel) {
    const int prev = G.filter_kernel;
    G.filter_kernel = kernel == 2 ? 2 : 3;
    return prev;
}

int sa_set_seed_shape(const char *pattern) {
    if (!pattern) return fail(SA_ERR_ARG, "pattern is NULL");
    size_t n = strlen(pattern);
    if (n == 0 || n > 32) return fail(SA_ERR_ARG, "seed span must be 1..32");
    ShapeDesc sh = {};
    sh.span = (int)n;
    for (size_t i = 0; i < n; i++) { // ntcoding.cpp:21-37
        if (pattern[i] == '1' || pattern[i] == 'T') {
            sh.pos[sh.weight] = (uint8_t)i;
            sh.trans[sh.weight] = pattern[i] == 'T';
            if (pattern[i] == 'T') sh.tvar[sh.num_trans++] = (uint8_t)sh.weight;
            sh.weight++;
        }
    }
    if (sh.weight <= 3 || sh.weight > 15) return fail(SA_ERR_ARG, "seed weight must be 4..15"); // seed_pos_table.cu:51-52
    G.shape = sh;
    G.shape_set = true;
    return sh.weight;
}

int sa_send_ref(const char *seq, size_t start_addr, uint32_t len) {
    if (!G.interface_ready) return fail(SA_ERR_STATE, "InitializeInterface has not been called");
    if (G.ref_loaded) return fail(SA_ERR_STATE, "ClearRef must precede a second SendRefWriteRequest");
    if (!seq && len) return fail(SA_ERR_ARG, "seq is NULL");
    G.ref_len = len;
    const auto t0 = std::chrono::steady_clock::now();
    TRY(upload_block_all_gpus(seq + start_addr, len, -1, "ref_seq"));
    G.ascii_holds_ref = true;
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    {
        std::lock_guard<std::mutex> l(G.stats_mu);
        G.stats.ms_ref_encode += ms; // wall time of the whole pool (the GPUs work concurrently)
    }
    G.ref_loaded = true;
    return SA_OK;
}

namespace {
// device buffers of one GPU's table build; whatever is still owned is freed on every exit path
struct TableBuild {
    uint32_t *keys_a = nullptr, *keys_b = nullptr, *vals_a = nullptr, *vals_b = nullptr, *index = nullptr;
    void *temp = nullptr;
    uint32_t *num_pos_host = nullptr; // pinned
    int device = 0;
    ~TableBuild() {
        cudaSetDevice(device);
        cudaFree(keys_a); cudaFree(keys_b); cudaFree(vals_a); cudaFree(vals_b); cudaFree(index); cudaFree(temp);
        if (num_pos_host) cudaFreeHost(num_pos_host);
    }
};
} // namespace

int sa_generate_seed_pos_table(const char *ref_str, size_t start_addr, uint32_t ref_length,
                               uint32_t step, int shape_size, int kmer_size) {
    (void)ref_str; (void)start_addr;
    if (!G.ref_loaded) return fail(SA_ERR_STATE, "SendRefWriteRequest must precede GenerateSeedPosTable");
    if (!G.shape_set) return fail(SA_ERR_STATE, "seed shape not set (GenerateShapePos)");
    if (ref_length != G.ref_len) return fail(SA_ERR_ARG, "ref_length %u differs from the resident block (%u)", ref_length, G.ref_len);
    if (shape_size != G.shape.span || kmer_size != G.shape.weight) return fail(SA_ERR_ARG, "shape_size/kmer_size do not match the seed shape");
    if (step == 0) return fail(SA_ERR_ARG, "step must be > 0");
    if (G.table_ready) return fail(SA_ERR_STATE, "ClearRef must precede a second GenerateSeedPosTable");
    // seed_pos_table.cu:58-64
    uint32_t offset = ((uint32_t)shape_size + 1) % step;
    uint32_t start_offset = step - offset;
    uint32_t num_steps = ref_length >= (uint32_t)shape_size ? (ref_length - shape_size + offset) / step : 0;
    uint32_t index_size = 1u << (2 * kmer_size);
    int end_bit = 2 * kmer_size + 1;
    const auto t0 = std::chrono::steady_clock::now();
    const size_t n = std::max<uint32_t>(num_steps, 1);
    const size_t ngpu = G.gpus.size();
    // Every GPU builds its own copy from its resident encoded block (no table crosses PCIe or NVLink;
    // the reference builds on the host and uploads ~4 bytes per base to every GPU in turn,
    // common/seed_pos_table.cu:33-47).  Phase 1 enqueues the whole build on every control stream,
    // phase 2 waits: the GPUs run concurrently.
    std::vector<TableBuild> tb(ngpu);
    std::vector<cub::DoubleBuffer<uint32_t>> dks(ngpu), dvs(ngpu);
    uint64_t launches = 0;
    for (size_t i = 0; i < ngpu; i++) {
        GpuCtx &g = G.gpus[i];
        TableBuild &t = tb[i];
        t.device = g.device;
        CU(cudaSetDevice(g.device), SA_ERR_SET_DEVICE);
        cudaStream_t st = g.ctrl;
        CU(cudaMalloc((void **)&t.index, (size_t)index_size * sizeof(uint32_t)), SA_ERR_MALLOC);
        CU(cudaMemsetAsync(t.index, 0, (size_t)index_size * sizeof(uint32_t), st), SA_ERR_MEMCPY);
        CU(cudaMalloc((void **)&t.keys_a, n * 4), SA_ERR_MALLOC);
        CU(cudaMalloc((void **)&t.keys_b, n * 4), SA_ERR_MALLOC);
        CU(cudaMalloc((void **)&t.vals_a, n * 4), SA_ERR_MALLOC);
        CU(cudaMalloc((void **)&t.vals_b, n * 4), SA_ERR_MALLOC);
        CU(cudaMallocHost((void **)&t.num_pos_host, sizeof(uint32_t)), SA_ERR_MALLOC);
        if (num_steps > 0) {
            k_table_keys<<<grid_for(num_steps, 256), 256, 0, st>>>(g.ref.p2, g.ref.m1, G.shape, start_offset, step,
                                                                   num_steps, t.keys_a, t.vals_a, t.index);
            launches++;
        }
        // histogram -> inclusive end offsets (seed_pos_table.cu:83; device sees index_table+1)
        size_t bytes = 0, bytes2 = 0;
        CU(cub::DeviceScan::InclusiveSum(nullptr, bytes, t.index, t.index, (int)index_size, st), SA_ERR_KERNEL);
        dks[i] = cub::DoubleBuffer<uint32_t>(t.keys_a, t.keys_b);
        dvs[i] = cub::DoubleBuffer<uint32_t>(t.vals_a, t.vals_b);
        CU(cub::DeviceRadixSort::SortPairs(nullptr, bytes2, dks[i], dvs[i], (int)num_steps, 0, end_bit, st), SA_ERR_KERNEL);
        CU(cudaMalloc(&t.temp, std::max(bytes, bytes2) + 256), SA_ERR_MALLOC);
        CU(cub::DeviceScan::InclusiveSum(t.temp, bytes, t.index, t.index, (int)index_size, st), SA_ERR_KERNEL);
        launches += 2;
        CU(cudaMemcpyAsync(t.num_pos_host, t.index + index_size - 1, 4, cudaMemcpyDeviceToHost, st), SA_ERR_MEMCPY);
        if (num_steps > 0) {
            CU(cub::DeviceRadixSort::SortPairs(t.temp, bytes2, dks[i], dvs[i], (int)num_steps, 0, end_bit, st), SA_ERR_KERNEL);
            launches += 2 * ((end_bit + 7) / 8) + 1;
        }
        CU(cudaGetLastError(), SA_ERR_KERNEL);
    }
    for (size_t i = 0; i < ngpu; i++) {
        GpuCtx &g = G.gpus[i];
        TableBuild &t = tb[i];
        CU(cudaSetDevice(g.device), SA_ERR_SET_DEVICE);
        CU(cudaStreamSynchronize(g.ctrl), SA_ERR_KERNEL);
        // the sorted value buffer IS the position table (positions of invalid words sort behind num_pos)
        uint32_t *pos = dvs[i].Current();
        if (pos == t.vals_a) t.vals_a = nullptr; else t.vals_b = nullptr;
        g.d_pos = pos;
        g.d_index = t.index; t.index = nullptr;
        g.index_size = index_size;
        g.num_pos = *t.num_pos_host;
    }
    tb.clear(); // frees the key buffers, the spare value buffer and the temporaries
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    {
        std::lock_guard<std::mutex> l(G.stats_mu);
        G.stats.ms_table_build += ms; // wall time of the whole pool
        G.stats.launches += launches;
    }
    G.table_ready = true;
    return SA_OK;
}

int sa_clear_ref(void) {
    for (auto &g : G.gpus) {
        CU(cudaSetDevice(g.device), SA_ERR_SET_DEVICE);
        TRY(free_planes(g.ref, g.ctrl));
        if (g.d_index) CU(cudaFree(g.d_index), SA_ERR_FREE);
        if (g.d_pos) CU(cudaFree(g.d_pos), SA_ERR_FREE);
        g.d_index = g.d_pos = nullptr;
        g.index_size = g.num_pos = 0;
    }
    G.ref_loaded = G.table_ready = false;
    G.ascii_holds_ref = false;
    return SA_OK;
}

int sa_send_query(const char *query_base, size_t start_addr, uint32_t len, uint32_t buffer) {
    if (!G.interface_ready) return fail(SA_ERR_STATE, "InitializeInterface has not been called");
    if (buffer >= SA_BUFFER_DEPTH) return fail(SA_ERR_ARG, "buffer %u out of range", buffer);
    if (G.query_loaded[buffer]) return fail(SA_ERR_STATE, "ClearQuery(%u) must precede a refill", buffer);
    if (!query_base && len) return fail(SA_ERR_ARG, "query_base is NULL");
    G.query_len[buffer] = len;
    const auto t0 = std::chrono::steady_clock::now();
    G.ascii_holds_ref = false;
    TRY(upload_block_all_gpus(query_base + start_addr, len, (int)buffer, "query_seq"));
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    {
        std::lock_guard<std::mutex> l(G.stats_mu);
        G.stats.ms_query_encode += ms;
        G.stats.h2d_bytes += len; // the block crosses PCIe once; the other GPUs get it by peer copy
    }
    G.query_loaded[buffer] = true;
    return SA_OK;
}

int sa_clear_query(uint32_t buffer) {
    if (buffer >= SA_BUFFER_DEPTH) return fail(SA_ERR_ARG, "buffer %u out of range", buffer);
    for (auto &g : G.gpus) {
        CU(cudaSetDevice(g.device), SA_ERR_SET_DEVICE);
        TRY(free_planes(g.q_fwd[buffer], g.ctrl));
        TRY(free_planes(g.q_rc[buffer], g.ctrl));
    }
    G.query_loaded[buffer] = false;
    if (buffer == 0) G.rm_query = false;
    return SA_OK;
}

int sa_seed_and_filter(const uint64_t *seeds, uint32_t num_seeds, int rev, uint32_t buffer,
                       sa_segment **out, uint32_t *out_count) {
    if (!out || !out_count) return fail(SA_ERR_ARG, "out/out_count is NULL");
    TRY(check_call_state(buffer));
    if (num_seeds > G.max_seeds) { // seed_filter.cu:688-692
        printf("MAX_SEEDS exceeded\n");
        return fail(SA_ERR_MAX_SEEDS, "num_seeds %u > MAX_SEEDS %u", num_seeds, G.max_seeds);
    }
    if (num_seeds == 0) { // the reference's seeder never makes this call (seeder.cpp:76)
        sa_segment *res = (sa_segment *)calloc(1, sizeof(sa_segment));
        *out = res; *out_count = 1;
        return SA_OK;
    }
    if (!seeds) return fail(SA_ERR_ARG, "seeds is NULL");
    Workspace *w = acquire_ws();
    WsGuard guard(w);
    CU(cudaSetDevice(G.gpus[w->gpu].device), SA_ERR_SET_DEVICE);
    PhaseTimer pt(w);
    pt.mark(PH_START);
    TRY(ensure(w->d_seeds, w->seeds_cap, num_seeds, "seed_offsets"));
    TRY(ensure(w->d_prefix, w->prefix_cap, num_seeds, "hit_num"));
    // A seed vector in the seeder's own form (per position: exact word, then its transition variants,
    // src/seeder.cpp:57-74) is uploaded as its base words only and rebuilt on the device; any other
    // vector is copied as it is.  One pass over the caller's vector either way.
    const uint32_t per = 1u + (G.transition ? (uint32_t)G.shape.num_trans : 0u);
    bool compact = G.use_compact && per > 1 && num_seeds >= 4096 && num_seeds % per == 0;
    if (compact) {
        const uint32_t groups = num_seeds / per;
        if (w->h_bases_cap < groups) {
            if (w->h_bases) cudaFreeHost(w->h_bases);
            w->h_bases = nullptr; w->h_bases_cap = 0;
            const size_t cap = std::max<size_t>(groups, G.max_seeds / per + 1);
            CU(cudaMallocHost((void **)&w->h_bases, cap * sizeof(uint64_t)), SA_ERR_MALLOC);
            w->h_bases_cap = cap;
        }
        VariantMasks VM = {};
        for (uint32_t v = 1; v < per; v++) VM.xm[v] = (uint64_t)(2u << (2 * G.shape.tvar[v - 1])) << 32;
        uint64_t bad = 0;
        for (uint32_t gi = 0; gi < groups && !bad; gi++) {
            const uint64_t *p = seeds + (size_t)gi * per;
            const uint64_t base = p[0];
            for (uint32_t v = 1; v < per; v++) bad |= p[v] ^ base ^ VM.xm[v];
            w->h_bases[gi] = base;
        }
        compact = bad == 0;
        if (compact) {
            TRY(ensure(w->d_bases, w->d_bases_cap, groups, "seed_bases"));
            CU(cudaMemcpyAsync(w->d_bases, w->h_bases, (size_t)groups * sizeof(uint64_t), cudaMemcpyHostToDevice, w->stream), SA_ERR_MEMCPY);
            k_expand_bases<<<grid_for(num_seeds, 256), 256, 0, w->stream>>>(w->d_bases, num_seeds, per, VM, w->d_seeds);
            std::lock_guard<std::mutex> l(G.stats_mu);
            G.stats.h2d_bytes += (uint64_t)groups * sizeof(uint64_t);
            G.stats.launches += 1;
        }
    }
    if (!compact) {
        CU(cudaMemcpyAsync(w->d_seeds, seeds, (size_t)num_seeds * sizeof(uint64_t), cudaMemcpyHostToDevice, w->stream), SA_ERR_MEMCPY);
        std::lock_guard<std::mutex> l(G.stats_mu);
        G.stats.h2d_bytes += (uint64_t)num_seeds * sizeof(uint64_t);
    }
    w->h_small[20] = num_seeds; // pinned: the seed count travels to d_plan[2] on the stream
    CU(cudaMemcpyAsync(w->d_plan + 2, w->h_small + 20, sizeof(uint32_t), cudaMemcpyHostToDevice, w->stream), SA_ERR_MEMCPY);
    pt.mark(PH_SEEDS);
    CallInput in = {};
    in.src = SRC_SEEDS; in.max_items = num_seeds; in.per = 1; in.transition = G.transition;
    return run_pipeline(w, in, rev, buffer, out, out_count, nullptr, pt);
}

int sa_seed_and_filter_range(uint32_t q_start, uint32_t q_end, int transition, int rev,
                             uint32_t buffer, sa_segment **out, uint32_t *out_count,
                             uint32_t *out_num_seeds) {
    if (!out || !out_count) return fail(SA_ERR_ARG, "out/out_count is NULL");
    TRY(check_call_state(buffer));
    if (q_end < q_start) return fail(SA_ERR_ARG, "q_end < q_start");
    Workspace *w = acquire_ws();
    WsGuard guard(w);
    GpuCtx &g = G.gpus[w->gpu];
    CU(cudaSetDevice(g.device), SA_ERR_SET_DEVICE);
    const SeqPlanes &q = rev ? g.q_rc[buffer] : g.q_fwd[buffer];
    if (q_end > q.len) return fail(SA_ERR_ARG, "range [%u,%u) exceeds the query block (%u)", q_start, q_end, q.len);
    PhaseTimer pt(w);
    pt.mark(PH_START);
    const uint32_t n = q_end - q_start;
    const uint32_t per = 1u + (transition ? (uint32_t)G.shape.num_trans : 0u);
    const uint64_t max_items64 = (uint64_t)n * per;
    if (n == 0) {
        if (out_num_seeds) *out_num_seeds = 0;
        sa_segment *res = (sa_segment *)calloc(1, sizeof(sa_segment));
        *out = res; *out_count = 1;
        return SA_OK;
    }
    if (max_items64 > G.max_seeds) { // a full range could exceed MAX_SEEDS (seed_filter.cu:688-692)
        printf("MAX_SEEDS exceeded\n");
        return fail(SA_ERR_MAX_SEEDS, "range of %u positions x %u words > MAX_SEEDS %u", n, per, G.max_seeds);
    }
    CallInput in = {};
    in.src = SRC_RANGE; in.max_items = (uint32_t)max_items64;
    in.q_start = q_start; in.q_end = q_end; in.per = per; in.transition = transition;
    pt.mark(PH_SEEDS);
    return run_pipeline(w, in, rev, buffer, out, out_count, out_num_seeds, pt);
}

// ---------------------------------------------------------------- repeat-masker variant
// repeat_masker_src/seed_filter.cu:951-961 SendQueryWriteRequest(): the "query" is the resident block itself
// (plus strand) and its reverse complement, built on the device (minus strand).  Here both come out of one
// more encode pass over the block's ASCII bytes, which are still in the per-GPU staging buffers.
int sa_rm_send_query(void) {
    if (!G.ref_loaded) return fail(SA_ERR_STATE, "SendRefWriteRequest must precede the repeat masker's SendQueryWriteRequest");
    if (!G.ascii_holds_ref) return fail(SA_ERR_STATE, "the reference block's bytes are no longer staged: call right after SendRefWriteRequest");
    if (G.query_loaded[0]) return fail(SA_ERR_STATE, "ClearQuery must precede a second SendQueryWriteRequest");
    for (auto &g : G.gpus) {
        CU(cudaSetDevice(g.device), SA_ERR_SET_DEVICE);
        TRY(enqueue_encode(g, G.ref_len, g.q_fwd[0], &g.q_rc[0], "seq_rc"));
    }
    for (auto &g : G.gpus) {
        CU(cudaSetDevice(g.device), SA_ERR_SET_DEVICE);
        CU(cudaStreamSynchronize(g.ctrl), SA_ERR_KERNEL);
        TRY(read_soft_flag(g.q_fwd[0]));
        g.q_rc[0].has_flat = g.q_fwd[0].has_flat;
    }
    G.query_len[0] = G.ref_len;
    G.query_loaded[0] = true;
    G.rm_query = true;
    return SA_OK;
}
int sa_rm_clear_query(void) { return sa_clear_query(0); }

static int rm_check(uint32_t ref_start, uint32_t ref_end) {
    TRY(check_call_state(0));
    if (!G.rm_query) return fail(SA_ERR_STATE, "query slot 0 does not hold the block's own reverse complement (sa_rm_send_query)");
    if (ref_end < ref_start) return fail(SA_ERR_ARG, "ref_end < ref_start");
    return SA_OK;
}

int sa_rm_seed_and_filter(const uint64_t *seeds, uint32_t num_seeds, int rev, uint32_t ref_start, uint32_t ref_end,
                          sa_segment **out, uint32_t *out_count) {
    if (!out || !out_count) return fail(SA_ERR_ARG, "out/out_count is NULL");
    TRY(rm_check(ref_start, ref_end));
    if (num_seeds > G.max_seeds) { // repeat_masker_src/seed_filter.cu:730-734
        printf("MAX_SEEDS exceeded\n");
        return fail(SA_ERR_MAX_SEEDS, "num_seeds %u > MAX_SEEDS %u", num_seeds, G.max_seeds);
    }
    if (num_seeds == 0) {
        sa_segment *res = (sa_segment *)calloc(1, sizeof(sa_segment));
        *out = res; *out_count = 1;
        return SA_OK;
    }
    if (!seeds) return fail(SA_ERR_ARG, "seeds is NULL");
    Workspace *w = acquire_ws();
    WsGuard guard(w);
    CU(cudaSetDevice(G.gpus[w->gpu].device), SA_ERR_SET_DEVICE);
    PhaseTimer pt(w);
    pt.mark(PH_START);
    TRY(ensure(w->d_seeds, w->seeds_cap, num_seeds, "seed_offsets"));
    TRY(ensure(w->d_prefix, w->prefix_cap, num_seeds, "hit_num"));
    CU(cudaMemcpyAsync(w->d_seeds, seeds, (size_t)num_seeds * sizeof(uint64_t), cudaMemcpyHostToDevice, w->stream), SA_ERR_MEMCPY);
    {
        std::lock_guard<std::mutex> l(G.stats_mu);
        G.stats.h2d_bytes += (uint64_t)num_seeds * sizeof(uint64_t);
    }
    w->h_small[20] = num_seeds;
    CU(cudaMemcpyAsync(w->d_plan + 2, w->h_small + 20, sizeof(uint32_t), cudaMemcpyHostToDevice, w->stream), SA_ERR_MEMCPY);
    pt.mark(PH_SEEDS);
    CallInput in = {};
    in.src = SRC_SEEDS; in.max_items = num_seeds; in.per = 1; in.transition = G.transition;
    in.rm = true; in.win_lo = ref_start; in.win_hi = ref_end;
    return run_pipeline(w, in, rev, 0, out, out_count, nullptr, pt);
}

int sa_rm_seed_and_filter_range(uint32_t q_start, uint32_t q_end, int transition, int rev, uint32_t ref_start,
                                uint32_t ref_end, sa_segment **out, uint32_t *out_count, uint32_t *out_num_seeds) {
    if (!out || !out_count) return fail(SA_ERR_ARG, "out/out_count is NULL");
    TRY(rm_check(ref_start, ref_end));
    if (q_end < q_start) return fail(SA_ERR_ARG, "q_end < q_start");
    if (q_end > G.ref_len) return fail(SA_ERR_ARG, "range [%u,%u) exceeds the block (%u)", q_start, q_end, G.ref_len);
    const uint32_t n = q_end - q_start;
    const uint32_t per = 1u + (transition ? (uint32_t)G.shape.num_trans : 0u);
    if (n == 0) {
        if (out_num_seeds) *out_num_seeds = 0;
        sa_segment *res = (sa_segment *)calloc(1, sizeof(sa_segment));
        *out = res; *out_count = 1;
        return SA_OK;
    }
    if ((uint64_t)n * per > G.max_seeds) {
        printf("MAX_SEEDS exceeded\n");
        return fail(SA_ERR_MAX_SEEDS, "range of %u positions x %u words > MAX_SEEDS %u", n, per, G.max_seeds);
    }
    Workspace *w = acquire_ws();
    WsGuard guard(w);
    CU(cudaSetDevice(G.gpus[w->gpu].device), SA_ERR_SET_DEVICE);
    PhaseTimer pt(w);
    pt.mark(PH_START);
    CallInput in = {};
    in.src = SRC_RANGE; in.max_items = n * per;
    in.q_start = q_start; in.q_end = q_end; in.per = per; in.transition = transition;
    in.rm = true; in.win_lo = ref_start; in.win_hi = ref_end;
    pt.mark(PH_SEEDS);
    return run_pipeline(w, in, rev, 0, out, out_count, out_num_seeds, pt);
}

void sa_release_result(sa_segment *out) { free(out); }

int sa_shutdown_processor(void) {
    for (auto &g : G.gpus) {
        cudaSetDevice(g.device);
        for (auto *w : g.ws) destroy_workspace(w);
        g.ws.clear();
        free_planes(g.ref, g.ctrl);
        for (int b = 0; b < SA_BUFFER_DEPTH; b++) { free_planes(g.q_fwd[b], g.ctrl); free_planes(g.q_rc[b], g.ctrl); }
        if (g.ctrl) cudaStreamSynchronize(g.ctrl);
        { cudaMemPool_t pool; if (cudaDeviceGetDefaultMemPool(&pool, g.device) == cudaSuccess) cudaMemPoolTrimTo(pool, 0); }
        cudaFree(g.d_index); cudaFree(g.d_pos); cudaFree(g.d_sub_mat); cudaFree(g.d_ascii);
        g.d_index = g.d_pos = nullptr; g.d_sub_mat = nullptr; g.d_ascii = nullptr; g.ascii_cap = 0;
        if (g.ev_up) cudaEventDestroy(g.ev_up);
        g.ev_up = nullptr;
        if (g.ctrl) cudaStreamDestroy(g.ctrl);
        g.ctrl = nullptr;
    }
    {
        std::lock_guard<std::mutex> lk(G.mu);
        G.free_ws.clear();
    }
    G.gpus.clear();
    G.interface_ready = G.processor_ready = G.ref_loaded = G.table_ready = false;
    G.ascii_holds_ref = G.rm_query = false;
    for (int b = 0; b < SA_BUFFER_DEPTH; b++) G.query_loaded[b] = false;
    return SA_OK;
}

int sa_debug_get_table(uint32_t *index_size, uint32_t *num_pos, uint32_t *index_out, uint32_t *pos_out) {
    if (!G.table_ready) return fail(SA_ERR_STATE, "no seed position table");
    GpuCtx &g = G.gpus[0];
    CU(cudaSetDevice(g.device), SA_ERR_SET_DEVICE);
    if (index_size) *index_size = g.index_size;
    if (num_pos) *num_pos = g.num_pos;
    if (index_out) CU(cudaMemcpy(index_out, g.d_index, (size_t)g.index_size * 4, cudaMemcpyDeviceToHost), SA_ERR_MEMCPY);
    if (pos_out && g.num_pos) CU(cudaMemcpy(pos_out, g.d_pos, (size_t)g.num_pos * 4, cudaMemcpyDeviceToHost), SA_ERR_MEMCPY);
    return SA_OK;
}

int sa_debug_get_encoded(int which, uint32_t buffer, uint8_t *out, uint32_t len) {
    if (!G.interface_ready || G.gpus.empty()) return fail(SA_ERR_STATE, "not initialised");
    GpuCtx &g = G.gpus[0];
    CU(cudaSetDevice(g.device), SA_ERR_SET_DEVICE);
    const SeqPlanes *p = nullptr;
    if (which == 0) p = &g.ref;
    else if (buffer < SA_BUFFER_DEPTH) p = which == 1 ? &g.q_fwd[buffer] : &g.q_rc[buffer];
    if (!p || !p->b8 || len > p->len) return fail(SA_ERR_ARG, "no such encoded block");
    CU(cudaMemcpy(out, p->b8, len, cudaMemcpyDeviceToHost), SA_ERR_MEMCPY);
    return SA_OK;
}

int sa_get_stats(sa_stats *out) {
    if (!out) return fail(SA_ERR_ARG, "out is NULL");
    std::lock_guard<std::mutex> l(G.stats_mu);
    *out = G.stats;
    return SA_OK;
}
int sa_reset_stats(void) {
    std::lock_guard<std::mutex> l(G.stats_mu);
    G.stats = sa_stats();
    for (auto &g : G.gpus) g.calls = 0;
    return SA_OK;
}
int sa_set_profiling(int enabled) { G.profiling = enabled != 0; return SA_OK; }

int sa_get_gpu_calls(uint64_t *out, int cap) {
    if (!out || cap < 0) return fail(SA_ERR_ARG, "out is NULL");
    std::lock_guard<std::mutex> l(G.stats_mu);
    const int n = (int)G.gpus.size();
    for (int i = 0; i < n && i < cap; i++) out[i] = G.gpus[i].calls;
    return n;
}

// Host-side seed words of one chunk, the loop of src/seeder.cpp:57-74 with
// GetKmerIndexAtPos (common/ntcoding.cpp:43-61) restated as a rolling validity window.
// Pure host code (no CUDA): it exists so that callers that still hand over seed vectors
// (the reference ABI) can build them at memory speed from many threads.
static const struct NtLut { uint8_t v[256]; NtLut() { memset(v, 4, 256); v['A'] = 0; v['C'] = 1; v['G'] = 2; v['T'] = 3; } } g_nt_lut;

#if defined(__x86_64__)
// BMI2 path: the span is kept as a rolling 2-bit window (newest base in the low bits, so that the
// first care position ends up most significant) and the k-mer is one PEXT of it.
__attribute__((target("bmi2"))) static size_t host_chunk_seeds_bmi2(const unsigned char *s, uint32_t j0, uint32_t j1,
                                                                    const ShapeDesc &sh, int transition, uint64_t *out) {
    const int span = sh.span, w = sh.weight;
    uint64_t care = 0;
    for (int i = 0; i < w; i++) care |= 3ull << (2 * (span - 1 - sh.pos[i]));
    const uint64_t keep = span >= 32 ? ~0ull : ((1ull << (2 * span)) - 1ull);
    uint64_t xm[32];
    int nv = 0;
    if (transition)
        for (int t = 0; t < w; t++)
            if (sh.trans[t]) xm[nv++] = ((uint64_t)2 << (2 * t)) << 32;
    uint64_t win = 0;
    int64_t last_bad = -1;
    for (uint32_t i = j0; i + 1 < j0 + (uint32_t)span; i++) {
        uint32_t c = g_nt_lut.v[s[i]];
        if (c > 3) { last_bad = i; c = 0; }
        win = ((win << 2) | c) & keep;
    }
    size_t n = 0;
    for (uint32_t j = j0; j < j1; j++) {
        const uint32_t tail = j + (uint32_t)span - 1;
        uint32_t c = g_nt_lut.v[s[tail]];
        if (c > 3) { last_bad = tail; c = 0; }
        win = ((win << 2) | c) & keep;
        if (last_bad >= (int64_t)j) continue;
        const uint64_t base = (__builtin_ia32_pext_di(win, care) << 32) + j;
        out[n++] = base;
        for (int v = 0; v < nv; v++) out[n++] = base ^ xm[v];
    }
    return n;
}
#endif

size_t sa_host_chunk_seeds(const char *seq, size_t block_start, uint32_t j0, uint32_t j1,
                           int transition, uint64_t *out) {
    if (!G.shape_set || !seq || !out || j1 <= j0) return 0;
    const ShapeDesc sh = G.shape;
    const int span = sh.span, w = sh.weight;
    const unsigned char *s = reinterpret_cast<const unsigned char *>(seq) + block_start;
#if defined(__x86_64__)
    static const bool has_bmi2 = __builtin_cpu_supports("bmi2");
    if (has_bmi2 && span <= 32) return host_chunk_seeds_bmi2(s, j0, j1, sh, transition, out);
#endif
    size_t n = 0;
    // last_bad = index of the most recent non-ACGT character seen in [j0, j+span)
    int64_t last_bad = -1;
    for (uint32_t i = j0; i + 1 < j0 + (uint32_t)span && i < j1 + (uint32_t)span - 1; i++)
        if (g_nt_lut.v[s[i]] > 3) last_bad = i;
    for (uint32_t j = j0; j < j1; j++) {
        uint32_t tail = j + (uint32_t)span - 1;
        if (g_nt_lut.v[s[tail]] > 3) last_bad = tail;
        if (last_bad >= (int64_t)j) continue;
        uint64_t kmer = 0;
        for (int i = 0; i < w; i++) kmer = (kmer << 2) | g_nt_lut.v[s[j + sh.pos[i]]];
        out[n++] = (kmer << 32) + j;
        if (transition) {
            for (int t = 0; t < w; t++)
                if (sh.trans[t]) out[n++] = ((kmer ^ ((uint64_t)2 << (2 * t))) << 32) + j;
        }
    }
    return n;
}

#include "host_segments.inc"
#include "host_pipeline.inc"

} // extern "C"
