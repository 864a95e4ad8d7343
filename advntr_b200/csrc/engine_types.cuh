// engine_types.cuh: error plumbing, device descriptors, context / model handles -- part of libadvhmm.so (see advhmm.cu for the overview)
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <memory>
#include <mutex>
#include <new>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

#include "locus_compile.hpp"

using namespace advhmm;

// error plumbing
// =============================================================================================
namespace {

thread_local std::string g_last_error;

int set_error(int code, const char* fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}

#define CU_TRY(expr)                                                                          \
    do {                                                                                      \
        cudaError_t e__ = (expr);                                                             \
        if (e__ != cudaSuccess)                                                               \
            return set_error(e__ == cudaErrorMemoryAllocation ? ADVHMM_ENOMEM : ADVHMM_ECUDA, \
                             "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__),         \
                             __FILE__, __LINE__);                                             \
    } while (0)

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 4 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) { cudaGetLastError(); e = cudaMalloc(&p, want = bytes); }
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct PinnedBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 4 + 256;
        cudaError_t e = cudaHostAlloc(&p, want, cudaHostAllocDefault);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

// =============================================================================================
// device-side model descriptors
// =============================================================================================
struct DevGeneric {
    int m, S, K, start, end, finite, n_levels, max_in_degree;
    const int32_t* in_off;
    const int32_t* in_src;
    const double* in_w;
    const double* emis;
    const double* v0;
    const int32_t* tb0;
    const int32_t* lvl_off;
    const int32_t* lvl_state;
    const uint8_t* classes;       // [m] state class bytes for the on-device path reducers
};

struct DevBanded {
    int NC, P, S, m, NF, end_final, acc_col, n_acc;
    int start, end, image_bytes, pad0;
    double logp_empty;            // v0[end]: the answer for an empty read
    const unsigned char* image;   // smem image, see kImg* (208 bytes per column)
    const int32_t* st;            // [3*NC] slot -> state
    const int32_t* tb1;           // [4*S]
    const int32_t* acc_src_col;   // [n_acc]
    const int32_t* fin_state;     // [NF]
    const int32_t* fin_off;       // [NF+1]
    const int32_t* fin_src;
    const double* fin_w;
    const int32_t* tb0;           // [m]
    // fp32 twin (ADVHMM_FP32): float image (112 bytes per column) and the host-evaluated tables
    // re-evaluated in float arithmetic
    const unsigned char* image_f;
    int image_f_bytes, pad1;
    float logp_empty_f, pad2;
    const int32_t* tb1_f;
    const int32_t* tb0_f;
    const float* fin_w_f;
    const uint8_t* classes;       // [m]
    // forward (log_probability): first-row table [sym][P]{I, M} evaluated with pair_lse, value of the empty read
    const double* f1;
    double logp_empty_fwd;
};

struct Tile {
    const void* model;   // DevBanded* or DevGeneric*
    int32_t first;       // first index into order[]
    int32_t cnt;         // reads in this tile (<= warps per block)
};

constexpr int kMaxRPL = 10;             // read positions per lane: reads up to 320 bases on the banded path
constexpr int kBandedWarpsMax = 16;     // reads per CTA (banded): run-time choice, see ctx->banded_warps
constexpr int kGenericWarpsMax = 8;

}  // namespace

// one device allocation holding the tables of every model of a batch (advhmm_models_create_for_loci);
// freed when the last model of the batch is destroyed
// Allocated and freed with the stream-ordered allocator: the free is queued on the context's compute
// stream behind the last kernel that read the tables, so destroying the models of one batch neither
// waits for the device nor stalls the batch that is being decoded (cudaFree would do both).
struct DeviceArena {
    int device = -1;
    void* p = nullptr;
    size_t bytes = 0;
    cudaStream_t free_stream = nullptr;                    // the owning context's compute stream
    std::shared_ptr<std::atomic<bool>> ctx_alive;          // false once that context (and stream) is gone
    ~DeviceArena()
    {
        if (!p) return;
        cudaSetDevice(device);
        if (ctx_alive && ctx_alive->load()) cudaFreeAsync(p, free_stream);
        else cudaFree(p);
    }
};

// structural device tables of a shape, shared by every locus model of that shape on a context
struct DevShape {
    std::shared_ptr<const rm::ShapeStructure> shape;   // kept alive: the map below is keyed by its address
    DevBuf blob;
    const int32_t* st = nullptr;
    const int32_t* acc_src_col = nullptr;
    const int32_t* fin_state = nullptr;
    const int32_t* fin_off = nullptr;
    const int32_t* fin_src = nullptr;
};

struct advhmm_context {
    int device = -1;             // < 0: host-only context (model analysis without a GPU)
    cudaStream_t stream = nullptr;
    bool owns_stream = false;
    int sm_count = 0;
    size_t smem_optin = 0;
    int64_t launches = 0;
    size_t workspace_budget = 0;       // traceback workspace a batch may take (short-read / generic families)
    size_t workspace_budget_long = 0;  // ... the long-read family: ~100 MB per 20 kb read, see advhmm_context_create
    DevBuf d_seqs, d_seq_off, d_pk, d_meta, d_work, d_out, d_paths, d_flags;
    DevBuf d_badflag;            // device-buffer calls: first read with a code outside the alphabet
    PinnedBuf h_meta, h_out;
    cudaEvent_t meta_done = nullptr;
    // host-buffer calls: state paths go home chunk by chunk on a second stream while the next chunk
    // is decoded (marks = cursor snapshots + events recorded after every backtrack launch)
    cudaStream_t copy_stream = nullptr;
    std::vector<cudaEvent_t> chunk_events;
    std::vector<cudaEvent_t> h2d_events;     // sequences of sub-batch k are on the device (run_host)
    PinnedBuf h_cursors;
    size_t n_marks = 0;
    bool mark_chunks = false;
    size_t host_chunks = 8;      // chunks a host-buffer call asks run_batch for (path copies overlap decoding)
    // optional per-kernel timing (bench.py roofline): event pairs around every fill launch
    bool profile = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_events[2];   // [0] banded fill, [1] backtrack
    size_t prof_used[2] = {0, 0};
    // locus models: per-shape structural tables on this device, pinned staging for batched uploads
    std::map<const rm::ShapeStructure*, std::unique_ptr<DevShape>> shape_dev;
    PinnedBuf h_stage[2];
    cudaEvent_t stage_done[2] = {nullptr, nullptr};
    cudaStream_t bt_stream = nullptr;        // long reads: backtrack of chunk k next to the fill of chunk k+1
    cudaEvent_t fill_done[2] = {nullptr, nullptr}, bt_done[2] = {nullptr, nullptr};
    cudaStream_t upload_stream = nullptr;    // model tables travel here, next to the decoding of the previous batch
    cudaEvent_t upload_done = nullptr;
    std::shared_ptr<std::atomic<bool>> alive = std::make_shared<std::atomic<bool>>(true);
    int banded_warps = 8;        // reads per CTA of the banded kernel
    int short_max_len = 32 * kMaxRPL;   // longer reads take the striped long-read kernel (ADVHMM_SHORT_MAX_LEN)
    std::mutex mu;
};

struct advhmm_model {
    advhmm_context* ctx = nullptr;
    int K = 4, m = 0, NCpad = 0; // alphabet size, states, padded columns (0: not banded)
    CompiledModel cm;            // graph analysis of THIS model (models made from a descriptor; locus
                                 // models share their shape's analysis and fill this only on demand)
    // locus models (advhmm_models_create_for_loci): the shape + the few KB of numbers of the locus
    rm::LocusValues locus;
    bool lean = false;           // device tables = the banded kernel's only (image, first-row tables, row 0)
    std::shared_ptr<DeviceArena> arena;
    DevBuf blob;                 // all device tables of this model
    DevGeneric* d_generic = nullptr;
    DevGeneric* d_generic_fwd = nullptr;   // same tables, row 0 closed with pair_lse (forward)
    DevBanded* d_banded = nullptr;
    uint8_t* d_classes = nullptr;  // [n_states] state class bytes (zero until set_state_classes)
    int banded_smem = 0;         // image bytes (0: not banded / does not fit)
    advhmm_model_info info{};
};

// =============================================================================================
// small device helpers
// =============================================================================================
