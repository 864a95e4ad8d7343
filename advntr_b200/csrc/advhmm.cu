// advhmm.cu -- B200 (sm_100a) Viterbi / forward engine for adVNTR's profile HMMs + its C-ABI.
//
// Replaces the vendored pomegranate's `_viterbi` / `_forward`
// (/root/reference/pomegranate/hmm.pyx:1970-2136, 1371-1484) for batches of reads.
// See DESIGN.md for the data layout and the roofline of each kernel.
//
// Kernels (all hand-written, no tensor cores -- this is max-plus DP, not a contraction):
//   pack_reads_kernel        byte codes -> 2-bit packed reads (+ reverse complement, validation)
//   banded_fill_kernel<RPL>  profile-shaped models: one warp per read, lanes own blocks of RPL
//                            read positions, columns sweep as a register wavefront (skew 1
//                            column per lane, 3 shuffles per step); model tables staged into
//                            shared memory with one TMA bulk copy per CTA; 6-bit traceback per
//                            (position, column) packed into one word per lane and step.
//   banded_backtrack_kernel  device backtrack to the state path (one thread per read)
//   generic_fill_kernel      any baked model: row-synchronous CSR kernel, silent states by level
//   generic_backtrack_kernel
//   generic_forward_kernel   log_probability (sum-product with the reference's pair_lse)
//
// Exactness: every DP value is produced by the same IEEE-754 double operations in the same
// order as the reference ((v + t) + e per edge, strict '>' in candidate order), so paths and
// log-probabilities are bit-identical; see model_compile.hpp for why the banded schedule may
// skip / reorder the candidates it does.
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

#include "model_compile.hpp"

using namespace advhmm;

// =============================================================================================
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

struct advhmm_context {
    int device = -1;             // < 0: host-only context (model analysis without a GPU)
    cudaStream_t stream = nullptr;
    bool owns_stream = false;
    int sm_count = 0;
    size_t smem_optin = 0;
    int64_t launches = 0;
    size_t workspace_budget = 0;
    DevBuf d_seqs, d_seq_off, d_pk, d_meta, d_work, d_out, d_paths, d_flags;
    PinnedBuf h_meta, h_out;
    cudaEvent_t meta_done = nullptr;
    // optional per-kernel timing (bench.py roofline): event pairs around every fill launch
    bool profile = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_events[2];   // [0] banded fill, [1] backtrack
    size_t prof_used[2] = {0, 0};
    int banded_smem_set[kMaxRPL + 4] = {0};   // dynamic-smem opt-in already applied per kernel variant
    int banded_warps = 8;        // reads per CTA of the banded kernel
    bool int_compare = false;    // integer-pipe compares (ADVHMM_ICMP=1); needs all tables <= 0
    bool launch_int_compare = false;   // ... and every model of the current batch qualifies
    int generic_smem_set = 0;
    int banded_f32_smem_set[kMaxRPL + 1] = {0};
    std::mutex mu;
};

struct advhmm_model {
    advhmm_context* ctx = nullptr;
    CompiledModel cm;
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
namespace {

__device__ __forceinline__ uint32_t smem_u32(const void* p)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// TMA bulk copy global -> shared, completion signalled on the mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!done);
}

__device__ __forceinline__ double shfl_up_f64(double v, int delta)
{
    int lo = __double2loint(v), hi = __double2hiint(v);
    lo = __shfl_up_sync(0xffffffffu, lo, delta);
    hi = __shfl_up_sync(0xffffffffu, hi, delta);
    return __hiloint2double(hi, lo);
}
__device__ __forceinline__ double shfl_xor_f64(double v, int mask)
{
    int lo = __double2loint(v), hi = __double2hiint(v);
    lo = __shfl_xor_sync(0xffffffffu, lo, mask);
    hi = __shfl_xor_sync(0xffffffffu, hi, mask);
    return __hiloint2double(hi, lo);
}

// symbol i of a 2-bit packed read
__device__ __forceinline__ int packed_sym(const uint32_t* __restrict__ pk, int i)
{
    return (pk[i >> 4] >> ((i & 15) * 2)) & 3;
}

// lexicographic (max value, min index) warp all-reduce; result in every lane
__device__ __forceinline__ void warp_argmax_first(double& v, int& idx)
{
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        double ov = shfl_xor_f64(v, off);
        int oi = __shfl_xor_sync(0xffffffffu, idx, off);
        if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
    }
}

// =============================================================================================
// pack kernel: one CTA (32 threads) per result read
// =============================================================================================
struct PackArgs {
    const uint8_t* seqs;
    const int64_t* seq_off;     // [n_reads+1]
    const int64_t* pk_off;      // [n_out] word offsets
    uint32_t* pk;
    int32_t* rlen;              // [n_out]
    int32_t* bad;               // [1]: first read index with a code >= n_symbols (atomicMin)
    int n_out, strands, n_symbols;
};

__global__ void __launch_bounds__(32) pack_reads_kernel(PackArgs a)
{
    const int q = blockIdx.x;
    if (q >= a.n_out) return;
    const int r = q / a.strands;
    const bool rc = (a.strands == 2) && (q & 1);
    const int64_t s0 = a.seq_off[r];
    const int n = (int)(a.seq_off[r + 1] - s0);
    const uint8_t* __restrict__ s = a.seqs + s0;
    uint32_t* __restrict__ out = a.pk + a.pk_off[q];
    const int words = (n + 15) / 16 + 1;
    if (threadIdx.x == 0) a.rlen[q] = n;
    bool bad = false;
    for (int w = threadIdx.x; w < words; w += 32) {
        uint32_t word = 0;
        const int base = w * 16;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const int p = base + k;
            if (p < n) {
                int code = rc ? s[n - 1 - p] : s[p];
                if (code >= a.n_symbols) { bad = true; code = 0; }
                if (rc) code = 3 - code;          // A<->T, C<->G under codes A,C,G,T = 0..3
                word |= (uint32_t)(code & 3) << (2 * k);
            }
        }
        out[w] = word;
    }
    if (bad) atomicMin(a.bad, r);
}

// =============================================================================================
// banded fill kernel
// =============================================================================================
struct BandedArgs {
    const Tile* tiles;
    const int32_t* order;       // result-read id of every work item
    int32_t chunk_base;         // first work item of this chunk (workspace slot = item - chunk_base)
    const uint32_t* pk;
    const int64_t* pk_off;
    const int32_t* rlen;
    double* logp;               // [n_out]
    uint32_t* tbw;              // traceback words, per slot 32 * Pmax * (RPL > 5 ? 2 : 1) words
    size_t tbw_stride;          // 32-bit words per slot
    uint16_t* acc_tb;           // collector choice (source column) per (slot, position)
    int acc_stride;             // entries per slot (32 * RPL)
    double* vfin;               // last-row values, per slot 3 * Pmax
    size_t vfin_stride;
    int32_t* ftb;               // final-state choices, per slot 32
};

// shared-memory image of a banded model (byte offsets from the start of dynamic smem), P = NCpad:
//   [0, 80P)          w10[c][10] : wII wIM wID | wMI wMM wMD | wDI wDM wDD | accw      (doubles)
//   [80P, 144P)       e2[sym][c][2] : emission log-prob of the I and M slot           (doubles)
//   [144P, 208P)      v1[sym][c][2] : first-row values of the I and M slot            (doubles)
// One 16-byte row per (sym, c) means a lane fetches both emissions with one LDS.128, and all ten
// weights of a column with five LDS.128 off a single address register (80-byte stride: the
// quarter-warp's eight 16-byte accesses fall into disjoint bank groups).
constexpr int kImgW = 0, kImgE = 80, kImgV1 = 144, kImgBytesPerCol = 208;

__device__ __forceinline__ double2 lds128(uint32_t addr)
{
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
    return v;
}

// First strict maximum of three candidates in candidate order (hmm.pyx:2039 `if cand > best`).
// Two traceback bits: bit0 = (a1 > a0), bit1 = (a2 > max(a0, a1));  source = bit1 ? 2 : bit0.
template <int J, int N, typename F>
__device__ __forceinline__ void static_for(F&& f)
{
    if constexpr (J < N) {
        f(std::integral_constant<int, J>{});
        static_for<J + 1, N>(f);
    }
}

template <int SH, bool ICMP>
__device__ __forceinline__ double max3_first(double a0, double a1, double a2, uint32_t& bits)
{
    double m;
    if (ICMP) {
        // All DP values are <= 0 (sums of log-probabilities; checked when the model is compiled),
        // so a > b  <=>  bits(a) < bits(b) as unsigned 64-bit integers (+0.0 is the largest value
        // and the smallest pattern, -inf the smallest value and the largest non-NaN pattern).
        // The compares then run on the integer pipe and leave the fp64 pipe to the adds.
        asm("{\n\t"
            ".reg .pred p1, p2;\n\t"
            ".reg .b64 t, x0, x1, x2;\n\t"
            "mov.b64 x0, %2;\n\t"
            "mov.b64 x1, %3;\n\t"
            "mov.b64 x2, %4;\n\t"
            "setp.lt.u64 p1, x1, x0;\n\t"
            "selp.b64 t, x1, x0, p1;\n\t"
            "@p1 or.b32 %1, %1, %5;\n\t"
            "setp.lt.u64 p2, x2, t;\n\t"
            "selp.b64 t, x2, t, p2;\n\t"
            "@p2 or.b32 %1, %1, %6;\n\t"
            "mov.b64 %0, t;\n\t"
            "}"
            : "=d"(m), "+r"(bits)
            : "d"(a0), "d"(a1), "d"(a2), "n"(1u << SH), "n"(2u << SH));
    } else {
        // written in PTX so that each compare costs one DSETP, one 64-bit select and one predicated OR
        asm("{\n\t"
            ".reg .pred p1, p2;\n\t"
            ".reg .f64 t;\n\t"
            "setp.gt.f64 p1, %3, %2;\n\t"
            "selp.f64 t, %3, %2, p1;\n\t"
            "@p1 or.b32 %1, %1, %5;\n\t"
            "setp.gt.f64 p2, %4, t;\n\t"
            "selp.f64 %0, %4, t, p2;\n\t"
            "@p2 or.b32 %1, %1, %6;\n\t"
            "}"
            : "=d"(m), "+r"(bits)
            : "d"(a0), "d"(a1), "d"(a2), "n"(1u << SH), "n"(2u << SH));
    }
    return m;
}

// ALIGNED: the read length is a multiple of RPL, so the last read position is the last row of a
// lane and its values can be stored from fixed registers.
template <int RPL, bool ALIGNED, bool ICMP>
__device__ __forceinline__ void banded_sweep(const int NC, const int P, const int nl, const int ln, const int jn,
                                             const int lane, const uint32_t s_base, const uint32_t symbits,
                                             const int acc_col, uint32_t* __restrict__ tbw,
                                             uint16_t* __restrict__ acc_tb, double* __restrict__ vfin)
{
    constexpr int NW = RPL > 5 ? 2 : 1;
    uint32_t eaddr[RPL];                           // smem address of e2[sym_j][0]
#pragma unroll
    for (int j = 0; j < RPL; ++j)
        eaddr[j] = s_base + (uint32_t)kImgE * P + ((symbits >> (2 * j)) & 3u) * (uint32_t)(16 * P);
    const uint32_t v1_delta = (uint32_t)(kImgV1 - kImgE) * P;

    double cI[RPL], cM[RPL], cD[RPL], acc[RPL];
    int accarg[RPL];
#pragma unroll
    for (int j = 0; j < RPL; ++j) { cI[j] = cM[j] = cD[j] = acc[j] = kNegInf; accarg[j] = 0; }
    double bI = kNegInf, bM = kNegInf, bD = kNegInf;   // row above my block, previous column
    const bool store_last = ALIGNED && lane == ln;
    // lane-skewed views: index them with the step t (column c = t - lane).  The empty asm keeps
    // ptxas from re-deriving these addresses inside the loop.
    uint32_t* tbw_t = tbw - (ptrdiff_t)lane * NW;
    double* vfin_t = vfin - lane;
    uint32_t w_t = s_base - (uint32_t)lane * 80u;       // + 80 t  -> w10[c]
    uint32_t e_t[RPL];                                  // + 16 t  -> e2[sym_j][c]
#pragma unroll
    for (int j = 0; j < RPL; ++j) { e_t[j] = eaddr[j] - (uint32_t)lane * 16u; asm volatile("" : "+r"(e_t[j])); }
    asm volatile("" : "+l"(tbw_t), "+l"(vfin_t), "+r"(w_t));

    const int steps = NC + nl - 1;
#pragma unroll 1
    for (int t = 0; t < steps; ++t) {
        // the row above my block at column c was finished by lane-1 in the previous step
        const double uI0 = shfl_up_f64(cI[RPL - 1], 1);
        const double uM0 = shfl_up_f64(cM[RPL - 1], 1);
        const double uD0 = shfl_up_f64(cD[RPL - 1], 1);
        const int c = t - lane;
        if (c < 0 || c >= NC || lane >= nl) continue;

        const uint32_t wa = w_t + (uint32_t)t * 80u;
        const double2 w01 = lds128(wa), w23 = lds128(wa + 16), w45 = lds128(wa + 32);
        const double2 w67 = lds128(wa + 48), w89 = lds128(wa + 64);
        const double wII = w01.x, wIM = w01.y, wID = w23.x, wMI = w23.y, wMM = w45.x, wMD = w45.y;
        const double wDI = w67.x, wDM = w67.y, wDD = w89.x, aw = w89.y;
        const uint32_t cb = (uint32_t)t * 16u;

        // M and D slots depend only on values of the previous column / previous step
        double nM[RPL], nD[RPL], eIr[RPL];
        uint32_t word[NW];
#pragma unroll
        for (int k = 0; k < NW; ++k) word[k] = 0;
        static_for<0, RPL>([&](auto jc) {
            constexpr int j = decltype(jc)::value;
            const double2 e = lds128(e_t[j] + cb);             // {eI, eM}
            eIr[j] = e.x;
            const double oI = j ? cI[j ? j - 1 : 0] : bI, oM = j ? cM[j ? j - 1 : 0] : bM, oD = j ? cD[j ? j - 1 : 0] : bD;
            nM[j] = max3_first<6 * (j % 5) + 2, ICMP>((oI + wMI) + e.y, (oM + wMM) + e.y, (oD + wMD) + e.y, word[j / 5]);
            nD[j] = max3_first<6 * (j % 5) + 4, ICMP>(cI[j] + wDI, cM[j] + wDM, cD[j] + wDD, word[j / 5]);
        });
        if (lane == 0) {                                       // first read position: from row 0
            const double2 f = lds128(e_t[0] + v1_delta + cb);
            nM[0] = f.y;
            eIr[0] = f.x;                                      // (carries vI of row 1, see below)
        }
        // collector (end_repeating_pattern_match): D of its column is the best unit_end so far.
        // Both cases are rare per lane (1 and `copies` columns of NC), hence real branches.
        if (c == acc_col) {
#pragma unroll
            for (int j = 0; j < RPL; ++j) nD[j] = acc[j];
        }
        if (aw > kNegInf) {                                    // a unit_end column
#pragma unroll
            for (int j = 0; j < RPL; ++j) {
                const double cand = nD[j] + aw;
                if (cand > acc[j]) { acc[j] = cand; accarg[j] = c; }
            }
        }
        // I slots chain down the rows of this column
        double uI = uI0, uM = uM0, uD = uD0;
        static_for<0, RPL>([&](auto jc) {
            constexpr int j = decltype(jc)::value;
            double vI = max3_first<6 * (j % 5), ICMP>((uI + wII) + eIr[j], (uM + wIM) + eIr[j], (uD + wID) + eIr[j], word[j / 5]);
            if (j == 0 && lane == 0) vI = eIr[0];
            uI = vI; uM = nM[j]; uD = nD[j];
            cI[j] = vI; cM[j] = nM[j]; cD[j] = nD[j];
        });
        bI = uI0; bM = uM0; bD = uD0;
        if (NW == 1) tbw_t[t] = word[0];
        else reinterpret_cast<uint2*>(tbw_t)[t] = make_uint2(word[0], word[NW - 1]);
        if (ALIGNED) {
            if (store_last) { vfin_t[t] = cI[RPL - 1]; vfin_t[P + t] = cM[RPL - 1]; vfin_t[2 * P + t] = cD[RPL - 1]; }
        } else if (lane == ln) {
            double fI = cI[0], fM = cM[0], fD = cD[0];
#pragma unroll
            for (int j = 1; j < RPL; ++j)
                if (j == jn) { fI = cI[j]; fM = cM[j]; fD = cD[j]; }
            vfin_t[t] = fI; vfin_t[P + t] = fM; vfin_t[2 * P + t] = fD;
        }
    }
    // which unit_end fed the collector, per read position (read by the backtrack only)
    if (lane < nl) {
#pragma unroll
        for (int j = 0; j < RPL; ++j) acc_tb[j] = (uint16_t)accarg[j];
    }
}

template <int RPL, int WPB, bool ICMP>
__global__ void __launch_bounds__(WPB * 32, 2)
banded_fill_kernel(const BandedArgs a)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ double s_fval[WPB][32];

    const Tile tile = a.tiles[blockIdx.x];
    const DevBanded* __restrict__ M = reinterpret_cast<const DevBanded*>(tile.model);
    const int P = M->P, NC = M->NC;

    // ---- stage the model tables: one elected thread issues TMA bulk copies -----------------
    if (threadIdx.x == 0) mbar_init(&s_bar, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t bytes = (uint32_t)M->image_bytes;
        mbar_expect_tx(&s_bar, bytes);
#pragma unroll 1
        for (uint32_t o = 0; o < bytes; o += 65536u)
            tma_bulk_g2s(smem_raw + o, M->image + o, min(65536u, bytes - o), &s_bar);
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    mbar_wait(&s_bar, 0);
    if (warp >= tile.cnt) return;

    const int item = tile.first + warp;
    const int q = a.order[item];
    const size_t slot = (size_t)(item - a.chunk_base);
    const int n = a.rlen[q];
    if (n == 0) {
        if (lane == 0) a.logp[q] = M->logp_empty;
        return;
    }
    const int nl = (n + RPL - 1) / RPL;                 // lanes in use
    const int ln = (n - 1) / RPL, jn = (n - 1) % RPL;   // owner of the last read position

    // ---- my RPL symbols (2 bits each) ------------------------------------------------------
    const uint32_t* __restrict__ pk = a.pk + a.pk_off[q];
    uint32_t symbits;
    {
        const int bit = 2 * lane * RPL;
        const int w = bit >> 5, sh = bit & 31;
        const int last_word = (n + 15) / 16;            // allocation has (n+15)/16 + 1 words
        const uint32_t lo = (w <= last_word) ? pk[w] : 0u;
        const uint32_t hi = (w + 1 <= last_word) ? pk[w + 1] : 0u;
        symbits = __funnelshift_r(lo, hi, sh);
    }
    constexpr int NW = RPL > 5 ? 2 : 1;
    uint32_t* __restrict__ tbw = a.tbw + slot * a.tbw_stride + (size_t)lane * P * NW;
    uint16_t* __restrict__ acc_tb = a.acc_tb + slot * (size_t)a.acc_stride + lane * RPL;
    double* __restrict__ vfin = a.vfin + slot * a.vfin_stride;
    const uint32_t s_base = smem_u32(smem_raw);
    if (jn == RPL - 1)
        banded_sweep<RPL, true, ICMP>(NC, P, nl, ln, jn, lane, s_base, symbits, M->acc_col, tbw, acc_tb, vfin);
    else
        banded_sweep<RPL, false, ICMP>(NC, P, nl, ln, jn, lane, s_base, symbits, M->acc_col, tbw, acc_tb, vfin);
    __syncwarp();

    // ---- final-only silent states on the last row (hub reductions across the warp) ---------
    const int NF = M->NF;
    int32_t* __restrict__ ftb = a.ftb + slot * 32;
    for (int f = 0; f < NF; ++f) {
        const int k0 = M->fin_off[f], k1 = M->fin_off[f + 1];
        double best = kNegInf;
        int arg = 0x7fffffff;
        for (int k = k0 + lane; k < k1; k += 32) {
            const int code = M->fin_src[k];
            const double sv = code < 0 ? s_fval[warp][-(code + 1)] : vfin[code];
            const double cand = sv + M->fin_w[k];
            if (cand > best) { best = cand; arg = k; }
        }
        warp_argmax_first(best, arg);
        if (lane == 0) {
            s_fval[warp][f] = best;
            ftb[f] = (best > kNegInf) ? M->fin_src[arg] : 0;
        }
        __syncwarp();
    }
    if (lane == 0) a.logp[q] = s_fval[warp][M->end_final];
}

// =============================================================================================
// fp32 variant of the banded fill kernel (optional mode, ADVHMM_FP32): same schedule and
// operation order with float tables / float arithmetic.  Image: 112 bytes per column
//   [0, 48P)    w12[c][12] floats: the ten weights of the fp64 image + 2 pad
//   [48P, 80P)  e2[sym][c][2] floats        [80P, 112P)  v1[sym][c][2] floats
// =============================================================================================
constexpr int kImgFE = 48, kImgFV1 = 80, kImgFBytesPerCol = 112;

__device__ __forceinline__ float4 lds128f(uint32_t addr)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ float2 lds64f(uint32_t addr)
{
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
    return v;
}

template <int SH>
__device__ __forceinline__ float max3_first_f32(float a0, float a1, float a2, uint32_t& bits)
{
    float m;
    asm("{\n\t"
        ".reg .pred p1, p2;\n\t"
        ".reg .f32 t;\n\t"
        "setp.gt.f32 p1, %3, %2;\n\t"
        "selp.f32 t, %3, %2, p1;\n\t"
        "@p1 or.b32 %1, %1, %5;\n\t"
        "setp.gt.f32 p2, %4, t;\n\t"
        "selp.f32 %0, %4, t, p2;\n\t"
        "@p2 or.b32 %1, %1, %6;\n\t"
        "}"
        : "=f"(m), "+r"(bits)
        : "f"(a0), "f"(a1), "f"(a2), "n"(1u << SH), "n"(2u << SH));
    return m;
}

template <int RPL>
__global__ void __launch_bounds__(8 * 32, 2)
banded_fill_f32_kernel(const BandedArgs a)
{
    constexpr int WPB = 8, NW = RPL > 5 ? 2 : 1;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ float s_fval[WPB][32];
    const float kNI = -__int_as_float(0x7f800000);

    const Tile tile = a.tiles[blockIdx.x];
    const DevBanded* __restrict__ M = reinterpret_cast<const DevBanded*>(tile.model);
    const int P = M->P, NC = M->NC, acc_col = M->acc_col;
    if (threadIdx.x == 0) mbar_init(&s_bar, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t bytes = (uint32_t)M->image_f_bytes;
        mbar_expect_tx(&s_bar, bytes);
#pragma unroll 1
        for (uint32_t o = 0; o < bytes; o += 65536u)
            tma_bulk_g2s(smem_raw + o, M->image_f + o, min(65536u, bytes - o), &s_bar);
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    mbar_wait(&s_bar, 0);
    if (warp >= tile.cnt) return;
    const int item = tile.first + warp;
    const int q = a.order[item];
    const size_t slot = (size_t)(item - a.chunk_base);
    const int n = a.rlen[q];
    if (n == 0) {
        if (lane == 0) a.logp[q] = (double)M->logp_empty_f;
        return;
    }
    const int nl = (n + RPL - 1) / RPL;
    const int ln = (n - 1) / RPL, jn = (n - 1) % RPL;
    const uint32_t* __restrict__ pk = a.pk + a.pk_off[q];
    uint32_t symbits;
    {
        const int bit = 2 * lane * RPL;
        const int w = bit >> 5, sh = bit & 31;
        const int last_word = (n + 15) / 16;
        const uint32_t lo = (w <= last_word) ? pk[w] : 0u;
        const uint32_t hi = (w + 1 <= last_word) ? pk[w + 1] : 0u;
        symbits = __funnelshift_r(lo, hi, sh);
    }
    uint32_t* tbw_t = a.tbw + slot * a.tbw_stride + (size_t)lane * P * NW - (ptrdiff_t)lane * NW;
    uint16_t* __restrict__ acc_tb = a.acc_tb + slot * (size_t)a.acc_stride + lane * RPL;
    float* __restrict__ vfin = reinterpret_cast<float*>(a.vfin + slot * a.vfin_stride);
    float* vfin_t = vfin - lane;
    const uint32_t s_base = smem_u32(smem_raw);
    uint32_t w_t = s_base - (uint32_t)lane * 48u;
    uint32_t e_t[RPL];
#pragma unroll
    for (int j = 0; j < RPL; ++j) {
        e_t[j] = s_base + (uint32_t)kImgFE * P + ((symbits >> (2 * j)) & 3u) * (uint32_t)(8 * P) - (uint32_t)lane * 8u;
        asm volatile("" : "+r"(e_t[j]));
    }
    asm volatile("" : "+l"(tbw_t), "+l"(vfin_t), "+r"(w_t));
    const uint32_t v1_delta = (uint32_t)(kImgFV1 - kImgFE) * P;

    float cI[RPL], cM[RPL], cD[RPL], acc[RPL];
    int accarg[RPL];
#pragma unroll
    for (int j = 0; j < RPL; ++j) { cI[j] = cM[j] = cD[j] = acc[j] = kNI; accarg[j] = 0; }
    float bI = kNI, bM = kNI, bD = kNI;

    const int steps = NC + nl - 1;
#pragma unroll 1
    for (int t = 0; t < steps; ++t) {
        const float uI0 = __shfl_up_sync(0xffffffffu, cI[RPL - 1], 1);
        const float uM0 = __shfl_up_sync(0xffffffffu, cM[RPL - 1], 1);
        const float uD0 = __shfl_up_sync(0xffffffffu, cD[RPL - 1], 1);
        const int c = t - lane;
        if (c < 0 || c >= NC || lane >= nl) continue;
        const uint32_t wa = w_t + (uint32_t)t * 48u;
        const float4 w0 = lds128f(wa), w1 = lds128f(wa + 16), w2 = lds128f(wa + 32);
        const float wII = w0.x, wIM = w0.y, wID = w0.z, wMI = w0.w, wMM = w1.x, wMD = w1.y;
        const float wDI = w1.z, wDM = w1.w, wDD = w2.x, aw = w2.y;
        const uint32_t cb = (uint32_t)t * 8u;
        float nM[RPL], nD[RPL], eIr[RPL];
        uint32_t word[NW];
#pragma unroll
        for (int k = 0; k < NW; ++k) word[k] = 0;
        static_for<0, RPL>([&](auto jc) {
            constexpr int j = decltype(jc)::value;
            const float2 e = lds64f(e_t[j] + cb);
            eIr[j] = e.x;
            const float oI = j ? cI[j ? j - 1 : 0] : bI, oM = j ? cM[j ? j - 1 : 0] : bM, oD = j ? cD[j ? j - 1 : 0] : bD;
            nM[j] = max3_first_f32<6 * (j % 5) + 2>((oI + wMI) + e.y, (oM + wMM) + e.y, (oD + wMD) + e.y, word[j / 5]);
            nD[j] = max3_first_f32<6 * (j % 5) + 4>(cI[j] + wDI, cM[j] + wDM, cD[j] + wDD, word[j / 5]);
        });
        if (lane == 0) {
            const float2 f = lds64f(e_t[0] + v1_delta + cb);
            nM[0] = f.y;
            eIr[0] = f.x;
        }
        if (c == acc_col) {
#pragma unroll
            for (int j = 0; j < RPL; ++j) nD[j] = acc[j];
        }
        if (aw > kNI) {
#pragma unroll
            for (int j = 0; j < RPL; ++j) {
                const float cand = nD[j] + aw;
                if (cand > acc[j]) { acc[j] = cand; accarg[j] = c; }
            }
        }
        float uI = uI0, uM = uM0, uD = uD0;
        static_for<0, RPL>([&](auto jc) {
            constexpr int j = decltype(jc)::value;
            float vI = max3_first_f32<6 * (j % 5)>((uI + wII) + eIr[j], (uM + wIM) + eIr[j], (uD + wID) + eIr[j], word[j / 5]);
            if (j == 0 && lane == 0) vI = eIr[0];
            uI = vI; uM = nM[j]; uD = nD[j];
            cI[j] = vI; cM[j] = nM[j]; cD[j] = nD[j];
        });
        bI = uI0; bM = uM0; bD = uD0;
        if (NW == 1) tbw_t[t] = word[0];
        else reinterpret_cast<uint2*>(tbw_t)[t] = make_uint2(word[0], word[NW - 1]);
        if (lane == ln) {
            float fI = cI[0], fM = cM[0], fD = cD[0];
#pragma unroll
            for (int j = 1; j < RPL; ++j)
                if (j == jn) { fI = cI[j]; fM = cM[j]; fD = cD[j]; }
            vfin_t[t] = fI; vfin_t[P + t] = fM; vfin_t[2 * P + t] = fD;
        }
    }
    if (lane < nl) {
#pragma unroll
        for (int j = 0; j < RPL; ++j) acc_tb[j] = (uint16_t)accarg[j];
    }
    __syncwarp();
    const int NF = M->NF;
    int32_t* __restrict__ ftb = a.ftb + slot * 32;
    for (int f = 0; f < NF; ++f) {
        const int k0 = M->fin_off[f], k1 = M->fin_off[f + 1];
        float best = kNI;
        int arg = 0x7fffffff;
        for (int k = k0 + lane; k < k1; k += 32) {
            const int code = M->fin_src[k];
            const float sv = code < 0 ? s_fval[warp][-(code + 1)] : vfin[code];
            const float cand = sv + M->fin_w_f[k];
            if (cand > best) { best = cand; arg = k; }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, best, off);
            const int oi = __shfl_xor_sync(0xffffffffu, arg, off);
            if (ov > best || (ov == best && oi < arg)) { best = ov; arg = oi; }
        }
        if (lane == 0) {
            s_fval[warp][f] = best;
            ftb[f] = (best > kNI) ? M->fin_src[arg] : 0;
        }
        __syncwarp();
    }
    if (lane == 0) a.logp[q] = (double)s_fval[warp][M->end_final];
}

// =============================================================================================
// banded fill kernel for long reads / large models (PacBio-like, BASELINE config 3)
//
// Same recurrence, tables and traceback encoding as banded_fill_kernel, but
//   * a read is cut into stripes of 32*RPL positions that one warp sweeps one after the other;
//     the last position of a stripe is carried to the next stripe through a per-read buffer in
//     global memory (3 doubles per column, updated in place: writes trail reads by >= 32 columns)
//     that lane 0 consumes through a double-buffered 32-column ring in shared memory;
//   * the model tables are read from global memory through the read-only path (the image of a
//     100-copy model is > 1 MB); the warps of an SM walk the columns together, so L1 serves them.
// =============================================================================================
struct LongArgs {
    const Tile* tiles;
    const int32_t* order;
    int32_t chunk_base;
    const uint32_t* pk;
    const int64_t* pk_off;
    const int32_t* rlen;
    double* logp;
    uint32_t* tbw;              // per slot: stripes_max * 32 * Pmax * 2 words
    size_t tbw_stride;
    uint16_t* acc_tb;           // per slot: stripes_max * 32 * RPL entries
    size_t acc_stride;
    double* vfin;               // per slot 3 * Pmax
    size_t vfin_stride;
    double* carry;              // per slot 3 * Pmax
    size_t carry_stride;
    int32_t* ftb;               // per slot 32
};

constexpr int kLongRPL = 8;
constexpr int kLongWarps = 4;

__device__ __forceinline__ double2 ldg128(const unsigned char* p)
{
    double2 v;
    asm volatile("ld.global.nc.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}

__global__ void __launch_bounds__(kLongWarps * 32, 2)
banded_long_kernel(const LongArgs a)
{
    constexpr int RPL = kLongRPL, NW = 2, H = 32 * RPL;
    __shared__ double s_ring[kLongWarps][2][3][32];
    __shared__ double s_fval[kLongWarps][32];

    const Tile tile = a.tiles[blockIdx.x];
    const DevBanded* __restrict__ M = reinterpret_cast<const DevBanded*>(tile.model);
    const int P = M->P, NC = M->NC, acc_col = M->acc_col;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp >= tile.cnt) return;
    const int item = tile.first + warp;
    const int q = a.order[item];
    const size_t slot = (size_t)(item - a.chunk_base);
    const int n = a.rlen[q];
    if (n == 0) {
        if (lane == 0) a.logp[q] = M->logp_empty;
        return;
    }
    const unsigned char* __restrict__ img = M->image;
    const unsigned char* __restrict__ img_e = img + (size_t)kImgE * P;
    const unsigned char* __restrict__ img_v1 = img + (size_t)kImgV1 * P;
    const uint32_t* __restrict__ pk = a.pk + a.pk_off[q];
    double* __restrict__ vfin = a.vfin + slot * a.vfin_stride;
    double* __restrict__ carry = a.carry + slot * a.carry_stride;
    const int n_stripes = (n + H - 1) / H;
    const int last_word = (n + 15) / 16;

    for (int s = 0; s < n_stripes; ++s) {
        const int rows = min(H, n - s * H);
        const int nl = (rows + RPL - 1) / RPL;
        const bool last_stripe = (s == n_stripes - 1);
        const int ln = (rows - 1) / RPL, jn = (rows - 1) % RPL;
        uint32_t symbits;
        {
            const int bit = 2 * (s * H + lane * RPL);
            const int w = bit >> 5;
            symbits = (w <= last_word) ? (pk[w] >> (bit & 31)) & 0xffffu : 0u;
        }
        size_t eoff[RPL];
#pragma unroll
        for (int j = 0; j < RPL; ++j) eoff[j] = (size_t)((symbits >> (2 * j)) & 3u) * 16u * P;

        uint32_t* __restrict__ tbw = a.tbw + slot * a.tbw_stride + ((size_t)(s * 32 + lane) * P) * NW;
        uint16_t* __restrict__ acc_tb = a.acc_tb + slot * a.acc_stride + (size_t)s * H + lane * RPL;

        double cI[RPL], cM[RPL], cD[RPL], acc[RPL];
        int accarg[RPL];
#pragma unroll
        for (int j = 0; j < RPL; ++j) { cI[j] = cM[j] = cD[j] = acc[j] = kNegInf; accarg[j] = 0; }
        double bI = kNegInf, bM = kNegInf, bD = kNegInf;

        if (s > 0) {                           // ring block 0 = carried values of columns 0..31
#pragma unroll
            for (int k = 0; k < 3; ++k) s_ring[warp][0][k][lane] = carry[(size_t)k * P + lane];
        }
        __syncwarp();

        const int steps = NC + nl - 1;
#pragma unroll 1
        for (int t = 0; t < steps; ++t) {
            if (s > 0 && (t & 31) == 0) {      // prefetch the next 32 carried columns
                const int col = t + 32 + lane;
                const int buf = ((t >> 5) + 1) & 1;
#pragma unroll
                for (int k = 0; k < 3; ++k) s_ring[warp][buf][k][lane] = (col < P) ? carry[(size_t)k * P + col] : kNegInf;
                __syncwarp();
            }
            double uI0 = shfl_up_f64(cI[RPL - 1], 1);
            double uM0 = shfl_up_f64(cM[RPL - 1], 1);
            double uD0 = shfl_up_f64(cD[RPL - 1], 1);
            const int c = t - lane;
            if (c < 0 || c >= NC || lane >= nl) continue;
            if (lane == 0 && s > 0) {
                const int buf = (t >> 5) & 1;
                uI0 = s_ring[warp][buf][0][t & 31];
                uM0 = s_ring[warp][buf][1][t & 31];
                uD0 = s_ring[warp][buf][2][t & 31];
            }
            const unsigned char* wp = img + (size_t)c * 80u;
            const double2 w01 = ldg128(wp), w23 = ldg128(wp + 16), w45 = ldg128(wp + 32);
            const double2 w67 = ldg128(wp + 48), w89 = ldg128(wp + 64);
            const double wII = w01.x, wIM = w01.y, wID = w23.x, wMI = w23.y, wMM = w45.x, wMD = w45.y;
            const double wDI = w67.x, wDM = w67.y, wDD = w89.x, aw = w89.y;
            const size_t cb = (size_t)c * 16u;

            double nM[RPL], nD[RPL], eIr[RPL];
            uint32_t word[NW] = {0u, 0u};
            static_for<0, RPL>([&](auto jc) {
                constexpr int j = decltype(jc)::value;
                const double2 e = ldg128(img_e + eoff[j] + cb);
                eIr[j] = e.x;
                const double oI = j ? cI[j ? j - 1 : 0] : bI, oM = j ? cM[j ? j - 1 : 0] : bM, oD = j ? cD[j ? j - 1 : 0] : bD;
                nM[j] = max3_first<6 * (j % 5) + 2, false>((oI + wMI) + e.y, (oM + wMM) + e.y, (oD + wMD) + e.y, word[j / 5]);
                nD[j] = max3_first<6 * (j % 5) + 4, false>(cI[j] + wDI, cM[j] + wDM, cD[j] + wDD, word[j / 5]);
            });
            const bool first_row = (lane == 0 && s == 0);
            if (first_row) {
                const double2 f = ldg128(img_v1 + eoff[0] + cb);
                nM[0] = f.y;
                eIr[0] = f.x;
            }
            if (c == acc_col) {
#pragma unroll
                for (int j = 0; j < RPL; ++j) nD[j] = acc[j];
            }
            if (aw > kNegInf) {
#pragma unroll
                for (int j = 0; j < RPL; ++j) {
                    const double cand = nD[j] + aw;
                    if (cand > acc[j]) { acc[j] = cand; accarg[j] = c; }
                }
            }
            double uI = uI0, uM = uM0, uD = uD0;
            static_for<0, RPL>([&](auto jc) {
                constexpr int j = decltype(jc)::value;
                double vI = max3_first<6 * (j % 5), false>((uI + wII) + eIr[j], (uM + wIM) + eIr[j], (uD + wID) + eIr[j], word[j / 5]);
                if (j == 0 && first_row) vI = eIr[0];
                uI = vI; uM = nM[j]; uD = nD[j];
                cI[j] = vI; cM[j] = nM[j]; cD[j] = nD[j];
            });
            bI = uI0; bM = uM0; bD = uD0;
            reinterpret_cast<uint2*>(tbw)[c] = make_uint2(word[0], word[1]);
            if (!last_stripe) {
                if (lane == 31) {              // full stripe: lane 31 owns its last position
                    carry[c] = cI[RPL - 1]; carry[(size_t)P + c] = cM[RPL - 1]; carry[(size_t)2 * P + c] = cD[RPL - 1];
                }
            } else if (lane == ln) {
                double fI = cI[0], fM = cM[0], fD = cD[0];
#pragma unroll
                for (int j = 1; j < RPL; ++j)
                    if (j == jn) { fI = cI[j]; fM = cM[j]; fD = cD[j]; }
                vfin[c] = fI; vfin[(size_t)P + c] = fM; vfin[(size_t)2 * P + c] = fD;
            }
        }
        if (lane < nl) {
#pragma unroll
            for (int j = 0; j < RPL; ++j) acc_tb[j] = (uint16_t)accarg[j];
        }
        __syncwarp();
    }

    // final-only silent states on the last row
    const int NF = M->NF;
    int32_t* __restrict__ ftb = a.ftb + slot * 32;
    for (int f = 0; f < NF; ++f) {
        const int k0 = M->fin_off[f], k1 = M->fin_off[f + 1];
        double best = kNegInf;
        int arg = 0x7fffffff;
        for (int k = k0 + lane; k < k1; k += 32) {
            const int code = M->fin_src[k];
            const double sv = code < 0 ? s_fval[warp][-(code + 1)] : vfin[code];
            const double cand = sv + M->fin_w[k];
            if (cand > best) { best = cand; arg = k; }
        }
        warp_argmax_first(best, arg);
        if (lane == 0) {
            s_fval[warp][f] = best;
            ftb[f] = (best > kNegInf) ? M->fin_src[arg] : 0;
        }
        __syncwarp();
    }
    if (lane == 0) a.logp[q] = s_fval[warp][M->end_final];
}

// =============================================================================================
// on-device path reducers (hmm_utils.py:155-286): what adVNTR derives from a state path, computed
// while the backtrack walks it (from the end of the read to its start)
// =============================================================================================
struct PathReducer {
    const uint8_t* __restrict__ cls;
    const uint32_t* __restrict__ pk;
    int n, rem = 0;                      // read length; emitting states seen so far (from the end)
    int n_match = 0, repeat_bp = 0, left_bp = 0, right_bp = 0, left_hits = 0, right_hits = 0;
    int starts = 0, ends = 0, first_start = -1, last_start = -1, first_end = -1, last_end = -1;

    __device__ __forceinline__ PathReducer(const uint8_t* c, const uint32_t* p, int len) : cls(c), pk(p), n(len) {}

    __device__ __forceinline__ void visit(int state)
    {
        const int c = cls[state];
        const int kind = c & 7, part = (c >> 3) & 3;
        if (kind == 1 || kind == 2) {                       // M or I: emits read base n - rem - 1
            if (kind == 1) {
                ++n_match;
                if (part == 1 || part == 2) {
                    const int hit = packed_sym(pk, n - rem - 1) == ((c >> 5) & 3);
                    if (part == 1) left_hits += hit; else right_hits += hit;
                }
            }
            if (part == 1) ++left_bp; else if (part == 2) ++right_bp; else ++repeat_bp;
            ++rem;
        } else if (kind == 4) {                             // unit_start: >= 3 bases still to come
            if (rem >= 3) { ++starts; const int bp = n - rem; if (last_start < 0) last_start = bp; first_start = bp; }
        } else if (kind == 5) {                             // unit_end: >= 3 bases consumed
            const int bp = n - rem;
            if (bp >= 3) { ++ends; if (last_end < 0) last_end = bp; first_end = bp; }
        }
    }
    __device__ __forceinline__ void store(advhmm_read_summary* out) const
    {
        int delta = 0;
        if (first_start >= 0 && first_end >= 0 && first_end < first_start && last_start > last_end) delta = 1;
        advhmm_read_summary s;
        s.repeats = max(starts, ends) + delta;
        s.n_match = n_match; s.repeat_bp = repeat_bp; s.left_bp = left_bp; s.right_bp = right_bp;
        s.left_hits = left_hits; s.right_hits = right_hits;
        s.unit_starts_ends = starts | (ends << 16);
        *out = s;
    }
};

// =============================================================================================
// banded backtrack kernel: one thread per read
// =============================================================================================
struct BandedBtArgs {
    const Tile* tiles;
    const int32_t* order;
    int32_t chunk_base;
    int32_t n_items;            // work items in this chunk
    int32_t rpl;                // RPL the fill kernel ran with
    int32_t fp32;               // fill ran in fp32: use the float-evaluated predecessor tables
    const uint32_t* pk;
    const int64_t* pk_off;
    const int32_t* rlen;
    const double* logp;
    const uint32_t* tbw;
    size_t tbw_stride;
    const uint16_t* acc_tb;
    size_t acc_stride;
    const int32_t* ftb;
    const int32_t* item_tile;   // tile index of every work item
    int32_t* path_len;          // [n_out]
    int64_t* path_off;          // [n_out]
    int32_t* path;              // NULL: no paths wanted (summaries only)
    int64_t path_cap;
    unsigned long long* cursor; // total path entries
    advhmm_read_summary* summaries;   // NULL or [n_out]
};

template <typename Emit>
__device__ __forceinline__ void banded_walk(const DevBanded* __restrict__ M, const BandedBtArgs& a,
                                            size_t slot, int n, int sym0, Emit emit)
{
    const int P = M->P, rpl = a.rpl;
    const int32_t* __restrict__ ftb = a.ftb + slot * 32;
    int state;
    if (n == 0) {
        state = M->end;
    } else {
        int f = M->end_final, code;
        for (;;) {
            emit(M->fin_state[f]);
            code = ftb[f];
            if (code >= 0) break;
            f = -(code + 1);
        }
        int sl = code / P, c = code - sl * P, r = n;
        state = -1;
        const int nw = rpl > 5 ? 2 : 1;
        const uint32_t* tbw = a.tbw + slot * a.tbw_stride;
        const uint16_t* acc_tb = a.acc_tb + slot * a.acc_stride;
        while (r >= 1) {
            const int s = M->st[sl * M->NC + c];
            emit(s);
            const int ln = (r - 1) / rpl, j = (r - 1) - ln * rpl;   // ln = stripe * 32 + lane
            const uint32_t t = tbw[((size_t)ln * P + c) * nw + j / 5] >> (6 * (j % 5));
            // two bits per slot: bit0 = second candidate beat the first, bit1 = third beat both
            const int kI = (t & 2u) ? 2 : (int)(t & 1u);
            const int kM = (t & 8u) ? 2 : (int)((t >> 2) & 1u);
            const int kD = (t & 32u) ? 2 : (int)((t >> 4) & 1u);
            if (sl == SLOT_D) {
                if (c == M->acc_col) c = acc_tb[r - 1];
                else { sl = kD; c -= 1; }
            } else if (r == 1) {
                state = (a.fp32 ? M->tb1_f : M->tb1)[sym0 * M->S + s];
                r = 0;
            } else if (sl == SLOT_M) { sl = kM; c -= 1; r -= 1; }
            else { sl = kI; r -= 1; }
        }
    }
    // row 0: silent closure back to the start state
    int guard = M->m + 1;
    const int32_t* __restrict__ tb0 = a.fp32 ? M->tb0_f : M->tb0;
    while (state != M->start && state >= 0 && guard-- > 0) { emit(state); state = tb0[state]; }
    emit(state);
}

__global__ void __launch_bounds__(128) banded_backtrack_kernel(const BandedBtArgs a)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = i < a.n_items;
    int len = 0, q = 0, n = 0, sym0 = 0;
    size_t slot = 0;
    const DevBanded* M = nullptr;
    bool possible = false;
    if (active) {
        const int item = a.chunk_base + i;
        q = a.order[item];
        slot = (size_t)i;
        M = reinterpret_cast<const DevBanded*>(a.tiles[a.item_tile[item]].model);
        n = a.rlen[q];
        possible = a.logp[q] > kNegInf;
        if (possible) {
            if (n > 0) sym0 = packed_sym(a.pk + a.pk_off[q], 0);
            if (a.summaries) {
                // the walk starts at the model's end state and finishes at its start state; like
                // the reference's vpath[1:-1] both are silent "other" states and reduce to nothing
                PathReducer red(M->classes, a.pk + a.pk_off[q], n);
                banded_walk(M, a, slot, n, sym0, [&](int s) { ++len; red.visit(s); });
                red.store(a.summaries + q);
            } else {
                banded_walk(M, a, slot, n, sym0, [&](int) { ++len; });
            }
        } else if (a.summaries) {
            advhmm_read_summary z = {};
            z.repeats = -1;
            a.summaries[q] = z;
        }
    }
    if (!a.path) {                       // summaries only
        if (active) { a.path_len[q] = possible ? len : -1; a.path_off[q] = 0; }
        return;
    }
    // warp-aggregated allocation of output space
    const unsigned lane = threadIdx.x & 31;
    int incl = len;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= (unsigned)o) incl += v;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    unsigned long long base = 0;
    if (lane == 31 && total > 0) base = atomicAdd(a.cursor, (unsigned long long)total);
    base = __shfl_sync(0xffffffffu, base, 31);
    if (!active) return;
    if (!possible) { a.path_len[q] = -1; a.path_off[q] = 0; return; }
    const int64_t off = (int64_t)base + (incl - len);
    a.path_off[q] = off;
    if (off + len > a.path_cap) { a.path_len[q] = -2; return; }   // caller buffer too small
    a.path_len[q] = len;
    int32_t* out = a.path + off;
    int w = len;
    banded_walk(M, a, slot, n, sym0, [&](int s) { out[--w] = s; });
}

// =============================================================================================
// generic kernels (any baked model): one warp per read, rows in shared or global memory
// =============================================================================================
struct GenericArgs {
    const Tile* tiles;
    const int32_t* order;
    int32_t chunk_base;
    const uint32_t* pk;
    const int64_t* pk_off;
    const int32_t* rlen;
    double* logp;
    int32_t* end_state;         // [n_out] state the path ends in
    uint16_t* tb;               // per slot: max_n * m slots (rows 1..n)
    size_t tb_stride;
    double* rows;               // global DP rows (2 * m per slot) when they do not fit in smem
    size_t rows_stride;
    int warps;                  // warps per CTA
    int rows_in_smem;
};

template <bool FWD>
__device__ __forceinline__ double pair_lse_dev(double x, double y)
{
    // utils.pyx:72-90
    if (x == kNegInf) return y;
    if (y == kNegInf) return x;
    if (x > y) return x + log(exp(y - x) + 1.0);
    return y + log(exp(x - y) + 1.0);
}

// FWD = false: Viterbi (max, traceback); FWD = true: forward (pair_lse, no traceback)
template <bool FWD>
__global__ void __launch_bounds__(kGenericWarpsMax * 32) generic_fill_kernel(const GenericArgs a)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const Tile tile = a.tiles[blockIdx.x];
    const DevGeneric* __restrict__ G = reinterpret_cast<const DevGeneric*>(tile.model);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp >= tile.cnt) return;
    const int item = tile.first + warp;
    const int q = a.order[item];
    const size_t slot = (size_t)(item - a.chunk_base);
    const int n = a.rlen[q];
    const int m = G->m, S = G->S, K = G->K;
    double* prev;
    double* cur;
    if (a.rows_in_smem) {
        prev = reinterpret_cast<double*>(smem_raw) + (size_t)warp * 2 * m;
    } else {
        prev = a.rows + slot * a.rows_stride;
    }
    cur = prev + m;
    const int32_t* __restrict__ in_off = G->in_off;
    const int32_t* __restrict__ in_src = G->in_src;
    const double* __restrict__ in_w = G->in_w;
    const uint32_t* __restrict__ pk = a.pk + a.pk_off[q];
    uint16_t* __restrict__ tb = FWD ? nullptr : a.tb + slot * a.tb_stride;

    for (int l = lane; l < m; l += 32) prev[l] = G->v0[l];
    __syncwarp();
    // NOTE: the forward recurrence's row 0 uses pair_lse instead of max; for FWD the caller passes
    // a model whose v0 was computed with pair_lse (DevGeneric::v0 of the forward table set).
    for (int i = 0; i < n; ++i) {
        const int x = packed_sym(pk, i);
        // emitting states (hmm.pyx:2026-2042 / 1427-1444)
        for (int l = lane; l < S; l += 32) {
            const double e = G->emis[(size_t)l * K + x];
            const int k0 = in_off[l], k1 = in_off[l + 1];
            double best = kNegInf;
            int code = 0;
            for (int k = k0; k < k1; ++k) {
                if (FWD) {
                    best = pair_lse_dev<true>(best, prev[in_src[k]] + in_w[k]);
                } else {
                    const double cand = (prev[in_src[k]] + in_w[k]) + e;
                    if (cand > best) { best = cand; code = k - k0; }
                }
            }
            if (FWD) best = best + e;
            cur[l] = best;
            if (!FWD) tb[(size_t)i * m + l] = (uint16_t)code;
        }
        __syncwarp();
        // silent states, level by level (states of one level do not feed each other)
        const int nlv = G->n_levels;
        for (int L = 0; L < nlv; ++L) {
            const int lo = G->lvl_off[L], hi = G->lvl_off[L + 1];
            for (int p = lo + lane; p < hi; p += 32) {
                const int l = G->lvl_state[p];
                const int k0 = in_off[l], k1 = in_off[l + 1];
                double best = kNegInf;
                int code = 0;
                if (FWD) {
                    // pass 1 (emitting sources) and pass 2 (silent sources) are summed separately
                    // and then combined (hmm.pyx:1446-1480)
                    double acc2 = kNegInf;
                    for (int k = k0; k < k1; ++k) {
                        const int src = in_src[k];
                        const double t = cur[src] + in_w[k];
                        if (src < S) best = pair_lse_dev<true>(best, t);
                        else acc2 = pair_lse_dev<true>(acc2, t);
                    }
                    best = pair_lse_dev<true>(best, acc2);
                } else {
                    for (int k = k0; k < k1; ++k) {
                        const double cand = cur[in_src[k]] + in_w[k];
                        if (cand > best) { best = cand; code = k - k0; }
                    }
                    tb[(size_t)i * m + l] = (uint16_t)code;
                }
                cur[l] = best;
            }
            __syncwarp();
        }
        double* t = prev; prev = cur; cur = t;
    }
    // termination (hmm.pyx:2089-2098 / 1300-1313)
    if (G->finite) {
        if (lane == 0) {
            a.logp[q] = prev[G->end];
            if (!FWD) a.end_state[q] = G->end;
        }
    } else if (FWD) {
        if (lane == 0) {
            double s = kNegInf;
            for (int l = 0; l < S; ++l) s = pair_lse_dev<true>(s, prev[l]);
            a.logp[q] = s;
        }
    } else {
        double best = kNegInf;
        int arg = 0x7fffffff;
        for (int l = lane; l < m; l += 32)
            if (prev[l] > best) { best = prev[l]; arg = l; }
        warp_argmax_first(best, arg);
        if (lane == 0) { a.logp[q] = best; a.end_state[q] = (best > kNegInf) ? arg : -1; }
    }
}

struct GenericBtArgs {
    const Tile* tiles;
    const int32_t* order;
    int32_t chunk_base;
    int32_t n_items;
    const int32_t* rlen;
    const double* logp;
    const int32_t* end_state;
    const uint16_t* tb;
    size_t tb_stride;
    const int32_t* item_tile;
    int32_t* path_len;
    int64_t* path_off;
    int32_t* path;
    int64_t path_cap;
    unsigned long long* cursor;
    const uint32_t* pk;
    const int64_t* pk_off;
    advhmm_read_summary* summaries;
};

template <typename Emit>
__device__ __forceinline__ void generic_walk(const DevGeneric* __restrict__ G, const uint16_t* __restrict__ tb,
                                             int n, int end, Emit emit)
{
    const int m = G->m, S = G->S;
    int px = n, py = end;
    while (px > 0) {
        emit(py);
        const int src = G->in_src[G->in_off[py] + tb[(size_t)(px - 1) * m + py]];
        if (py < S) --px;
        py = src;
    }
    int guard = m + 1;
    while (py != G->start && py >= 0 && guard-- > 0) { emit(py); py = G->tb0[py]; }
    emit(py);
}

__global__ void __launch_bounds__(128) generic_backtrack_kernel(const GenericBtArgs a)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = i < a.n_items;
    int len = 0, q = 0, n = 0, end = 0;
    const DevGeneric* G = nullptr;
    const uint16_t* tb = nullptr;
    bool possible = false;
    if (active) {
        const int item = a.chunk_base + i;
        q = a.order[item];
        G = reinterpret_cast<const DevGeneric*>(a.tiles[a.item_tile[item]].model);
        n = a.rlen[q];
        tb = a.tb + (size_t)i * a.tb_stride;
        end = a.end_state[q];
        possible = a.logp[q] > kNegInf && end >= 0;
        if (possible) {
            if (a.summaries) {
                PathReducer red(G->classes, a.pk + a.pk_off[q], n);
                generic_walk(G, tb, n, end, [&](int s) { ++len; red.visit(s); });
                red.store(a.summaries + q);
            } else {
                generic_walk(G, tb, n, end, [&](int) { ++len; });
            }
        } else if (a.summaries) {
            advhmm_read_summary z = {};
            z.repeats = -1;
            a.summaries[q] = z;
        }
    }
    if (!a.path) {
        if (active) { a.path_len[q] = possible ? len : -1; a.path_off[q] = 0; }
        return;
    }
    const unsigned lane = threadIdx.x & 31;
    int incl = len;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= (unsigned)o) incl += v;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    unsigned long long base = 0;
    if (lane == 31 && total > 0) base = atomicAdd(a.cursor, (unsigned long long)total);
    base = __shfl_sync(0xffffffffu, base, 31);
    if (!active) return;
    if (!possible) { a.path_len[q] = -1; a.path_off[q] = 0; return; }
    const int64_t off = (int64_t)base + (incl - len);
    a.path_off[q] = off;
    if (off + len > a.path_cap) { a.path_len[q] = -2; return; }
    a.path_len[q] = len;
    int32_t* out = a.path + off;
    int w = len;
    generic_walk(G, tb, n, end, [&](int s) { out[--w] = s; });
}

// =============================================================================================
// host side: model upload
// =============================================================================================
struct BlobBuilder {
    std::vector<unsigned char> bytes;
    size_t add(const void* src, size_t n)
    {
        size_t off = (bytes.size() + 255) / 256 * 256;
        bytes.resize(off + n);
        if (n && src) memcpy(bytes.data() + off, src, n);
        return off;
    }
    template <typename T> size_t add(const std::vector<T>& v) { return add(v.data(), v.size() * sizeof(T)); }
};

// forward-algorithm row 0 (hmm.pyx:1402-1424): same closure as Viterbi's with pair_lse
std::vector<double> forward_row0(const GenericTables& g)
{
    auto lse = [](double x, double y) {
        if (x == kNegInf) return y;
        if (y == kNegInf) return x;
        if (x > y) return x + std::log(std::exp(y - x) + 1.0);
        return y + std::log(std::exp(x - y) + 1.0);
    };
    std::vector<double> f(g.m, kNegInf);
    f[g.start] = 0.0;
    for (int l = g.S; l < g.m; ++l) {
        if (l == g.start) continue;
        double acc = kNegInf;
        for (int k = g.in_off[l]; k < g.in_off[l + 1]; ++k)
            if (g.in_src[k] >= g.S) acc = lse(acc, f[g.in_src[k]] + g.in_w[k]);
        f[l] = acc;
    }
    return f;
}

}  // namespace

namespace {

int upload_model(advhmm_model* mod)
{
    advhmm_context* ctx = mod->ctx;
    const GenericTables& g = mod->cm.g;
    const BandedTables& b = mod->cm.b;
    BlobBuilder bb;
    // generic tables
    const size_t o_in_off = bb.add(g.in_off), o_in_src = bb.add(g.in_src), o_in_w = bb.add(g.in_w);
    const size_t o_emis = bb.add(g.emis), o_v0 = bb.add(g.v0), o_tb0 = bb.add(g.tb0);
    const size_t o_lvl_off = bb.add(g.lvl_off), o_lvl_state = bb.add(g.lvl_state);
    const std::vector<double> f0 = forward_row0(g);
    const size_t o_f0 = bb.add(f0);
    // banded tables
    size_t o_image = 0, o_st = 0, o_tb1 = 0, o_acc = 0, o_fs = 0, o_fo = 0, o_fsrc = 0, o_fw = 0;
    size_t o_image_f = 0, o_tb1f = 0, o_tb0f = 0, o_fwf = 0;
    int image_bytes = 0, image_f_bytes = 0;
    if (b.valid) {
        const size_t P = b.NCpad;
        std::vector<unsigned char> image((size_t)kImgBytesPerCol * P, 0);
        double* w10 = reinterpret_cast<double*>(image.data() + (size_t)kImgW * P);
        double* e2 = reinterpret_cast<double*>(image.data() + (size_t)kImgE * P);
        double* v12 = reinterpret_cast<double*>(image.data() + (size_t)kImgV1 * P);
        for (size_t c = 0; c < P; ++c) {
            for (int k = 0; k < 9; ++k) w10[c * 10 + k] = b.w[(size_t)k * P + c];
            w10[c * 10 + 9] = b.accw[c];
            for (int x = 0; x < 4; ++x)
                for (int sl = 0; sl < 2; ++sl) {       // sl: 0 = I slot, 1 = M slot
                    e2[((size_t)x * P + c) * 2 + sl] = b.e[((size_t)sl * 4 + x) * P + c];
                    v12[((size_t)x * P + c) * 2 + sl] = b.v1[((size_t)sl * 4 + x) * P + c];
                }
        }
        image_bytes = (int)image.size();
        o_image = bb.add(image);
        std::vector<int32_t> st(3 * (size_t)b.NC);
        for (int t = 0; t < 3; ++t) memcpy(st.data() + (size_t)t * b.NC, b.st[t].data(), sizeof(int32_t) * b.NC);
        o_st = bb.add(st); o_tb1 = bb.add(b.tb1); o_acc = bb.add(b.acc_src_col);
        o_fs = bb.add(b.fin_state); o_fo = bb.add(b.fin_off); o_fsrc = bb.add(b.fin_src); o_fw = bb.add(b.fin_w);
        // fp32 twin
        const BandedF32& f = mod->cm.f;
        std::vector<unsigned char> imf((size_t)kImgFBytesPerCol * P, 0);
        float* w12 = reinterpret_cast<float*>(imf.data());
        float* e2f = reinterpret_cast<float*>(imf.data() + (size_t)kImgFE * P);
        float* v1f = reinterpret_cast<float*>(imf.data() + (size_t)kImgFV1 * P);
        for (size_t c = 0; c < P; ++c) {
            for (int k = 0; k < 9; ++k) w12[c * 12 + k] = f.w[(size_t)k * P + c];
            w12[c * 12 + 9] = f.accw[c];
            for (int x = 0; x < 4; ++x)
                for (int sl = 0; sl < 2; ++sl) {
                    e2f[((size_t)x * P + c) * 2 + sl] = f.e[((size_t)sl * 4 + x) * P + c];
                    v1f[((size_t)x * P + c) * 2 + sl] = f.v1[((size_t)sl * 4 + x) * P + c];
                }
        }
        image_f_bytes = (int)imf.size();
        o_image_f = bb.add(imf); o_tb1f = bb.add(f.tb1); o_tb0f = bb.add(f.tb0); o_fwf = bb.add(f.fin_w);
    }
    const std::vector<uint8_t> zero_classes((size_t)g.m, 0);
    const size_t o_cls = bb.add(zero_classes);
    const size_t o_dg = bb.add(nullptr, sizeof(DevGeneric));
    const size_t o_dgf = bb.add(nullptr, sizeof(DevGeneric));
    const size_t o_db = bb.add(nullptr, sizeof(DevBanded));

    mod->info.kind = ADVHMM_KIND_GENERIC;
    mod->info.n_states = g.m;
    mod->info.n_edges = g.in_off[g.m];
    mod->info.n_columns = g.n_levels;
    mod->info.n_final_states = 0;
    mod->info.smem_bytes = 0;
    mod->info.max_in_degree = g.max_in_degree;
    const size_t smem_limit = ctx->device >= 0 ? ctx->smem_optin : (size_t)232448;
    if (b.valid) {
        mod->info.kind = ADVHMM_KIND_BANDED;
        mod->info.n_columns = b.NC;
        mod->info.n_final_states = (int)b.fin_state.size();
        mod->info.smem_bytes = image_bytes;
        // 0: the image does not fit in shared memory -> every read of this model takes the
        // long-read kernel, which streams the tables from global memory
        mod->banded_smem = ((size_t)image_bytes + 4096 <= smem_limit) ? image_bytes : 0;
    }
    if (ctx->device < 0) return ADVHMM_OK;   // host-only context: analysis only

    CU_TRY(cudaSetDevice(ctx->device));
    CU_TRY(mod->blob.ensure(bb.bytes.size()));
    unsigned char* base = mod->blob.as<unsigned char>();
    auto P8 = [&](size_t off) { return base + off; };
    DevGeneric dg{};
    dg.m = g.m; dg.S = g.S; dg.K = g.K; dg.start = g.start; dg.end = g.end; dg.finite = g.finite;
    dg.n_levels = g.n_levels; dg.max_in_degree = g.max_in_degree;
    dg.in_off = (const int32_t*)P8(o_in_off); dg.in_src = (const int32_t*)P8(o_in_src);
    dg.in_w = (const double*)P8(o_in_w); dg.emis = (const double*)P8(o_emis);
    dg.v0 = (const double*)P8(o_v0); dg.tb0 = (const int32_t*)P8(o_tb0);
    dg.lvl_off = (const int32_t*)P8(o_lvl_off); dg.lvl_state = (const int32_t*)P8(o_lvl_state);
    dg.classes = (const uint8_t*)P8(o_cls);
    mod->d_classes = (uint8_t*)P8(o_cls);
    memcpy(bb.bytes.data() + o_dg, &dg, sizeof dg);
    DevGeneric dgf = dg;
    dgf.v0 = (const double*)P8(o_f0);
    memcpy(bb.bytes.data() + o_dgf, &dgf, sizeof dgf);
    if (b.valid) {
        DevBanded db{};
        db.NC = b.NC; db.P = b.NCpad; db.S = b.S; db.m = g.m; db.NF = (int)b.fin_state.size();
        db.end_final = b.end_final; db.acc_col = b.acc_col; db.n_acc = (int)b.acc_src_col.size();
        db.start = g.start; db.end = g.end; db.image_bytes = image_bytes;
        db.logp_empty = g.v0[g.end];
        db.image = P8(o_image); db.st = (const int32_t*)P8(o_st); db.tb1 = (const int32_t*)P8(o_tb1);
        db.acc_src_col = (const int32_t*)P8(o_acc); db.fin_state = (const int32_t*)P8(o_fs);
        db.fin_off = (const int32_t*)P8(o_fo); db.fin_src = (const int32_t*)P8(o_fsrc);
        db.fin_w = (const double*)P8(o_fw); db.tb0 = (const int32_t*)P8(o_tb0);
        db.image_f = P8(o_image_f); db.image_f_bytes = image_f_bytes;
        db.logp_empty_f = mod->cm.f.v0[g.end];
        db.tb1_f = (const int32_t*)P8(o_tb1f); db.tb0_f = (const int32_t*)P8(o_tb0f);
        db.fin_w_f = (const float*)P8(o_fwf);
        db.classes = (const uint8_t*)P8(o_cls);
        memcpy(bb.bytes.data() + o_db, &db, sizeof db);
    }
    CU_TRY(cudaMemcpyAsync(base, bb.bytes.data(), bb.bytes.size(), cudaMemcpyHostToDevice, ctx->stream));
    CU_TRY(cudaStreamSynchronize(ctx->stream));
    mod->d_generic = reinterpret_cast<DevGeneric*>(P8(o_dg));
    mod->d_generic_fwd = reinterpret_cast<DevGeneric*>(P8(o_dgf));
    mod->d_banded = b.valid ? reinterpret_cast<DevBanded*>(P8(o_db)) : nullptr;
    return ADVHMM_OK;
}

// =============================================================================================
// host side: batch planning and launches
// =============================================================================================
// work items (result reads) of one kernel family, in launch order
struct Family {
    std::vector<int32_t> items;      // result-read ids
    std::vector<int32_t> model;      // model index of every item
    int first_item = 0;              // position of the family inside Plan::order
    int max_len = 0;
};

struct Plan {
    std::vector<int64_t> pk_off;        // [n_out]
    int64_t pk_words = 0;
    std::vector<int32_t> order;         // short-banded items, then long-banded, then generic
    std::vector<int32_t> item_tile;     // tile of every item
    std::vector<Tile> tiles;
    int max_P_short = 0, max_smem_short = 0, max_P_long = 0, max_m_generic = 0;
};

template <typename T> size_t vec_bytes(const std::vector<T>& v) { return v.size() * sizeof(T); }

// event pair bracketing one kernel launch when profiling is on (kind: 0 banded fill, 1 backtrack)
struct ProfScope {
    advhmm_context* ctx; int kind; cudaEvent_t stop = nullptr;
    ProfScope(advhmm_context* c, int k) : ctx(c), kind(k)
    {
        if (!ctx->profile) return;
        auto& ev = ctx->prof_events[kind];
        if (ctx->prof_used[kind] == ev.size()) {
            cudaEvent_t a, b;
            if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) return;
            ev.emplace_back(a, b);
        }
        auto& pr = ev[ctx->prof_used[kind]++];
        cudaEventRecord(pr.first, ctx->stream);
        stop = pr.second;
    }
    ~ProfScope() { if (stop) cudaEventRecord(stop, ctx->stream); }
};

template <int RPL, int WPB, bool ICMP>
int launch_banded_variant(advhmm_context* ctx, int grid, int smem, const BandedArgs& args, int slot)
{
    if (smem > ctx->banded_smem_set[slot]) {
        CU_TRY(cudaFuncSetAttribute(banded_fill_kernel<RPL, WPB, ICMP>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        ctx->banded_smem_set[slot] = smem;
    }
    banded_fill_kernel<RPL, WPB, ICMP><<<grid, WPB * 32, smem, ctx->stream>>>(args);
    return ADVHMM_OK;
}

int launch_banded_chunk(advhmm_context* ctx, int rpl, int grid, int smem, const BandedArgs& args)
{
    ProfScope prof(ctx, 0);
    int rc = ADVHMM_OK;
    const bool ic = ctx->launch_int_compare;
    // 8 reads per CTA, 2 CTAs per SM (126 registers/thread).  Wider CTAs (10 / 12 warps, 96 / 80
    // registers) spill and measured 25-40 % slower on B200; see profiles/r1_variants.md.
#define ADV_CASE(R)                                                                              \
    case R: rc = ic ? launch_banded_variant<R, 8, true>(ctx, grid, smem, args, R)                 \
                    : launch_banded_variant<R, 8, false>(ctx, grid, smem, args, R); break;
    switch (rpl) {
        ADV_CASE(1) ADV_CASE(2) ADV_CASE(3) ADV_CASE(4) ADV_CASE(5)
        ADV_CASE(6) ADV_CASE(7) ADV_CASE(8) ADV_CASE(9) ADV_CASE(10)
        default: return set_error(ADVHMM_EINVAL, "unsupported rows-per-lane %d", rpl);
    }
#undef ADV_CASE
    if (rc) return rc;
    CU_TRY(cudaGetLastError());
    ctx->launches++;
    return ADVHMM_OK;
}

int launch_banded_f32_chunk(advhmm_context* ctx, int rpl, int grid, int smem, const BandedArgs& args)
{
    ProfScope prof(ctx, 0);
    // the fp32 image is smaller than the fp64 one, so the fp64 size is a safe dynamic-smem request
#define ADV_CASE(R)                                                                                   \
    case R:                                                                                           \
        if (smem > ctx->banded_f32_smem_set[R]) {                                                     \
            CU_TRY(cudaFuncSetAttribute(banded_fill_f32_kernel<R>,                                    \
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, smem));          \
            ctx->banded_f32_smem_set[R] = smem;                                                       \
        }                                                                                             \
        banded_fill_f32_kernel<R><<<grid, 8 * 32, smem, ctx->stream>>>(args);                         \
        break;
    switch (rpl) {
        ADV_CASE(1) ADV_CASE(2) ADV_CASE(3) ADV_CASE(4) ADV_CASE(5)
        ADV_CASE(6) ADV_CASE(7) ADV_CASE(8) ADV_CASE(9) ADV_CASE(10)
        default: return set_error(ADVHMM_EINVAL, "unsupported rows-per-lane %d", rpl);
    }
#undef ADV_CASE
    CU_TRY(cudaGetLastError());
    ctx->launches++;
    return ADVHMM_OK;
}

struct OutPtrs {
    double* logp; int32_t* path_len; int64_t* path_off; int32_t* path; int64_t path_cap;
    unsigned long long* cursor;
    advhmm_read_summary* summaries;   // NULL unless ADVHMM_WANT_SUMMARY
};

// Runs the whole batch on the context's stream.  All pointers in `out` and d_seqs are device
// pointers.  seq_off / group_off are HOST arrays (planning metadata).
int run_batch(advhmm_context* ctx, advhmm_model* const* models, int n_models, const int64_t* group_off,
              const uint8_t* d_seqs, const int64_t* seq_off, int n_reads, uint32_t flags,
              const OutPtrs& out, bool forward, int32_t* d_bad)
{
    const bool want_path = (flags & ADVHMM_WANT_PATH) && !forward;
    const bool want_walk = want_path || (out.summaries && !forward);   // backtrack kernel needed
    const int strands = (flags & ADVHMM_BOTH_STRANDS) ? 2 : 1;
    const int n_out = n_reads * strands;
    if (n_out == 0) return ADVHMM_OK;
    const bool fp32 = (flags & ADVHMM_FP32) != 0;
    if (fp32 && forward) return set_error(ADVHMM_EUNSUPPORTED, "fp32 mode exists for Viterbi only");
    auto al = [](size_t x) { return (x + 255) / 256 * 256; };

    // ---- plan: packed-read offsets, kernel family of every result read, tiles ----------------
    Plan pl;
    pl.pk_off.resize(n_out);
    int64_t words = 0;
    for (int r = 0; r < n_reads; ++r) {
        const int64_t len = seq_off[r + 1] - seq_off[r];
        if (len < 0 || len > 0x3fffffff) return set_error(ADVHMM_EINVAL, "bad seq_off at read %d", r);
        for (int s = 0; s < strands; ++s) {
            pl.pk_off[(size_t)r * strands + s] = words;
            words += (len + 15) / 16 + 1;
        }
    }
    pl.pk_words = words;

    Family fam_short, fam_long, fam_generic;
    bool all_nonpositive = true;
    for (int gi = 0; gi < n_models; ++gi) {
        advhmm_model* mod = models[gi];
        if (!mod || mod->ctx != ctx) return set_error(ADVHMM_EINVAL, "model %d does not belong to this context", gi);
        const int64_t r0 = group_off[gi], r1 = group_off[gi + 1];
        if (r0 < 0 || r1 < r0 || r1 > n_reads) return set_error(ADVHMM_EINVAL, "bad group_off at model %d", gi);
        const bool banded = !forward && mod->d_banded && !(flags & ADVHMM_FORCE_GENERIC);
        for (int64_t r = r0; r < r1; ++r) {
            const int len = (int)(seq_off[r + 1] - seq_off[r]);
            Family* f = &fam_generic;
            if (banded) {
                if (mod->banded_smem > 0 && len <= 32 * kMaxRPL) {
                    f = &fam_short;
                    pl.max_P_short = std::max(pl.max_P_short, mod->cm.b.NCpad);
                    pl.max_smem_short = std::max(pl.max_smem_short, mod->banded_smem);
                    all_nonpositive = all_nonpositive && mod->cm.b.nonpositive;
                } else {
                    f = &fam_long;
                    pl.max_P_long = std::max(pl.max_P_long, mod->cm.b.NCpad);
                }
            } else {
                pl.max_m_generic = std::max(pl.max_m_generic, mod->cm.g.m);
            }
            f->max_len = std::max(f->max_len, len);
            for (int s = 0; s < strands; ++s) {
                f->items.push_back((int32_t)(r * strands + s));
                f->model.push_back(gi);
            }
        }
    }
    ctx->launch_int_compare = ctx->int_compare && all_nonpositive;
    if (fp32 && (!fam_long.items.empty() || !fam_generic.items.empty()))
        return set_error(ADVHMM_EUNSUPPORTED, "fp32 mode is implemented for the short-read banded kernel only "
                                              "(profile-shaped model, reads <= %d bases)", 32 * kMaxRPL);

    // generic launch geometry: as many warps per CTA as DP rows fit in shared memory
    int gwarps = kGenericWarpsMax, rows_in_smem = 1;
    if (!fam_generic.items.empty()) {
        const size_t per_warp = (size_t)pl.max_m_generic * 2 * sizeof(double);
        const size_t budget = ctx->smem_optin > 8192 ? ctx->smem_optin - 8192 : 0;
        gwarps = (int)std::min<size_t>(kGenericWarpsMax, per_warp ? budget / per_warp : kGenericWarpsMax);
        if (gwarps < 1) { gwarps = 4; rows_in_smem = 0; }
    }
    pl.order.reserve(n_out);
    pl.item_tile.reserve(n_out);
    auto make_tiles = [&](Family& f, int per_tile, int kind) {
        f.first_item = (int)pl.order.size();
        int in_tile = 0, last_model = -1;
        for (size_t i = 0; i < f.items.size(); ++i) {
            const int gi = f.model[i];
            if (in_tile == 0 || in_tile == per_tile || gi != last_model) {
                const void* dm = kind == 2 ? (forward ? (const void*)models[gi]->d_generic_fwd
                                                      : (const void*)models[gi]->d_generic)
                                           : (const void*)models[gi]->d_banded;
                pl.tiles.push_back(Tile{dm, (int32_t)pl.order.size(), 0});
                in_tile = 0;
            }
            last_model = gi;
            pl.tiles.back().cnt = ++in_tile;
            pl.item_tile.push_back((int32_t)pl.tiles.size() - 1);
            pl.order.push_back(f.items[i]);
        }
    };
    make_tiles(fam_short, ctx->banded_warps, 0);
    make_tiles(fam_long, kLongWarps, 1);
    make_tiles(fam_generic, gwarps, 2);
    const int n_short = (int)fam_short.items.size(), n_long = (int)fam_long.items.size();
    const int n_generic = (int)fam_generic.items.size();

    // ---- metadata upload (pinned staging, one H2D) -----------------------------------------
    const size_t b_seq_off = (size_t)(n_reads + 1) * sizeof(int64_t);
    const size_t b_pk_off = vec_bytes(pl.pk_off), b_order = vec_bytes(pl.order);
    const size_t b_item_tile = vec_bytes(pl.item_tile), b_tiles = vec_bytes(pl.tiles);
    const size_t o_seq_off = 0, o_pk_off = o_seq_off + al(b_seq_off), o_order = o_pk_off + al(b_pk_off);
    const size_t o_item_tile = o_order + al(b_order), o_tiles = o_item_tile + al(b_item_tile);
    const size_t meta_bytes = o_tiles + al(b_tiles);
    if (ctx->meta_done) CU_TRY(cudaEventSynchronize(ctx->meta_done));
    CU_TRY(ctx->h_meta.ensure(meta_bytes));
    CU_TRY(ctx->d_meta.ensure(meta_bytes));
    unsigned char* hm = static_cast<unsigned char*>(ctx->h_meta.p);
    memcpy(hm + o_seq_off, seq_off, b_seq_off);
    memcpy(hm + o_pk_off, pl.pk_off.data(), b_pk_off);
    memcpy(hm + o_order, pl.order.data(), b_order);
    memcpy(hm + o_item_tile, pl.item_tile.data(), b_item_tile);
    memcpy(hm + o_tiles, pl.tiles.data(), b_tiles);
    CU_TRY(cudaMemcpyAsync(ctx->d_meta.p, hm, meta_bytes, cudaMemcpyHostToDevice, ctx->stream));
    if (!ctx->meta_done) CU_TRY(cudaEventCreateWithFlags(&ctx->meta_done, cudaEventDisableTiming));
    CU_TRY(cudaEventRecord(ctx->meta_done, ctx->stream));
    unsigned char* dm = ctx->d_meta.as<unsigned char>();
    const int64_t* d_seq_off = reinterpret_cast<const int64_t*>(dm + o_seq_off);
    const int64_t* d_pk_off = reinterpret_cast<const int64_t*>(dm + o_pk_off);
    const int32_t* d_order = reinterpret_cast<const int32_t*>(dm + o_order);
    const int32_t* d_item_tile = reinterpret_cast<const int32_t*>(dm + o_item_tile);
    const Tile* d_tiles = reinterpret_cast<const Tile*>(dm + o_tiles);

    // ---- pack ------------------------------------------------------------------------------
    const size_t pk_words_al = (size_t)(pl.pk_words + 63) / 64 * 64;
    CU_TRY(ctx->d_pk.ensure(pk_words_al * sizeof(uint32_t) + (size_t)n_out * sizeof(int32_t)));
    uint32_t* d_pk = ctx->d_pk.as<uint32_t>();
    int32_t* d_rlen = reinterpret_cast<int32_t*>(d_pk + pk_words_al);
    {
        PackArgs pa{d_seqs, d_seq_off, d_pk_off, d_pk, d_rlen, d_bad, n_out, strands, 0};
        pa.n_symbols = models[0]->cm.g.K;
        pack_reads_kernel<<<n_out, 32, 0, ctx->stream>>>(pa);
        CU_TRY(cudaGetLastError());
        ctx->launches++;
    }
    if (want_path) CU_TRY(cudaMemsetAsync(out.cursor, 0, sizeof(unsigned long long), ctx->stream));

    // ---- workspace: sized once for all families (they run one after the other on the stream),
    //      chunks end on tile boundaries --------------------------------------------------------
    const int rpl = std::max(1, (fam_short.max_len + 31) / 32);
    const size_t Ps = (size_t)pl.max_P_short, Pl = (size_t)pl.max_P_long;
    const size_t nw_short = rpl > 5 ? 2 : 1;
    const size_t s_per_item = n_short ? 32 * Ps * nw_short * 4 + 3 * Ps * 8 + (size_t)32 * rpl * 2 + 32 * 4 : 0;
    const size_t stripes_max = n_long ? ((size_t)std::max(fam_long.max_len, 1) + 32 * kLongRPL - 1) / (32 * kLongRPL) : 0;
    const size_t l_tbw_words = stripes_max * 32 * Pl * 2;
    const size_t l_acc = stripes_max * 32 * kLongRPL;
    const size_t l_per_item = n_long ? l_tbw_words * 4 + 6 * Pl * 8 + l_acc * 2 + 32 * 4 : 0;
    const size_t gm = (size_t)pl.max_m_generic;
    const size_t g_tb_per = (want_walk && n_generic) ? (size_t)std::max(fam_generic.max_len, 1) * gm * sizeof(uint16_t) : 0;
    const size_t g_rows_per = (n_generic && !rows_in_smem) ? 2 * gm * sizeof(double) : 0;
    const size_t g_per_item = n_generic ? g_tb_per + g_rows_per + 8 : 0;
    auto chunk_of = [&](size_t per_item, int n_items, size_t floor_items) {
        if (!n_items) return (size_t)0;
        size_t c = std::max<size_t>(ctx->workspace_budget / per_item, floor_items);
        return std::min<size_t>(c, (size_t)n_items);
    };
    const size_t s_chunk = chunk_of(s_per_item, n_short, (size_t)kBandedWarpsMax * ctx->sm_count);
    const size_t l_chunk = chunk_of(l_per_item, n_long, (size_t)kLongWarps);
    const size_t g_chunk = chunk_of(g_per_item, n_generic, (size_t)gwarps);
    // short layout
    const size_t so_tbw = 0, so_vfin = al(s_chunk * 32 * Ps * nw_short * 4);
    const size_t so_acc = so_vfin + al(s_chunk * 3 * Ps * 8), so_ftb = so_acc + al(s_chunk * 32 * rpl * 2);
    const size_t s_bytes = so_ftb + al(s_chunk * 32 * 4);
    // long layout
    const size_t lo_tbw = 0, lo_vfin = al(l_chunk * l_tbw_words * 4), lo_carry = lo_vfin + al(l_chunk * 3 * Pl * 8);
    const size_t lo_acc = lo_carry + al(l_chunk * 3 * Pl * 8), lo_ftb = lo_acc + al(l_chunk * l_acc * 2);
    const size_t l_bytes = lo_ftb + al(l_chunk * 32 * 4);
    // generic layout
    const size_t go_tb = 0, go_rows = al(g_chunk * g_tb_per);
    const size_t g_bytes = go_rows + al(g_chunk * g_rows_per);
    const size_t o_end = std::max(s_bytes, std::max(l_bytes, g_bytes));   // end_state[n_out], whole batch
    CU_TRY(ctx->d_work.ensure(o_end + al((size_t)n_out * sizeof(int32_t))));
    unsigned char* w = ctx->d_work.as<unsigned char>();

    // [lo, hi) item ranges that start and end on tile boundaries and hold <= cap items
    auto next_chunk = [&](int lo, int end, size_t cap) {
        int hi = (int)std::min<size_t>((size_t)end, (size_t)lo + cap);
        if (hi < end)
            while (hi > lo && pl.item_tile[hi] == pl.item_tile[hi - 1]) --hi;
        if (hi == lo) {
            hi = lo + 1;
            while (hi < end && pl.item_tile[hi] == pl.item_tile[hi - 1]) ++hi;
        }
        return hi;
    };
    auto launch_backtrack = [&](int lo, int items, int bt_rpl, const uint32_t* tbw, size_t tbw_stride,
                                const uint16_t* acc, size_t acc_stride, const int32_t* ftb) -> int {
        BandedBtArgs ba{};
        ba.tiles = d_tiles; ba.order = d_order; ba.chunk_base = lo; ba.n_items = items; ba.rpl = bt_rpl;
        ba.fp32 = fp32 ? 1 : 0;
        ba.pk = d_pk; ba.pk_off = d_pk_off; ba.rlen = d_rlen; ba.logp = out.logp;
        ba.tbw = tbw; ba.tbw_stride = tbw_stride; ba.acc_tb = acc; ba.acc_stride = acc_stride; ba.ftb = ftb;
        ba.item_tile = d_item_tile;
        ba.path_len = out.path_len; ba.path_off = out.path_off; ba.path = want_path ? out.path : nullptr;
        ba.path_cap = out.path_cap; ba.cursor = out.cursor; ba.summaries = out.summaries;
        {
            ProfScope prof(ctx, 1);
            banded_backtrack_kernel<<<(items + 127) / 128, 128, 0, ctx->stream>>>(ba);
        }
        CU_TRY(cudaGetLastError());
        ctx->launches++;
        return ADVHMM_OK;
    };

    // ---- short banded reads (the Illumina path) ---------------------------------------------
    for (int lo = fam_short.first_item, end = lo + n_short; lo < end;) {
        const int hi = next_chunk(lo, end, s_chunk);
        const int tile0 = pl.item_tile[lo], tile1 = pl.item_tile[hi - 1] + 1;
        BandedArgs fa{};
        fa.tiles = d_tiles + tile0; fa.order = d_order; fa.chunk_base = lo;
        fa.pk = d_pk; fa.pk_off = d_pk_off; fa.rlen = d_rlen; fa.logp = out.logp;
        fa.tbw = reinterpret_cast<uint32_t*>(w + so_tbw); fa.tbw_stride = 32 * Ps * nw_short;
        fa.acc_tb = reinterpret_cast<uint16_t*>(w + so_acc); fa.acc_stride = 32 * rpl;
        fa.vfin = reinterpret_cast<double*>(w + so_vfin); fa.vfin_stride = 3 * Ps;
        fa.ftb = reinterpret_cast<int32_t*>(w + so_ftb);
        int rc = fp32 ? launch_banded_f32_chunk(ctx, rpl, tile1 - tile0, pl.max_smem_short, fa)
                      : launch_banded_chunk(ctx, rpl, tile1 - tile0, pl.max_smem_short, fa);
        if (rc) return rc;
        if (want_walk) {
            rc = launch_backtrack(lo, hi - lo, rpl, fa.tbw, fa.tbw_stride, fa.acc_tb, (size_t)fa.acc_stride, fa.ftb);
            if (rc) return rc;
        }
        lo = hi;
    }

    // ---- long banded reads / models too large for shared memory ------------------------------
    for (int lo = fam_long.first_item, end = lo + n_long; lo < end;) {
        const int hi = next_chunk(lo, end, l_chunk);
        const int tile0 = pl.item_tile[lo], tile1 = pl.item_tile[hi - 1] + 1;
        LongArgs la{};
        la.tiles = d_tiles + tile0; la.order = d_order; la.chunk_base = lo;
        la.pk = d_pk; la.pk_off = d_pk_off; la.rlen = d_rlen; la.logp = out.logp;
        la.tbw = reinterpret_cast<uint32_t*>(w + lo_tbw); la.tbw_stride = l_tbw_words;
        la.acc_tb = reinterpret_cast<uint16_t*>(w + lo_acc); la.acc_stride = l_acc;
        la.vfin = reinterpret_cast<double*>(w + lo_vfin); la.vfin_stride = 3 * Pl;
        la.carry = reinterpret_cast<double*>(w + lo_carry); la.carry_stride = 3 * Pl;
        la.ftb = reinterpret_cast<int32_t*>(w + lo_ftb);
        {
            ProfScope prof(ctx, 0);
            banded_long_kernel<<<tile1 - tile0, kLongWarps * 32, 0, ctx->stream>>>(la);
        }
        CU_TRY(cudaGetLastError());
        ctx->launches++;
        if (want_walk) {
            int rc = launch_backtrack(lo, hi - lo, kLongRPL, la.tbw, la.tbw_stride, la.acc_tb, la.acc_stride, la.ftb);
            if (rc) return rc;
        }
        lo = hi;
    }

    // ---- everything else: generic kernel ------------------------------------------------------
    if (n_generic > 0) {
        const int smem = rows_in_smem ? (int)(gwarps * 2 * gm * sizeof(double)) : 0;
        if (smem > 48 * 1024 && smem > ctx->generic_smem_set) {
            CU_TRY(cudaFuncSetAttribute(generic_fill_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            CU_TRY(cudaFuncSetAttribute(generic_fill_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            ctx->generic_smem_set = smem;
        }
        for (int lo = fam_generic.first_item, end = lo + n_generic; lo < end;) {
            const int hi = next_chunk(lo, end, g_chunk);
            const int items = hi - lo;
            const int tile0 = pl.item_tile[lo], tile1 = pl.item_tile[hi - 1] + 1;
            GenericArgs ga{};
            ga.tiles = d_tiles + tile0; ga.order = d_order; ga.chunk_base = lo;
            ga.pk = d_pk; ga.pk_off = d_pk_off; ga.rlen = d_rlen; ga.logp = out.logp;
            ga.end_state = reinterpret_cast<int32_t*>(w + o_end);
            ga.tb = reinterpret_cast<uint16_t*>(w + go_tb); ga.tb_stride = g_tb_per / sizeof(uint16_t);
            ga.rows = reinterpret_cast<double*>(w + go_rows); ga.rows_stride = 2 * gm;
            ga.warps = gwarps; ga.rows_in_smem = rows_in_smem;
            if (forward) generic_fill_kernel<true><<<tile1 - tile0, gwarps * 32, smem, ctx->stream>>>(ga);
            else generic_fill_kernel<false><<<tile1 - tile0, gwarps * 32, smem, ctx->stream>>>(ga);
            CU_TRY(cudaGetLastError());
            ctx->launches++;
            if (want_walk) {
                GenericBtArgs ba{};
                ba.tiles = d_tiles; ba.order = d_order; ba.chunk_base = lo; ba.n_items = items;
                ba.rlen = d_rlen; ba.logp = out.logp; ba.end_state = ga.end_state;
                ba.tb = ga.tb; ba.tb_stride = ga.tb_stride; ba.item_tile = d_item_tile;
                ba.path_len = out.path_len; ba.path_off = out.path_off; ba.path = want_path ? out.path : nullptr;
                ba.path_cap = out.path_cap; ba.cursor = out.cursor;
                ba.pk = d_pk; ba.pk_off = d_pk_off; ba.summaries = out.summaries;
                generic_backtrack_kernel<<<(items + 127) / 128, 128, 0, ctx->stream>>>(ba);
                CU_TRY(cudaGetLastError());
                ctx->launches++;
            }
            lo = hi;
        }
    }
    return ADVHMM_OK;
}

// host-buffer front end shared by the three public decoding calls
int run_host(advhmm_context* ctx, advhmm_model* const* models, int n_models, const int64_t* group_off,
             const uint8_t* seqs, const int64_t* seq_off, int n_reads, uint32_t flags, bool forward,
             double* logp, int32_t* path_len, int64_t* path_off, int32_t* path, int64_t path_cap,
             int64_t* path_total, advhmm_read_summary* summaries = nullptr)
{
    if (!ctx || ctx->device < 0) return set_error(ADVHMM_ECUDA, "this context has no CUDA device (host-only analysis context)");
    if (n_reads < 0 || !seq_off || (n_reads > 0 && (!logp || !models || n_models <= 0)))
        return set_error(ADVHMM_EINVAL, "null or negative argument");
    const bool want_path = (flags & ADVHMM_WANT_PATH) && !forward;
    const bool want_sum = (flags & ADVHMM_WANT_SUMMARY) && !forward;
    if (want_path && (!path_len || !path_off || !path_total || (path_cap > 0 && !path)))
        return set_error(ADVHMM_EINVAL, "ADVHMM_WANT_PATH needs path_len, path_off, path and path_total");
    if (want_sum && !summaries) return set_error(ADVHMM_EINVAL, "ADVHMM_WANT_SUMMARY needs a summaries buffer");
    std::lock_guard<std::mutex> lock(ctx->mu);
    CU_TRY(cudaSetDevice(ctx->device));
    const int strands = (flags & ADVHMM_BOTH_STRANDS) ? 2 : 1;
    const int n_out = n_reads * strands;
    if (path_total) *path_total = 0;
    if (n_out == 0) return ADVHMM_OK;
    const int64_t n_bases = seq_off[n_reads];
    if (seq_off[0] != 0 || n_bases < 0) return set_error(ADVHMM_EINVAL, "seq_off must start at 0 and be non-decreasing");
    if (n_bases > 0 && !seqs) return set_error(ADVHMM_EINVAL, "seqs is null");

    CU_TRY(ctx->d_seqs.ensure((size_t)n_bases + 16));
    if (n_bases) CU_TRY(cudaMemcpyAsync(ctx->d_seqs.p, seqs, (size_t)n_bases, cudaMemcpyHostToDevice, ctx->stream));
    // outputs: logp | path_len | path_off | cursor | bad
    auto al = [](size_t x) { return (x + 255) / 256 * 256; };
    const size_t o_logp = 0, o_plen = al((size_t)n_out * 8), o_poff = o_plen + al((size_t)n_out * 4);
    const size_t o_cursor = o_poff + al((size_t)n_out * 8), o_bad = o_cursor + 256;
    const size_t o_sum = o_bad + 256;
    const size_t out_bytes = o_sum + (want_sum ? al((size_t)n_out * sizeof(advhmm_read_summary)) : 0);
    CU_TRY(ctx->d_out.ensure(out_bytes));
    unsigned char* d = ctx->d_out.as<unsigned char>();
    int64_t cap = want_path ? path_cap : 0;
    if (want_path) CU_TRY(ctx->d_paths.ensure((size_t)std::max<int64_t>(cap, 1) * sizeof(int32_t)));
    OutPtrs op{reinterpret_cast<double*>(d + o_logp), reinterpret_cast<int32_t*>(d + o_plen),
               reinterpret_cast<int64_t*>(d + o_poff), ctx->d_paths.as<int32_t>(), cap,
               reinterpret_cast<unsigned long long*>(d + o_cursor),
               want_sum ? reinterpret_cast<advhmm_read_summary*>(d + o_sum) : nullptr};
    int32_t* d_bad = reinterpret_cast<int32_t*>(d + o_bad);
    CU_TRY(cudaMemsetAsync(d_bad, 0x7f, sizeof(int32_t), ctx->stream));
    int rc = run_batch(ctx, models, n_models, group_off, ctx->d_seqs.as<uint8_t>(), seq_off, n_reads, flags, op,
                       forward, d_bad);
    if (rc) { cudaStreamSynchronize(ctx->stream); return rc; }
    // results back
    CU_TRY(ctx->h_out.ensure(out_bytes));
    CU_TRY(cudaMemcpyAsync(ctx->h_out.p, d, out_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(cudaStreamSynchronize(ctx->stream));
    const unsigned char* h = static_cast<const unsigned char*>(ctx->h_out.p);
    int32_t bad;
    memcpy(&bad, h + o_bad, sizeof bad);
    if (bad != 0x7f7f7f7f) return set_error(ADVHMM_ESYMBOL, "read %d contains a symbol code outside the model alphabet", bad);
    memcpy(logp, h + o_logp, (size_t)n_out * 8);
    if (want_sum) {
        memcpy(summaries, h + o_sum, (size_t)n_out * sizeof(advhmm_read_summary));
        if (path_len && !want_path) memcpy(path_len, h + o_plen, (size_t)n_out * 4);
    }
    if (want_path) {
        memcpy(path_len, h + o_plen, (size_t)n_out * 4);
        memcpy(path_off, h + o_poff, (size_t)n_out * 8);
        unsigned long long total;
        memcpy(&total, h + o_cursor, sizeof total);
        *path_total = (int64_t)total;
        if ((int64_t)total > path_cap)
            return set_error(ADVHMM_ECAPACITY, "path buffer too small: need %lld entries, have %lld",
                             (long long)total, (long long)path_cap);
        if (total) {
            CU_TRY(cudaMemcpyAsync(path, ctx->d_paths.p, (size_t)total * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
            CU_TRY(cudaStreamSynchronize(ctx->stream));
        }
    }
    return ADVHMM_OK;
}

}  // namespace

// =============================================================================================
// fp64 add/compare issue-rate microbenchmark (the roofline denominator of the fill kernel)
// =============================================================================================
namespace {
// Each thread runs 8 independent dependent-chains of DADD; with 1024 threads per SM resident the
// fp64 pipe is saturated.  ops = threads * iters * 8.
__global__ void __launch_bounds__(256) fp64_add_peak_kernel(double* out, int iters, double seed)
{
    double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double d = seed * 1e-9;
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            a0 = __dadd_rn(a0, d); a1 = __dadd_rn(a1, d); a2 = __dadd_rn(a2, d); a3 = __dadd_rn(a3, d);
            a4 = __dadd_rn(a4, d); a5 = __dadd_rn(a5, d); a6 = __dadd_rn(a6, d); a7 = __dadd_rn(a7, d);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}
}  // namespace

// =============================================================================================
// C-ABI
// =============================================================================================
extern "C" {

const char* advhmm_last_error(void) { return g_last_error.c_str(); }
int advhmm_abi_version(void) { return ADVHMM_ABI_VERSION; }

int64_t advhmm_encode_acgt(const char* ascii, int64_t n, uint8_t* codes)
{
    static const struct Lut {
        uint8_t t[256];
        Lut() { memset(t, 255, sizeof t); t['A'] = t['a'] = 0; t['C'] = t['c'] = 1; t['G'] = t['g'] = 2; t['T'] = t['t'] = 3; }
    } lut;
    for (int64_t i = 0; i < n; ++i) {
        const uint8_t c = lut.t[(unsigned char)ascii[i]];
        if (c == 255) return i;
        codes[i] = c;
    }
    return -1;
}

int advhmm_context_create(int device, void* stream, advhmm_context** out)
{
    if (!out) return set_error(ADVHMM_EINVAL, "out is null");
    *out = nullptr;
    std::unique_ptr<advhmm_context> ctx(new (std::nothrow) advhmm_context);
    if (!ctx) return set_error(ADVHMM_ENOMEM, "out of host memory");
    ctx->device = device;
    if (device >= 0) {
        int count = 0;
        CU_TRY(cudaGetDeviceCount(&count));
        if (device >= count) return set_error(ADVHMM_EINVAL, "device %d does not exist (%d visible)", device, count);
        CU_TRY(cudaSetDevice(device));
        cudaDeviceProp prop;
        CU_TRY(cudaGetDeviceProperties(&prop, device));
        if (prop.major < 10) return set_error(ADVHMM_EUNSUPPORTED, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
        ctx->sm_count = prop.multiProcessorCount;
        ctx->smem_optin = prop.sharedMemPerBlockOptin;
        if (stream) { ctx->stream = static_cast<cudaStream_t>(stream); }
        else { CU_TRY(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)); ctx->owns_stream = true; }
        size_t free_b = 0, total_b = 0;
        CU_TRY(cudaMemGetInfo(&free_b, &total_b));
        // traceback workspace: long reads need ~40 MB each to keep every SM busy, so by default
        // half of the free device memory (capped at 96 GB) may be used; ADVHMM_WORKSPACE_MB overrides
        ctx->workspace_budget = std::min<size_t>((size_t)96 << 30, free_b / 2);
        const char* env = getenv("ADVHMM_WORKSPACE_MB");
        if (env && atoll(env) > 0) ctx->workspace_budget = (size_t)atoll(env) << 20;
        env = getenv("ADVHMM_ICMP");
        if (env) ctx->int_compare = atoi(env) != 0;
    }
    *out = ctx.release();
    return ADVHMM_OK;
}

void advhmm_context_destroy(advhmm_context* ctx)
{
    if (!ctx) return;
    if (ctx->device >= 0) {
        cudaSetDevice(ctx->device);
        cudaStreamSynchronize(ctx->stream);
        for (DevBuf* b : {&ctx->d_seqs, &ctx->d_seq_off, &ctx->d_pk, &ctx->d_meta, &ctx->d_work, &ctx->d_out, &ctx->d_paths, &ctx->d_flags}) b->release();
        ctx->h_meta.release(); ctx->h_out.release();
        if (ctx->meta_done) cudaEventDestroy(ctx->meta_done);
        for (auto& v : ctx->prof_events)
            for (auto& pr : v) { cudaEventDestroy(pr.first); cudaEventDestroy(pr.second); }
        if (ctx->owns_stream) cudaStreamDestroy(ctx->stream);
    }
    delete ctx;
}

int advhmm_context_synchronize(advhmm_context* ctx)
{
    if (!ctx || ctx->device < 0) return ADVHMM_OK;
    CU_TRY(cudaStreamSynchronize(ctx->stream));
    return ADVHMM_OK;
}

void* advhmm_context_stream(advhmm_context* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

int advhmm_context_profile(advhmm_context* ctx, int enable)
{
    if (!ctx) return set_error(ADVHMM_EINVAL, "ctx is null");
    ctx->profile = enable != 0;
    ctx->prof_used[0] = ctx->prof_used[1] = 0;
    return ADVHMM_OK;
}

int advhmm_context_profile_read(advhmm_context* ctx, double* fill_ms, int64_t* fill_launches,
                                double* backtrack_ms, int64_t* backtrack_launches)
{
    if (!ctx || ctx->device < 0) return set_error(ADVHMM_EINVAL, "no device context");
    CU_TRY(cudaStreamSynchronize(ctx->stream));
    double ms[2] = {0, 0};
    for (int k = 0; k < 2; ++k)
        for (size_t i = 0; i < ctx->prof_used[k]; ++i) {
            float t = 0;
            CU_TRY(cudaEventElapsedTime(&t, ctx->prof_events[k][i].first, ctx->prof_events[k][i].second));
            ms[k] += t;
        }
    if (fill_ms) *fill_ms = ms[0];
    if (fill_launches) *fill_launches = (int64_t)ctx->prof_used[0];
    if (backtrack_ms) *backtrack_ms = ms[1];
    if (backtrack_launches) *backtrack_launches = (int64_t)ctx->prof_used[1];
    ctx->prof_used[0] = ctx->prof_used[1] = 0;
    return ADVHMM_OK;
}

int advhmm_fp64_add_peak(advhmm_context* ctx, double* gops)
{
    if (!ctx || ctx->device < 0 || !gops) return set_error(ADVHMM_EINVAL, "no device context");
    std::lock_guard<std::mutex> lock(ctx->mu);
    CU_TRY(cudaSetDevice(ctx->device));
    const int blocks = ctx->sm_count * 8, threads = 256, iters = 4096;
    CU_TRY(ctx->d_flags.ensure((size_t)blocks * threads * sizeof(double) + 256));
    double* out = reinterpret_cast<double*>(ctx->d_flags.as<unsigned char>() + 256);
    cudaEvent_t a, b;
    CU_TRY(cudaEventCreate(&a));
    CU_TRY(cudaEventCreate(&b));
    double best = 0;
    for (int rep = 0; rep < 5; ++rep) {
        CU_TRY(cudaEventRecord(a, ctx->stream));
        fp64_add_peak_kernel<<<blocks, threads, 0, ctx->stream>>>(out, iters, 1.0 + rep);
        CU_TRY(cudaEventRecord(b, ctx->stream));
        CU_TRY(cudaStreamSynchronize(ctx->stream));
        float ms = 0;
        CU_TRY(cudaEventElapsedTime(&ms, a, b));
        const double ops = (double)blocks * threads * iters * 64.0;
        if (rep > 0) best = std::max(best, ops / (ms * 1e-3) / 1e9);
        ctx->launches++;
    }
    cudaEventDestroy(a); cudaEventDestroy(b);
    *gops = best;
    return ADVHMM_OK;
}
int64_t advhmm_context_launch_count(advhmm_context* ctx) { return ctx ? ctx->launches : 0; }

int advhmm_model_create(advhmm_context* ctx, const advhmm_model_desc* desc, advhmm_model** out)
{
    if (!ctx || !desc || !out) return set_error(ADVHMM_EINVAL, "null argument");
    *out = nullptr;
    std::unique_ptr<advhmm_model> mod(new (std::nothrow) advhmm_model);
    if (!mod) return set_error(ADVHMM_ENOMEM, "out of host memory");
    mod->ctx = ctx;
    std::string err;
    if (!compile_model(*desc, mod->cm, err)) return set_error(ADVHMM_EINVAL, "%s", err.c_str());
    if (mod->cm.g.max_in_degree > 65535) return set_error(ADVHMM_EUNSUPPORTED, "in-degree %d exceeds the 16-bit traceback slot", mod->cm.g.max_in_degree);
    std::lock_guard<std::mutex> lock(ctx->mu);
    int rc = upload_model(mod.get());
    if (rc) return rc;
    *out = mod.release();
    return ADVHMM_OK;
}

void advhmm_model_destroy(advhmm_model* model)
{
    if (!model) return;
    if (model->ctx && model->ctx->device >= 0) {
        cudaSetDevice(model->ctx->device);
        cudaStreamSynchronize(model->ctx->stream);
        model->blob.release();
    }
    delete model;
}

int advhmm_model_info_get(const advhmm_model* model, advhmm_model_info* out)
{
    if (!model || !out) return set_error(ADVHMM_EINVAL, "null argument");
    *out = model->info;
    return ADVHMM_OK;
}

int advhmm_viterbi_batch(advhmm_model* model, const uint8_t* seqs, const int64_t* seq_off, int32_t n_reads,
                         uint32_t flags, double* logp, int32_t* path_len, int64_t* path_off,
                         int32_t* path, int64_t path_cap, int64_t* path_total)
{
    if (!model) return set_error(ADVHMM_EINVAL, "model is null");
    const int64_t group_off[2] = {0, n_reads};
    advhmm_model* models[1] = {model};
    return run_host(model->ctx, models, 1, group_off, seqs, seq_off, n_reads, flags & ~ADVHMM_DEVICE_BUFFERS, false,
                    logp, path_len, path_off, path, path_cap, path_total);
}

int advhmm_log_probability_batch(advhmm_model* model, const uint8_t* seqs, const int64_t* seq_off,
                                 int32_t n_reads, uint32_t flags, double* logp)
{
    if (!model) return set_error(ADVHMM_EINVAL, "model is null");
    const int64_t group_off[2] = {0, n_reads};
    advhmm_model* models[1] = {model};
    return run_host(model->ctx, models, 1, group_off, seqs, seq_off, n_reads,
                    flags & ~(ADVHMM_DEVICE_BUFFERS | ADVHMM_WANT_PATH | ADVHMM_BOTH_STRANDS), true,
                    logp, nullptr, nullptr, nullptr, 0, nullptr);
}

int advhmm_viterbi_multi(advhmm_context* ctx, advhmm_model* const* models, int32_t n_models,
                         const int64_t* group_off, const uint8_t* seqs, const int64_t* seq_off,
                         int32_t n_reads, uint32_t flags, double* logp, int32_t* path_len,
                         int64_t* path_off, int32_t* path, int64_t path_cap, int64_t* path_total)
{
    return advhmm_viterbi_multi_summary(ctx, models, n_models, group_off, seqs, seq_off, n_reads,
                                        flags & ~ADVHMM_WANT_SUMMARY, logp, path_len, path_off, path, path_cap,
                                        path_total, nullptr);
}

int advhmm_model_set_state_classes(advhmm_model* model, const uint8_t* classes)
{
    if (!model || !classes) return set_error(ADVHMM_EINVAL, "null argument");
    if (!model->ctx || model->ctx->device < 0) return ADVHMM_OK;      // host-only context: nothing to upload
    std::lock_guard<std::mutex> lock(model->ctx->mu);
    CU_TRY(cudaSetDevice(model->ctx->device));
    CU_TRY(cudaMemcpyAsync(model->d_classes, classes, (size_t)model->cm.g.m, cudaMemcpyHostToDevice, model->ctx->stream));
    CU_TRY(cudaStreamSynchronize(model->ctx->stream));
    return ADVHMM_OK;
}

int advhmm_viterbi_multi_summary(advhmm_context* ctx, advhmm_model* const* models, int32_t n_models,
                                 const int64_t* group_off, const uint8_t* seqs, const int64_t* seq_off,
                                 int32_t n_reads, uint32_t flags, double* logp, int32_t* path_len,
                                 int64_t* path_off, int32_t* path, int64_t path_cap, int64_t* path_total,
                                 advhmm_read_summary* summaries)
{
    if (!ctx || !models || !group_off || n_models <= 0) return set_error(ADVHMM_EINVAL, "null argument");
    if (!(flags & ADVHMM_DEVICE_BUFFERS))
        return run_host(ctx, models, n_models, group_off, seqs, seq_off, n_reads, flags, false,
                        logp, path_len, path_off, path, path_cap, path_total, summaries);
    // device-resident buffers: asynchronous on the context's stream, nothing is copied back
    if (ctx->device < 0) return set_error(ADVHMM_ECUDA, "this context has no CUDA device");
    const bool want_path = flags & ADVHMM_WANT_PATH;
    const bool want_sum = (flags & ADVHMM_WANT_SUMMARY) != 0;
    if (!seq_off || !logp || (want_path && (!path_len || !path_off || !path_total)) ||
        (want_sum && (!summaries || !path_len || !path_off)))
        return set_error(ADVHMM_EINVAL, "null argument");
    std::lock_guard<std::mutex> lock(ctx->mu);
    CU_TRY(cudaSetDevice(ctx->device));
    CU_TRY(ctx->d_flags.ensure(256));
    int32_t* d_bad = ctx->d_flags.as<int32_t>();
    CU_TRY(cudaMemsetAsync(d_bad, 0x7f, sizeof(int32_t), ctx->stream));
    OutPtrs op{logp, path_len, path_off, path, want_path ? path_cap : 0,
               reinterpret_cast<unsigned long long*>(path_total), want_sum ? summaries : nullptr};
    return run_batch(ctx, models, n_models, group_off, seqs, seq_off, n_reads, flags & ~ADVHMM_DEVICE_BUFFERS, op,
                     false, d_bad);
}

}  // extern "C"
