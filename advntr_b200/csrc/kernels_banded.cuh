// kernels_banded.cuh: banded Viterbi fill kernels (fp64, fp32, long reads), on-device path reducers, backtrack -- part of libadvhmm.so (see advhmm.cu for the overview)
#pragma once
#include "kernels_common.cuh"

#ifndef ADV_STEP_UNROLL
#define ADV_STEP_UNROLL 3      // unroll factor of the steady phase of banded_fill_kernel's column loop (RPL <= 5)
#endif

namespace {
constexpr int kStepUnroll = ADV_STEP_UNROLL;
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

template <int SH>
__device__ __forceinline__ double max3_first(double a0, double a1, double a2, uint32_t& bits)
{
    double m;
    // written in PTX so that each compare costs one DSETP, one 64-bit select and one predicated add.
    // (A variant that compares the bit patterns on the integer pipe -- every DP value is <= 0, so the
    // order of the patterns is the reverse order of the values -- measured 20 % slower: it overloads
    // the ALU pipe that already carries the selects; profiles/r1_variants.md.)
    //
    // Evaluation order: in every use a0 is the candidate coming from an I slot, which is the value
    // that arrives last (the I slots chain down the rows of a column and across lanes), so a1 and a2
    // are compared first and the late value meets their winner in ONE compare + select.  The winner
    // is still the first maximal candidate in in-edge order: a2 replaces a1 only if strictly
    // greater, and their winner replaces a0 only if strictly greater.
    // bits: (1 << SH) = "not a0", (2 << SH) = "a2 beat a1" (meaningful only with the first bit set).
    asm("{\n\t"
        ".reg .pred p1, p2;\n\t"
        ".reg .f64 t;\n\t"
        "setp.gt.f64 p2, %4, %3;\n\t"
        "selp.f64 t, %4, %3, p2;\n\t"
        "@p2 or.b32 %1, %1, %6;\n\t"
        "setp.gt.f64 p1, t, %2;\n\t"
        "selp.f64 %0, t, %2, p1;\n\t"
        "@p1 or.b32 %1, %1, %5;\n\t"
        "}"
        : "=d"(m), "+r"(bits)
        : "d"(a0), "d"(a1), "d"(a2), "n"(1u << SH), "n"(2u << SH));
    return m;
}

// ALIGNED: the read length is a multiple of RPL, so the last read position is the last row of a
// lane and its values can be stored from fixed registers.
template <int RPL, bool ALIGNED>
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
#pragma unroll
    for (int j = 0; j < RPL; ++j) { cI[j] = cM[j] = cD[j] = acc[j] = kNegInf; }
    double bI = kNegInf, bM = kNegInf, bD = kNegInf;   // row above my block, previous column
    // lane-skewed views: index them with the step t (column c = t - lane).  The empty asm keeps
    // ptxas from re-deriving these addresses inside the loop.
    uint32_t* tbw_t = tbw - (ptrdiff_t)lane * NW;
    double* vfin_t = vfin - 3 * lane;                   // interleaved: vfin[3 c + {0, 1, 2}] = I, M, D of column c
    const uint32_t store_flag = (uint32_t)(lane == ln);
    uint32_t w_t = s_base - (uint32_t)lane * 80u;       // + 80 t  -> w10[c]
    uint32_t e_t[RPL];                                  // + 16 t  -> e2[sym_j][c]
#pragma unroll
    for (int j = 0; j < RPL; ++j) { e_t[j] = eaddr[j] - (uint32_t)lane * 16u; asm volatile("" : "+r"(e_t[j])); }
    // lane 0 owns the first read position, whose I and M values come from the first-row table v1: its "emission"
    // address of row 0 points there, so the step loads {vI, vM} of row 1 with the load every lane issues anyway
    if (lane == 0) e_t[0] += v1_delta;
    asm volatile("" : "+l"(tbw_t), "+l"(vfin_t), "+r"(w_t));

    // One column step.  GUARD: some lanes are outside the column range (ramp-up / ramp-down of the
    // skewed wavefront).  The steady phase, where every lane has a column, runs without the test and
    // without the divergence scope around it, so that consecutive steps form one basic block and the
    // tail of the I-slot chain of step t can overlap the M / D work of step t+1 (kStepUnroll).
    auto step = [&](const int t, auto guard_c, auto acc_c) {
        constexpr bool GUARD = decltype(guard_c)::value;
        constexpr bool ACC = decltype(acc_c)::value;       // some lane may sit on the collector's column in this step
        // the row above my block at column c was finished by lane-1 in the previous step
        const double uI0 = shfl_up_f64(cI[RPL - 1], 1);
        const double uM0 = shfl_up_f64(cM[RPL - 1], 1);
        const double uD0 = shfl_up_f64(cD[RPL - 1], 1);
        const int c = t - lane;
        if (GUARD && (c < 0 || c >= NC || lane >= nl)) return;

        const uint32_t wa = w_t + (uint32_t)t * 80u;
        const double2 w01 = lds128(wa), w23 = lds128(wa + 16), w45 = lds128(wa + 32);
        const double2 w67 = lds128(wa + 48), w89 = lds128(wa + 64);
        const double wII = w01.x, wIM = w01.y, wID = w23.x, wMI = w23.y, wMM = w45.x, wMD = w45.y;
        const double wDI = w67.x, wDM = w67.y, wDD = w89.x, aw = w89.y;
        const uint32_t cb = (uint32_t)t * 16u;

        // M and D slots depend only on values of the previous column / previous step
        double nM[RPL], nD[RPL], eIr[RPL], eM0 = 0.0;
        uint32_t word[NW];
#pragma unroll
        for (int k = 0; k < NW; ++k) word[k] = 0;
        static_for<0, RPL>([&](auto jc) {
            constexpr int j = decltype(jc)::value;
            const double2 e = lds128(e_t[j] + cb);             // {eI, eM}; lane 0, row 0: {vI, vM} of the first row
            eIr[j] = e.x;
            if (j == 0) eM0 = e.y;
            const double oI = j ? cI[j ? j - 1 : 0] : bI, oM = j ? cM[j ? j - 1 : 0] : bM, oD = j ? cD[j ? j - 1 : 0] : bD;
            nM[j] = max3_first<6 * (j % 5) + 2>((oI + wMI) + e.y, (oM + wMM) + e.y, (oD + wMD) + e.y, word[j / 5]);
            nD[j] = max3_first<6 * (j % 5) + 4>(cI[j] + wDI, cM[j] + wDM, cD[j] + wDD, word[j / 5]);
        });
        if (lane == 0) nM[0] = eM0;                            // first read position: from the row-0 table (e_t[0] above)
        // collector (end_repeating_pattern_match): D of its column is the best unit_end so far.
        // Both cases are rare per lane (1 and `copies` columns of NC), hence real branches.
        if (ACC && c == acc_col) {
#pragma unroll
            for (int j = 0; j < RPL; ++j) nD[j] = acc[j];
        }
        if (aw > kNegInf) {                                    // a unit_end column
#pragma unroll
            for (int j = 0; j < RPL; ++j) {
                const double cand = nD[j] + aw;
                // the source column goes straight to the workspace (the last improvement is the one that
                // stays): no register per row, a predicated store instead of a select
                if (cand > acc[j]) { acc[j] = cand; acc_tb[j] = (uint16_t)c; }
            }
        }
        // I slots chain down the rows of this column
        double uI = uI0, uM = uM0, uD = uD0;
        static_for<0, RPL>([&](auto jc) {
            constexpr int j = decltype(jc)::value;
            double vI = max3_first<6 * (j % 5)>((uI + wII) + eIr[j], (uM + wIM) + eIr[j], (uD + wID) + eIr[j], word[j / 5]);
            if (j == 0 && lane == 0) vI = eIr[0];
            uI = vI; uM = nM[j]; uD = nD[j];
            cI[j] = vI; cM[j] = nM[j]; cD[j] = nD[j];
        });
        bI = uI0; bM = uM0; bD = uD0;
        if (NW == 1) tbw_t[t] = word[0];
        else reinterpret_cast<uint2*>(tbw_t)[t] = make_uint2(word[0], word[NW - 1]);
        if (ALIGNED) {
            // predicated stores, no branch: the step stays one basic block up to the collector
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %4, 0;\n\t"
                         "@p st.global.f64 [%0], %1;\n\t@p st.global.f64 [%0+8], %2;\n\t@p st.global.f64 [%0+16], %3;\n\t}"
                         :: "l"(vfin_t + 3 * t), "d"(cI[RPL - 1]), "d"(cM[RPL - 1]), "d"(cD[RPL - 1]), "r"(store_flag) : "memory");
        } else {
            // the last read position sits in row jn of lane ln; jn is warp-uniform, so this is a uniform jump to
            // three predicated stores from fixed registers (a runtime select of the row cost 25 instructions per step)
            double* const dst = vfin_t + 3 * t;
            static_for<0, RPL - 1>([&](auto jc) {
                constexpr int j = decltype(jc)::value;
                if (jn == j)
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %4, 0;\n\t"
                                 "@p st.global.f64 [%0], %1;\n\t@p st.global.f64 [%0+8], %2;\n\t@p st.global.f64 [%0+16], %3;\n\t}"
                                 :: "l"(dst), "d"(cI[j]), "d"(cM[j]), "d"(cD[j]), "r"(store_flag) : "memory");
            });
        }
    };
    // Lanes >= nl have no read positions: in the steady phase they run along (results land in their
    // own registers and in traceback words nobody reads), in the guarded phases they skip.
    const int NCu = __reduce_max_sync(0xffffffffu, NC);
    const int steps = __reduce_max_sync(0xffffffffu, NC + nl - 1);
    const int t_steady = steps < 31 ? steps : 31, t_down = NCu > t_steady ? NCu : t_steady;
    // lane L reaches the collector's column at step acc_col + L: only the 32 steps from acc_col on carry the select
    const int accu = __reduce_max_sync(0xffffffffu, acc_col);
    const int t_acc0 = min(max(accu, t_steady), t_down), t_acc1 = min(max(accu + 32, t_steady), t_down);
#pragma unroll 1
    for (int t = 0; t < t_steady; ++t) step(t, std::true_type{}, std::true_type{});
    // unrolled x3 for reads up to 160 bp (+4.5 % measured, 2..6 alike, 8 outgrows the instruction cache);
    // 6 and 7 rows per lane spill at 128 registers and stay rolled; 8..10 rows (255 registers) take x2 (+1..2 %)
#pragma unroll (RPL <= 5 ? kStepUnroll : RPL >= 8 ? 2 : 1)
    for (int t = t_steady; t < t_acc0; ++t) step(t, std::false_type{}, std::false_type{});
#pragma unroll 1
    for (int t = t_acc0; t < t_acc1; ++t) step(t, std::false_type{}, std::true_type{});
#pragma unroll (RPL <= 5 ? kStepUnroll : RPL >= 8 ? 2 : 1)
    for (int t = t_acc1; t < t_down; ++t) step(t, std::false_type{}, std::false_type{});
#pragma unroll 1
    for (int t = t_down; t < steps; ++t) step(t, std::true_type{}, std::true_type{});
    // (which unit_end fed the collector, per read position, was stored as it changed: acc_tb, read by
    // the backtrack only and only for rows whose collector value is finite)
}

template <int RPL, int WPB>
// 8..10 rows per lane (reads of 225..320 bases) need 184..210 registers: one 8-warp CTA per SM without
// spills beats sixteen warps at 128 registers with spills (+11 % at 250 bp, +33 % at 300 bp); up to 7 rows
// the 128-register build wins (profiles/r1_variants.md)
__global__ void __launch_bounds__(WPB * 32, (WPB > 8 || RPL >= 8) ? 1 : 2)
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
    if (__reduce_max_sync(0xffffffffu, (int)(warp >= tile.cnt))) return;

    const int item = tile.first + warp;
    const int q = a.order[item];
    const size_t slot = (size_t)(item - a.chunk_base);
    const int n = __reduce_max_sync(0xffffffffu, a.rlen[q]);
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
        banded_sweep<RPL, true>(NC, P, nl, ln, jn, lane, s_base, symbits, M->acc_col, tbw, acc_tb, vfin);
    else
        banded_sweep<RPL, false>(NC, P, nl, ln, jn, lane, s_base, symbits, M->acc_col, tbw, acc_tb, vfin);
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
            const double sv = code < 0 ? s_fval[warp][-(code + 1)] : vfin[3 * (code % P) + code / P];
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
// banded forward kernel (HiddenMarkovModel.log_probability / _forward, hmm.pyx:1258-1313,
// 1371-1484): the same register wavefront with log-sum-exp in place of max and no traceback.
// The reference folds a state's candidates with pair_lse in in-edge order (utils.pyx:72-90) and,
// for silent states, combines the emitting-source and silent-source partial sums at the end;
// here the three candidates of a slot are summed in one shifted log-sum-exp.  The contract for
// forward is 1e-9 relative (the device exp/log differ from glibc's in the last bits anyway), not
// bit identity.  Row 1 comes from the host-evaluated forward first-row table (global memory,
// read by lane 0 only); reads longer than 32 * kMaxRPL or models whose image does not fit in
// shared memory take the generic forward kernel.
// =============================================================================================
// not inlined on purpose: one column step evaluates 3 * RPL of these, and with exp / log expanded
// in place the loop body outgrows the instruction cache (ncu: no_instruction was the top stall)
__device__ __noinline__ double lse3(double a, double b, double c)
{
    const double m = fmax(a, fmax(b, c));
    if (m == kNegInf) return kNegInf;
    return m + log(exp(a - m) + exp(b - m) + exp(c - m));
}
__device__ __noinline__ double lse2(double a, double b)
{
    const double m = fmax(a, b);
    if (m == kNegInf) return kNegInf;
    return m + log(exp(a - m) + exp(b - m));
}

template <int RPL, int WPB>
__global__ void __launch_bounds__(WPB * 32, 2)
banded_forward_kernel(const BandedArgs a)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ double s_fval[WPB][32];

    const Tile tile = a.tiles[blockIdx.x];
    const DevBanded* __restrict__ M = reinterpret_cast<const DevBanded*>(tile.model);
    const int P = M->P, NC = M->NC;
    if (threadIdx.x == 0) mbar_init(&s_bar, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        // weights + emissions only: the Viterbi first-row table at the end of the image is not used
        const uint32_t bytes = (uint32_t)kImgV1 * (uint32_t)P;
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
        if (lane == 0) a.logp[q] = M->logp_empty_fwd;
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
    double* __restrict__ vfin = a.vfin + slot * a.vfin_stride;
    const uint32_t s_base = smem_u32(smem_raw);
    uint32_t eaddr[RPL];
#pragma unroll
    for (int j = 0; j < RPL; ++j)
        eaddr[j] = s_base + (uint32_t)kImgE * P + ((symbits >> (2 * j)) & 3u) * (uint32_t)(16 * P);
    const double2* __restrict__ f1 = reinterpret_cast<const double2*>(M->f1) + (size_t)(symbits & 3u) * P;   // lane 0: [c] -> {I, M}

    double cI[RPL], cM[RPL], cD[RPL], acc[RPL];
#pragma unroll
    for (int j = 0; j < RPL; ++j) cI[j] = cM[j] = cD[j] = acc[j] = kNegInf;
    double bI = kNegInf, bM = kNegInf, bD = kNegInf;
    const int acc_col = M->acc_col;
    const int steps = NC + nl - 1;
#pragma unroll 1
    for (int t = 0; t < steps; ++t) {
        const double uI0 = shfl_up_f64(cI[RPL - 1], 1);
        const double uM0 = shfl_up_f64(cM[RPL - 1], 1);
        const double uD0 = shfl_up_f64(cD[RPL - 1], 1);
        const int c = t - lane;
        if (c < 0 || c >= NC || lane >= nl) continue;
        const uint32_t wa = s_base + (uint32_t)c * 80u;
        const double2 w01 = lds128(wa), w23 = lds128(wa + 16), w45 = lds128(wa + 32);
        const double2 w67 = lds128(wa + 48), w89 = lds128(wa + 64);
        const double wII = w01.x, wIM = w01.y, wID = w23.x, wMI = w23.y, wMM = w45.x, wMD = w45.y;
        const double wDI = w67.x, wDM = w67.y, wDD = w89.x, aw = w89.y;
        double nM[RPL], nD[RPL], eIr[RPL];
#pragma unroll
        for (int j = 0; j < RPL; ++j) {
            const double2 e = lds128(eaddr[j] + (uint32_t)c * 16u);   // {eI, eM}
            eIr[j] = e.x;
            const double oI = j ? cI[j ? j - 1 : 0] : bI, oM = j ? cM[j ? j - 1 : 0] : bM, oD = j ? cD[j ? j - 1 : 0] : bD;
            nM[j] = lse3(oI + wMI, oM + wMM, oD + wMD) + e.y;           // emission after the sum (hmm.pyx:1444)
            nD[j] = lse3(cI[j] + wDI, cM[j] + wDM, cD[j] + wDD);
        }
        double first_I = kNegInf;
        if (lane == 0) {
            const double2 f = __ldg(f1 + c);
            nM[0] = f.y;
            first_I = f.x;
        }
        if (c == acc_col) {
#pragma unroll
            for (int j = 0; j < RPL; ++j) nD[j] = acc[j];
        }
        if (aw > kNegInf) {
#pragma unroll
            for (int j = 0; j < RPL; ++j) acc[j] = lse2(acc[j], nD[j] + aw);
        }
        double uI = uI0, uM = uM0, uD = uD0;
#pragma unroll
        for (int j = 0; j < RPL; ++j) {
            double vI = lse3(uI + wII, uM + wIM, uD + wID) + eIr[j];
            if (j == 0 && lane == 0) vI = first_I;
            uI = vI; uM = nM[j]; uD = nD[j];
            cI[j] = vI; cM[j] = nM[j]; cD[j] = nD[j];
        }
        bI = uI0; bM = uM0; bD = uD0;
        if (lane == ln) {
            double fI = cI[0], fM = cM[0], fD = cD[0];
#pragma unroll
            for (int j = 1; j < RPL; ++j)
                if (j == jn) { fI = cI[j]; fM = cM[j]; fD = cD[j]; }
            vfin[c] = fI; vfin[P + c] = fM; vfin[2 * P + c] = fD;
        }
    }
    __syncwarp();

    // final-only silent states on the last row: log-sum-exp over the candidate list, across the warp
    const int NF = M->NF;
    for (int f = 0; f < NF; ++f) {
        const int k0 = M->fin_off[f], k1 = M->fin_off[f + 1];
        double mx = kNegInf;
        for (int k = k0 + lane; k < k1; k += 32) {
            const int code = M->fin_src[k];
            const double sv = code < 0 ? s_fval[warp][-(code + 1)] : vfin[code];
            mx = fmax(mx, sv + M->fin_w[k]);
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) mx = fmax(mx, shfl_xor_f64(mx, off));
        double sum = 0.0;
        if (mx > kNegInf)
            for (int k = k0 + lane; k < k1; k += 32) {
                const int code = M->fin_src[k];
                const double sv = code < 0 ? s_fval[warp][-(code + 1)] : vfin[code];
                sum += exp(sv + M->fin_w[k] - mx);
            }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) sum += shfl_xor_f64(sum, off);
        if (lane == 0) s_fval[warp][f] = (mx > kNegInf) ? mx + log(sum) : kNegInf;
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
        "setp.gt.f32 p2, %4, %3;\n\t"
        "selp.f32 t, %4, %3, p2;\n\t"
        "@p2 or.b32 %1, %1, %6;\n\t"
        "setp.gt.f32 p1, t, %2;\n\t"
        "selp.f32 %0, t, %2, p1;\n\t"
        "@p1 or.b32 %1, %1, %5;\n\t"
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
#pragma unroll
    for (int j = 0; j < RPL; ++j) { cI[j] = cM[j] = cD[j] = acc[j] = kNI; }
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
                if (cand > acc[j]) { acc[j] = cand; acc_tb[j] = (uint16_t)c; }   // stored as it changes (see banded_sweep)
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
// Same recurrence, tables, rows per lane (5) and traceback encoding as banded_fill_kernel<5, 8>, but
//   * a read is cut into stripes of 32 * 5 = 160 positions that one warp sweeps one after the other;
//     the last position of a stripe is carried to the next stripe through a per-read buffer in
//     global memory (3 doubles per column, updated in place: writes trail reads by >= 16 columns)
//     that lanes 0..15 stage 16 columns ahead in shared memory;
//   * the model image (1.3 MB for the 100-copy model) does not fit in shared memory, and it does not
//     have to: at step t the 32 lanes of a warp touch columns t-31 .. t only.  Every warp keeps a
//     RING of 64 columns (+ 16 mirrored, see below) of the image in shared memory -- weights 80 B and
//     emissions 4 x 16 B per column -- that lane 0 refills 16 columns at a time with TMA bulk copies
//     (cp.async.bulk + one mbarrier per 16-column slot), one block ahead of the lane-0 front.  The
//     column loop is the short-read kernel's: LDS.128 from the ring instead of global loads.
//   * ring addressing without a wrap test per access: the steps run in blocks of 16; within a block a
//     lane walks 16 consecutive ring slots from ((16 b - lane) & 63), and the first 16 slots are
//     mirrored behind the 64th (the copy of every fourth column block is issued twice), so the walk
//     never wraps.
//   * a 20 kb read needs ~100 MB of traceback, so few reads can be in flight; the parallelism comes from
//     INSIDE a read instead: `wpr` warps of a CTA share one read and take its stripes round-robin, each
//     one trailing the warp that sweeps the stripe above by three 16-column blocks (it needs that
//     stripe's last row).  The hand-over is the carry buffer; how far a stripe has got is published in
//     shared memory by the lane that writes the carry (st.release) and polled by the consumer (ld.acquire).
// =============================================================================================
struct LongArgs {
    const Tile* tiles;
    const int32_t* order;
    int32_t chunk_base;
    const uint32_t* pk;
    const int64_t* pk_off;
    const int32_t* rlen;
    double* logp;
    uint32_t* tbw;              // per slot: stripes_max * 32 * Pmax words
    size_t tbw_stride;
    uint16_t* acc_tb;           // per slot: stripes_max * 32 * RPL entries
    size_t acc_stride;
    double* vfin;               // per slot 3 * Pmax
    size_t vfin_stride;
    double* carry;              // per slot 3 * Pmax
    size_t carry_stride;
    int32_t* ftb;               // per slot 32
    int32_t wpr;                // warps per read: 1, 2, 4 or 8 (the stripes of a read are dealt round-robin)
};

#ifndef ADV_LONG_UNROLL
#define ADV_LONG_UNROLL 2      // unroll factor of the steady 16-step blocks of banded_long_kernel (4: -4 %, the loop outgrows the instruction cache)
#endif
constexpr int kLongUnroll = ADV_LONG_UNROLL;
constexpr int kLongRPL = 5;
constexpr int kLongWarps = 8;
constexpr int kRingBlk = 16;                       // columns per TMA refill
constexpr int kRingCols = 64 + kRingBlk;           // 4 slots + the mirror of slot 0

// The kernel is written once for both arithmetic types: double (the bit-exact default) and float (the optional
// ADVHMM_FP32 mode, float image of 112 bytes per column, see banded_fill_f32_kernel).
template <typename T> struct LongT;
template <> struct LongT<double> {
    static constexpr int WB = 80, EB = 16;         // ring bytes per column: ten weights; {eI, eM} of one symbol
    static constexpr int kE = kImgE, kV1 = kImgV1; // byte offsets (per column) of the image sections
    static __device__ __forceinline__ const unsigned char* image(const DevBanded* M) { return M->image; }
    static __device__ __forceinline__ double empty(const DevBanded* M) { return M->logp_empty; }
    static __device__ __forceinline__ const double* fin_w(const DevBanded* M) { return M->fin_w; }
};
template <> struct LongT<float> {
    static constexpr int WB = 48, EB = 8;
    static constexpr int kE = kImgFE, kV1 = kImgFV1;
    static __device__ __forceinline__ const unsigned char* image(const DevBanded* M) { return M->image_f; }
    static __device__ __forceinline__ float empty(const DevBanded* M) { return M->logp_empty_f; }
    static __device__ __forceinline__ const float* fin_w(const DevBanded* M) { return M->fin_w_f; }
};

template <typename T>
struct __align__(128) LongRing {                   // per warp
    unsigned char w[kRingCols * LongT<T>::WB];     // w[slot][10 (+2 pad for float)]
    unsigned char e[4][kRingCols * LongT<T>::EB];  // e2[sym][slot][2]
    T carry[2][3][kRingBlk];                       // carried row of the previous stripe, double-buffered
    uint64_t bar[4];
};

__device__ __forceinline__ void ldg_pair(const unsigned char* p, double& x, double& y)
{
    asm volatile("ld.global.nc.v2.f64 {%0, %1}, [%2];" : "=d"(x), "=d"(y) : "l"(p));
}
__device__ __forceinline__ void ldg_pair(const unsigned char* p, float& x, float& y)
{
    asm volatile("ld.global.nc.v2.f32 {%0, %1}, [%2];" : "=f"(x), "=f"(y) : "l"(p));
}
__device__ __forceinline__ void lds_pair(uint32_t addr, double& x, double& y) { const double2 v = lds128(addr); x = v.x; y = v.y; }
__device__ __forceinline__ void lds_pair(uint32_t addr, float& x, float& y) { const float2 v = lds64f(addr); x = v.x; y = v.y; }
// the ten weights of a column: wII wIM wID | wMI wMM wMD | wDI wDM wDD | accw
__device__ __forceinline__ void lds_weights(uint32_t wa, double (&w)[10])
{
    const double2 a = lds128(wa), b = lds128(wa + 16), c = lds128(wa + 32), d = lds128(wa + 48), e = lds128(wa + 64);
    w[0] = a.x; w[1] = a.y; w[2] = b.x; w[3] = b.y; w[4] = c.x; w[5] = c.y; w[6] = d.x; w[7] = d.y; w[8] = e.x; w[9] = e.y;
}
__device__ __forceinline__ void lds_weights(uint32_t wa, float (&w)[10])
{
    const float4 a = lds128f(wa), b = lds128f(wa + 16), c = lds128f(wa + 32);
    w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w; w[8] = c.x; w[9] = c.y;
}
template <int SH> __device__ __forceinline__ double max3_t(double a0, double a1, double a2, uint32_t& bits) { return max3_first<SH>(a0, a1, a2, bits); }
template <int SH> __device__ __forceinline__ float max3_t(float a0, float a1, float a2, uint32_t& bits) { return max3_first_f32<SH>(a0, a1, a2, bits); }
__device__ __forceinline__ double shfl_up_t(double v) { return shfl_up_f64(v, 1); }
__device__ __forceinline__ float shfl_up_t(float v) { return __shfl_up_sync(0xffffffffu, v, 1); }
__device__ __forceinline__ double ldcg_t(const double* p) { return __ldcg(p); }
__device__ __forceinline__ float ldcg_t(const float* p) { return __ldcg(p); }
__device__ __forceinline__ void warp_argmax_first_t(double& v, int& idx) { warp_argmax_first(v, idx); }
__device__ __forceinline__ void warp_argmax_first_t(float& v, int& idx)
{
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, v, off);
        const int oi = __shfl_xor_sync(0xffffffffu, idx, off);
        if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
    }
}

// FWD (double only): the forward algorithm (log_probability) on the same stripes -- log-sum-exp in place of the
// max, emission added after the sum (hmm.pyx:1444), no traceback; first row from the forward table M->f1.
template <typename T, bool FWD = false>
__global__ void __launch_bounds__(kLongWarps * 32, 2)
banded_long_kernel(const LongArgs a)
{
    static_assert(!FWD || std::is_same<T, double>::value, "forward runs in double");
    using LT = LongT<T>;
    constexpr int RPL = kLongRPL, H = 32 * RPL, B = kRingBlk, WB = LT::WB, EB = LT::EB;
    const T kNI = (T)kNegInf;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ T s_fval[kLongWarps][32];
    __shared__ long long s_prog[kLongWarps];       // (stripe << 32 | carried columns written) of every warp

    const Tile tile = a.tiles[blockIdx.x];
    const DevBanded* __restrict__ M = reinterpret_cast<const DevBanded*>(tile.model);
    const int P = M->P, NC = M->NC, acc_col = M->acc_col;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wpr = a.wpr;
    if (threadIdx.x < kLongWarps) s_prog[threadIdx.x] = -1;
    __syncthreads();
    const int rd = warp / wpr, sub = warp - rd * wpr;              // read of this CTA, position among its warps
    if (rd >= tile.cnt) return;
    const int item = tile.first + rd;
    const int q = a.order[item];
    const size_t slot = (size_t)(item - a.chunk_base);
    const int n = a.rlen[q];
    if (n == 0) {
        if (lane == 0 && sub == 0) a.logp[q] = FWD ? M->logp_empty_fwd : (double)LT::empty(M);
        return;
    }
    // progress flags: release store by the lane that wrote the carried values, acquire load by the consumer
    auto publish = [&](long long v) {
        asm volatile("st.release.cta.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&s_prog[warp])), "l"(v) : "memory");
    };
    // the warp that sweeps the stripe above mine, and a wait until it has written `need` carried columns of it
    const int pw = rd * wpr + (sub + wpr - 1) % wpr;
    auto wait_carry = [&](int stripe_above, int need) {
        if (lane == 0) {
            const long long want = ((long long)stripe_above << 32) | (long long)need;
            const uint32_t addr = smem_u32(&s_prog[pw]);
            long long have;
            for (;;) {
                asm volatile("ld.acquire.cta.shared::cta.b64 %0, [%1];" : "=l"(have) : "r"(addr) : "memory");
                if (have >= want) break;
                __nanosleep(64);
            }
        }
        __syncwarp();
    };
    LongRing<T>& ring = reinterpret_cast<LongRing<T>*>(smem_raw)[warp];
    const unsigned char* __restrict__ img = LT::image(M);
    const unsigned char* __restrict__ img_e = img + (size_t)LT::kE * P;
    const unsigned char* __restrict__ img_v1 = FWD ? reinterpret_cast<const unsigned char*>(M->f1) : img + (size_t)LT::kV1 * P;
    const uint32_t* __restrict__ pk = a.pk + a.pk_off[q];
    T* __restrict__ vfin = reinterpret_cast<T*>(a.vfin + slot * a.vfin_stride);        // (the float build uses half of it)
    T* __restrict__ carry = reinterpret_cast<T*>(a.carry + slot * a.carry_stride);
    const int n_stripes = (n + H - 1) / H;
    const int last_word = (n + 15) / 16;
    const int n_cblocks = P / B;                                   // P is a multiple of 16
    const uint32_t w_ring = smem_u32(ring.w), e_ring = smem_u32(ring.e);

    if (lane == 0)
        for (int k = 0; k < 4; ++k) mbar_init(&ring.bar[k], 1);
    __syncwarp();
    uint32_t phase = 0;                                            // bit k: parity the next wait on slot k expects
    auto wait_block = [&](int j) {
        const int sl = j & 3;
        mbar_wait(&ring.bar[sl], (phase >> sl) & 1u);
        phase ^= 1u << sl;
    };

    // lane 0: queue the copies of column block j of the image into ring slot j & 3
    auto issue_block = [&](int j) {
        const int sl = j & 3;
        const uint32_t bytes = (uint32_t)(B * WB + 4 * B * EB) * (sl == 0 ? 2u : 1u);
        mbar_expect_tx(&ring.bar[sl], bytes);
        tma_bulk_g2s(&ring.w[(size_t)sl * B * WB], img + (size_t)j * B * WB, B * WB, &ring.bar[sl]);
#pragma unroll
        for (int x = 0; x < 4; ++x)
            tma_bulk_g2s(&ring.e[x][(size_t)sl * B * EB], img_e + (size_t)x * EB * P + (size_t)j * B * EB, B * EB, &ring.bar[sl]);
        if (sl == 0) {                                              // the mirror behind slot 3
            tma_bulk_g2s(&ring.w[(size_t)64 * WB], img + (size_t)j * B * WB, B * WB, &ring.bar[sl]);
#pragma unroll
            for (int x = 0; x < 4; ++x)
                tma_bulk_g2s(&ring.e[x][(size_t)64 * EB], img_e + (size_t)x * EB * P + (size_t)j * B * EB, B * EB, &ring.bar[sl]);
        }
    };

    // One stripe.  FIRST / LAST are compile-time: the middle stripes of a read -- all but two of ~100 -- carry
    // neither the first-row override nor the last-row stores through their column loop.
    auto sweep_stripe = [&](const int s, auto first_c, auto last_c) {
        constexpr bool FIRST = decltype(first_c)::value, LAST = decltype(last_c)::value;
        const int rows = min(H, n - s * H);
        const int nl = (rows + RPL - 1) / RPL;
        const int ln = (rows - 1) / RPL, jn = (rows - 1) % RPL;
        uint32_t symbits;
        {
            const int bit = 2 * (s * H + lane * RPL);
            const int w = bit >> 5, sh = bit & 31;
            const uint32_t lo = (w <= last_word) ? pk[w] : 0u;
            const uint32_t hi = (w + 1 <= last_word) ? pk[w + 1] : 0u;
            symbits = __funnelshift_r(lo, hi, sh);
        }
        // ring address of e2[sym_j][slot of my column at the start of the current block]; moved on by 16 slots
        // (mod 64) from block to block
        uint32_t e_blk[RPL];
        const uint32_t slot_init = (uint32_t)(0 - lane) & 63u;
#pragma unroll
        for (int j = 0; j < RPL; ++j)
            e_blk[j] = e_ring + ((symbits >> (2 * j)) & 3u) * (uint32_t)(kRingCols * EB) + slot_init * (uint32_t)EB;
        uint32_t w_blk = w_ring + slot_init * (uint32_t)WB;
        uint32_t slot0 = slot_init;
        const size_t v1_off = (size_t)(symbits & 3u) * (size_t)EB * P;

        uint32_t* tbw_t = a.tbw + slot * a.tbw_stride + ((size_t)(s * 32 + lane) * P) - lane;   // index with the step t
        uint16_t* __restrict__ acc_tb = a.acc_tb + slot * a.acc_stride + (size_t)s * H + lane * RPL;
        T* carry_t = carry - 31;                                    // lane 31 writes column t - 31: index with t
        asm volatile("" : "+l"(tbw_t), "+l"(carry_t));

        T cI[RPL], cM[RPL], cD[RPL], acc[RPL];
#pragma unroll
        for (int j = 0; j < RPL; ++j) { cI[j] = cM[j] = cD[j] = acc[j] = kNI; }
        T bI = kNI, bM = kNI, bD = kNI;
        const bool first_row = FIRST && lane == 0;
        const bool carried = !FIRST && lane == 0;
        const bool carry_out = !LAST && lane == 31;

        // prologue: column block 0 on its way, carried columns 0..15 staged
        __syncwarp();
        if (lane == 0) issue_block(0);
        if (!FIRST) {
            if (wpr > 1) wait_carry(s - 1, min(B, NC));
            if (lane < B) {
#pragma unroll
                for (int k = 0; k < 3; ++k) ring.carry[0][k][lane] = ldcg_t(&carry[(size_t)k * P + lane]);
            }
        }

        const int steps = NC + nl - 1;
        const int n_sblocks = (steps + B - 1) / B;
        for (int sb = 0; sb < n_sblocks; ++sb) {
            const int t0 = sb * B;
            __syncwarp();                                           // every lane is done with the slot refilled next
            if (sb + 1 < n_cblocks) {
                if (lane == 0) issue_block(sb + 1);
                if (!FIRST) {
                    // (keeping the three loaded values in registers across the 16 steps to hide the L2 round trip
                    // measured 6 % SLOWER: the kernel sits at its 128-register cap)
                    if (wpr > 1) wait_carry(s - 1, min((sb + 2) * B, NC));
                    if (lane < B) {
                        const int col = (sb + 1) * B + lane;
#pragma unroll
                        for (int k = 0; k < 3; ++k) ring.carry[(sb + 1) & 1][k][lane] = ldcg_t(&carry[(size_t)k * P + col]);
                    }
                }
            }
            if (sb < n_cblocks) wait_block(sb);
            __syncwarp();
            const T* cr = &ring.carry[sb & 1][0][0];

            auto step = [&](const int i, auto guard_c, auto acc_c) {
                constexpr bool GUARD = decltype(guard_c)::value;
                constexpr bool ACC = decltype(acc_c)::value;        // some lane may reach the collector's column in this block
                const int t = t0 + i;
                T uI0 = shfl_up_t(cI[RPL - 1]);
                T uM0 = shfl_up_t(cM[RPL - 1]);
                T uD0 = shfl_up_t(cD[RPL - 1]);
                const int c = t - lane;
                if (GUARD && (c < 0 || c >= NC || lane >= nl)) return;
                if (!FIRST && carried) { uI0 = cr[i]; uM0 = cr[B + i]; uD0 = cr[2 * B + i]; }
                // this lane's 16 consecutive ring slots of the block (mirror: no wrap inside the block)
                T wv[10];
                lds_weights(w_blk + (uint32_t)i * (uint32_t)WB, wv);
                const T wII = wv[0], wIM = wv[1], wID = wv[2], wMI = wv[3], wMM = wv[4], wMD = wv[5];
                const T wDI = wv[6], wDM = wv[7], wDD = wv[8], aw = wv[9];

                T nM[RPL], nD[RPL], eIr[RPL];
                uint32_t word = 0;
                static_for<0, RPL>([&](auto jc) {
                    constexpr int j = decltype(jc)::value;
                    T eI, eM;
                    lds_pair(e_blk[j] + (uint32_t)i * (uint32_t)EB, eI, eM);
                    eIr[j] = eI;
                    const T oI = j ? cI[j ? j - 1 : 0] : bI, oM = j ? cM[j ? j - 1 : 0] : bM, oD = j ? cD[j ? j - 1 : 0] : bD;
                    if constexpr (FWD) {
                        nM[j] = lse3(oI + wMI, oM + wMM, oD + wMD) + eM;
                        nD[j] = lse3(cI[j] + wDI, cM[j] + wDM, cD[j] + wDD);
                    } else {
                        nM[j] = max3_t<6 * j + 2>((oI + wMI) + eM, (oM + wMM) + eM, (oD + wMD) + eM, word);
                        nD[j] = max3_t<6 * j + 4>(cI[j] + wDI, cM[j] + wDM, cD[j] + wDD, word);
                    }
                });
                if (FIRST && first_row) {
                    T fI, fM;
                    ldg_pair(img_v1 + v1_off + (size_t)c * (size_t)EB, fI, fM);
                    nM[0] = fM;
                    eIr[0] = fI;
                }
                if (ACC && c == acc_col) {
#pragma unroll
                    for (int j = 0; j < RPL; ++j) nD[j] = acc[j];
                }
                if (aw > kNI) {
#pragma unroll
                    for (int j = 0; j < RPL; ++j) {
                        const T cand = nD[j] + aw;
                        if constexpr (FWD) acc[j] = lse2(acc[j], cand);
                        else if (cand > acc[j]) { acc[j] = cand; acc_tb[j] = (uint16_t)c; }   // stored as it changes (see banded_sweep)
                    }
                }
                T uI = uI0, uM = uM0, uD = uD0;
                static_for<0, RPL>([&](auto jc) {
                    constexpr int j = decltype(jc)::value;
                    T vI;
                    if constexpr (FWD) vI = lse3(uI + wII, uM + wIM, uD + wID) + eIr[j];
                    else vI = max3_t<6 * j>((uI + wII) + eIr[j], (uM + wIM) + eIr[j], (uD + wID) + eIr[j], word);
                    if (FIRST && j == 0 && first_row) vI = eIr[0];
                    uI = vI; uM = nM[j]; uD = nD[j];
                    cI[j] = vI; cM[j] = nM[j]; cD[j] = nD[j];
                });
                bI = uI0; bM = uM0; bD = uD0;
                if constexpr (!FWD) tbw_t[t] = word;
                if (!LAST && carry_out) {                           // full stripe: lane 31 owns its last position
                    carry_t[t] = cI[RPL - 1]; carry_t[(size_t)P + t] = cM[RPL - 1]; carry_t[(size_t)2 * P + t] = cD[RPL - 1];
                }
                if (LAST && lane == ln) {
                    T fI = cI[0], fM = cM[0], fD = cD[0];
#pragma unroll
                    for (int j = 1; j < RPL; ++j)
                        if (j == jn) { fI = cI[j]; fM = cM[j]; fD = cD[j]; }
                    vfin[c] = fI; vfin[(size_t)P + c] = fM; vfin[(size_t)2 * P + c] = fD;
                }
            };
            // steady blocks: every lane has a column at every step (lanes beyond the stripe's last row run
            // along, as in banded_sweep); boundary blocks test per step
            // (the lanes of a block sit on columns t0 - 31 .. t0 + B - 1: the three blocks that can touch the
            // collector's column take the guarded loop, which carries its select)
            if (t0 >= 31 && t0 + B <= NC && t0 + B <= steps && (acc_col < t0 - 31 || acc_col >= t0 + B)) {
#pragma unroll (kLongUnroll)
                for (int i = 0; i < B; ++i) step(i, std::false_type{}, std::false_type{});
            } else {
                const int i_end = min(B, steps - t0);
#pragma unroll 1
                for (int i = 0; i < i_end; ++i) step(i, std::true_type{}, std::true_type{});
            }
            // the next block starts 16 ring slots further on (mod 64)
            {
                const bool wrap = slot0 + B >= 64u;
                const uint32_t dw = wrap ? (uint32_t)(B * WB) - 64u * (uint32_t)WB : (uint32_t)(B * WB);
                const uint32_t de = wrap ? (uint32_t)(B * EB) - 64u * (uint32_t)EB : (uint32_t)(B * EB);
                slot0 = (slot0 + B) & 63u;
                w_blk += dw;
#pragma unroll
                for (int j = 0; j < RPL; ++j) e_blk[j] += de;
            }
            // tell the warp below how many carried columns of this stripe are in memory: after step t lane 31
            // has written columns 0 .. t - 31.  Published by the lane that wrote them (release).
            if (!LAST && wpr > 1 && carry_out) {
                const int done = min(max(t0 + B - 31, 0), NC);
                publish(((long long)s << 32) | (long long)done);
            }
        }
        if (!LAST && wpr > 1 && carry_out) publish(((long long)s << 32) | (long long)NC);
        // a block queued beyond the last one this stripe waited for (the prefetch runs one block ahead of the
        // steps; cannot happen while steps >= columns, kept for symmetry): drain it, every copy is waited for once
        if (n_sblocks < n_cblocks) wait_block(n_sblocks);
        __syncwarp();
    };
    for (int s = sub; s < n_stripes; s += wpr) {
        const bool first = (s == 0), last = (s == n_stripes - 1);
        if (first && last) sweep_stripe(s, std::true_type{}, std::true_type{});
        else if (first) sweep_stripe(s, std::true_type{}, std::false_type{});
        else if (last) sweep_stripe(s, std::false_type{}, std::true_type{});
        else sweep_stripe(s, std::false_type{}, std::false_type{});
    }

    // final-only silent states on the last row: the warp that swept the last stripe holds it
    if (sub != (n_stripes - 1) % wpr) return;
    const int NF = M->NF;
    int32_t* __restrict__ ftb = a.ftb + slot * 32;
    const T* __restrict__ fin_w = LT::fin_w(M);
    if constexpr (FWD) {
        for (int f = 0; f < NF; ++f) {
            const int k0 = M->fin_off[f], k1 = M->fin_off[f + 1];
            double mx = kNegInf;
            for (int k = k0 + lane; k < k1; k += 32) {
                const int code = M->fin_src[k];
                const double sv = code < 0 ? s_fval[warp][-(code + 1)] : vfin[code];
                mx = fmax(mx, sv + M->fin_w[k]);
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) mx = fmax(mx, shfl_xor_f64(mx, off));
            double sum = 0.0;
            if (mx > kNegInf)
                for (int k = k0 + lane; k < k1; k += 32) {
                    const int code = M->fin_src[k];
                    const double sv = code < 0 ? s_fval[warp][-(code + 1)] : vfin[code];
                    sum += exp(sv + M->fin_w[k] - mx);
                }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) sum += shfl_xor_f64(sum, off);
            if (lane == 0) s_fval[warp][f] = (mx > kNegInf) ? mx + log(sum) : kNegInf;
            __syncwarp();
        }
        if (lane == 0) a.logp[q] = s_fval[warp][M->end_final];
        return;
    }
    for (int f = 0; f < NF; ++f) {
        const int k0 = M->fin_off[f], k1 = M->fin_off[f + 1];
        T best = kNI;
        int arg = 0x7fffffff;
        for (int k = k0 + lane; k < k1; k += 32) {
            const int code = M->fin_src[k];
            const T sv = code < 0 ? s_fval[warp][-(code + 1)] : vfin[code];
            const T cand = sv + fin_w[k];
            if (cand > best) { best = cand; arg = k; }
        }
        warp_argmax_first_t(best, arg);
        if (lane == 0) {
            s_fval[warp][f] = best;
            ftb[f] = (best > kNI) ? M->fin_src[arg] : 0;
        }
        __syncwarp();
    }
    if (lane == 0) a.logp[q] = (double)s_fval[warp][M->end_final];
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
    // long reads: the walk is a serial pointer chase of n + columns steps; instead of walking twice (count,
    // then write) the states of the first walk are kept here, per work item, and copied out
    int32_t* scratch;           // NULL or n_items * scratch_stride entries
    int64_t scratch_stride;
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
            // two bits per slot (max3_first): bit0 = the first candidate lost, bit1 = the third beat the second
            const int kI = (t & 1u) ? 1 + (int)((t >> 1) & 1u) : 0;
            const int kM = (t & 4u) ? 1 + (int)((t >> 3) & 1u) : 0;
            const int kD = (t & 16u) ? 1 + (int)((t >> 5) & 1u) : 0;
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
        if (possible && a.scratch && a.path) {
            // one walk: states land at the END of this item's scratch region, in path order
            if (n > 0) sym0 = packed_sym(a.pk + a.pk_off[q], 0);
            int32_t* __restrict__ sc = a.scratch + (size_t)i * a.scratch_stride;
            const int64_t cap = a.scratch_stride;
            if (a.summaries) {
                PathReducer red(M->classes, a.pk + a.pk_off[q], n);
                banded_walk(M, a, slot, n, sym0, [&](int s) { if (len < cap) sc[cap - 1 - len] = s; ++len; red.visit(s); });
                red.store(a.summaries + q);
            } else {
                banded_walk(M, a, slot, n, sym0, [&](int s) { if (len < cap) sc[cap - 1 - len] = s; ++len; });
            }
        } else if (possible) {
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
    if (a.scratch && len <= a.scratch_stride) {
        const int32_t* __restrict__ sc = a.scratch + (size_t)i * a.scratch_stride + (a.scratch_stride - len);
#pragma unroll 4
        for (int k = 0; k < len; ++k) out[k] = sc[k];
        return;
    }
    int w = len;
    banded_walk(M, a, slot, n, sym0, [&](int s) { out[--w] = s; });
}

}  // namespace
