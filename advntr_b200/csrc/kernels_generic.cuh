// kernels_generic.cuh: generic CSR kernels for any baked model (Viterbi, forward, backtrack) -- part of libadvhmm.so (see advhmm.cu for the overview)
#pragma once
#include "kernels_banded.cuh"

namespace {
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

}  // namespace
