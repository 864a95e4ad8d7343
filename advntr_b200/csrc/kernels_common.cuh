// kernels_common.cuh: device helpers (mbarrier / TMA bulk copy, shuffles, packed reads) and the pack kernel -- part of libadvhmm.so (see advhmm.cu for the overview)
#pragma once
#include "engine_types.cuh"

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
    int read_base;              // index of this batch's first read in the caller's numbering (error reports)
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
    if (bad) atomicMin(a.bad, r + a.read_base);
}

}  // namespace
