// kernels_kfilter.cuh -- GPU keyword pre-filter, the step right before the Viterbi hot path
// (SURVEY.md section 8f rank 3).  Replaces the Aho-Corasick scan of the reference's
// `adVNTR-Filtering` binary (/root/reference/filtering/main.cc:229-300): for every unmapped read
// and every locus, count the occurrences of the locus's keywords in the read.
//
// An Aho-Corasick automaton reports, at every text position, every keyword that ends there.  With
// the keywords grouped by length ("classes") that is an exact k-mer lookup per class and position,
// which parallelises over POSITIONS instead of over reads:
//
//   * the reads are one flat byte stream (as the caller passes them, no separators), cut into warp
//     tiles of kKfTile bytes.  The kernel is persistent (one CTA per SM) and every warp is its own
//     pipeline: lane 0 stages tile + halo (the k_max-1 bytes before it) into the warp's shared-memory
//     buffer with ONE TMA bulk copy (cp.async.bulk, completion on the warp's mbarrier) while the
//     previous tile is scanned (two buffers per warp); no CTA-wide barrier after start-up.  The warp
//     converts the tile in place to 3-bit symbol codes (A,C,G,T = 0..3, anything else = 4: the
//     reference's char_to_num, main.cc:43-54; lower case is "anything else" there too).
//   * a lane scans 64 consecutive end positions.  k <= 21: the k-mer is a 3k-bit integer rolled by
//     shift/or (exact key).  k > 21: a 64-bit polynomial rolling hash, verified against the
//     keyword text on a hit (exact result either way).
//   * level 0: every position tests one bit of a 64 KB bitmap held in shared memory (copied once
//     per CTA); level 1: the survivors test one bit of a bitmap in global memory (64 bits
//     per keyword, L2 resident).  Eight positions are hashed before the first
//     test, so the probes of a group overlap.
//   * a filter hit is queued in shared memory and resolved when the warp's tile is done, one hit per
//     lane (resolving it where it occurs would run the whole warp through the lookup for one lane):
//     open-addressing table (key, class) -> list of loci that
//     own the keyword; which read the position belongs to is found by a binary search in seq_off
//     restricted to the reads that intersect the tile (device-computed tile_first[]); k-mers that
//     would span a read boundary are rejected there -- the main loop carries no boundary logic.
//   * (read, locus) occurrence counters live in a second open-addressing table (atomicCAS insert,
//     atomicAdd count): exact for any number of loci per read; a full table is reported to the
//     host, which retries with a larger one.  A compaction kernel emits the triples with
//     count >= min_matches.
// Integer / byte work: no tensor cores.  Per text byte the scan issues ~30 integer instructions,
// so the bound is the integer issue rate, not HBM (bench: tools/kbench_filter.py).
#pragma once
#include "kernels_common.cuh"

namespace {

constexpr int kKfMaxWarps = 32;                          // warps of the one persistent CTA per SM
constexpr int kKfPerThread = 64;
constexpr int kKfTile = 32 * kKfPerThread;               // 2,048 bytes of text per warp tile
constexpr int kKfMaxClasses = 16;                        // distinct keyword lengths per filter
constexpr int kKfMaxK = 4096;                            // longest keyword (halo in shared memory)
constexpr int kKfGroup = 8;                              // positions hashed before the first test
constexpr int kKfL0Bits = 1 << 19;                       // first-level filter: 64 KB bitmap in shared memory
constexpr int kKfL0Bytes = kKfL0Bits / 8;
constexpr unsigned long long kKfMul = 0x9E3779B97F4A7C15ull;
constexpr unsigned long long kKfBase = 0x100000001B3ull; // rolling-hash base (odd)

struct KfClass {
    int k;
    int exact;                       // 1: 3-bit packed key (k <= 21); 0: rolling hash + verification
    unsigned long long key_mask;     // 3k low bits (exact)
    unsigned long long salt;         // separates the classes in the shared tables
    unsigned long long bk;           // kKfBase^k (rolling hash: weight of the symbol that leaves)
};

struct KfEntry {                     // 32 bytes = one sector per probe
    unsigned long long key;
    uint32_t loci_off, loci_cnt;     // loci_cnt == 0: empty slot
    uint32_t text_off, cls;
    unsigned long long pad;
};

struct DevKFilter {
    int n_classes, halo;             // halo: bytes staged before a tile (multiple of 16, >= k_max - 1)
    KfClass cls[kKfMaxClasses];
    const uint32_t* l0;              // first-level bitmap (kKfL0Bits bits), copied to shared memory by every CTA
    const uint32_t* bloom;
    uint32_t bloom_shift, pad;       // word index = hash >> bloom_shift
    const KfEntry* table;
    unsigned long long table_mask;
    const int32_t* loci;
    const uint8_t* text;             // symbol codes of the hashed keywords (verification)
};

__host__ __device__ __forceinline__ int kf_code(unsigned char ch)
{
    return ch == 'A' ? 0 : ch == 'C' ? 1 : ch == 'G' ? 2 : ch == 'T' ? 3 : 4;
}

// h = key * M + salt (the salt separates the length classes and rides on the multiply-add for free).
// a = high word: first-level bit = its low bits, Bloom word = its top bits; the Bloom bit inside the
// word comes from the top of the low word.
__host__ __device__ __forceinline__ unsigned long long kf_hash(unsigned long long key, unsigned long long salt, uint32_t& a)
{
    const unsigned long long h = key * kKfMul + salt;
    a = (uint32_t)(h >> 32);
    return h;
}
__host__ __device__ __forceinline__ uint32_t kf_bloom_bit(unsigned long long h)
{
    return (uint32_t)h >> 27;
}
__host__ __device__ __forceinline__ unsigned long long kf_slot(unsigned long long h, unsigned long long mask)
{
    return (h >> 20) & mask;
}

// four ASCII bytes -> four symbol codes
__device__ __forceinline__ uint32_t kf_codes4(uint32_t w)
{
    const uint32_t mA = __vcmpeq4(w, 0x41414141u), mC = __vcmpeq4(w, 0x43434343u);
    const uint32_t mG = __vcmpeq4(w, 0x47474747u), mT = __vcmpeq4(w, 0x54545454u);
    return (mC & 0x01010101u) | (mG & 0x02020202u) | (mT & 0x03030303u) | (~(mA | mC | mG | mT) & 0x04040404u);
}

struct KfScanArgs {
    DevKFilter f;
    const unsigned char* seqs;          // ASCII reads, back to back; 16-byte aligned, readable to the next multiple of 16
    const int64_t* seq_off;             // [n_reads + 1]
    const int32_t* tile_first;          // [n_tiles + 1]: read that owns the first byte of the tile; last = n_reads - 1
    int64_t n_bases;
    int32_t n_reads, n_tiles;
    unsigned long long* cnt_keys;       // (read << 32 | locus), ~0ull = empty
    uint32_t* cnt_vals;
    unsigned long long cnt_mask;
    int32_t* overflow;                  // set to 1 when the counter table is full
};

__device__ __forceinline__ void kf_count(const KfScanArgs& a, int r, int32_t locus)
{
    const unsigned long long ck = ((unsigned long long)(uint32_t)r << 32) | (uint32_t)locus;
    unsigned long long cs = (ck * kKfMul >> 17) & a.cnt_mask;
    for (unsigned long long probes = 0; probes <= a.cnt_mask; ++probes) {
        const unsigned long long old = atomicCAS(&a.cnt_keys[cs], ~0ull, ck);
        if (old == ~0ull || old == ck) { atomicAdd(&a.cnt_vals[cs], 1u); return; }
        cs = (cs + 1) & a.cnt_mask;
    }
    *a.overflow = 1;
}

constexpr int kKfQueue = 128;                            // pending filter hits per warp and tile

// One filter hit (end position p, length class c): exact lookup, read lookup, boundary check, counting.
// The k-mer is re-read from the warp's tile buffer, so the scan loop keeps no per-position state.
__device__ __noinline__ void kf_resolve(const KfScanArgs& a, int c, int64_t p,
                                        const unsigned char* __restrict__ sm, int64_t base, int tile)
{
    const KfClass cl = a.f.cls[c];
    const int k = cl.k;
    if (p >= a.n_bases || p - k + 1 < 0) return;
    const unsigned char* __restrict__ s = sm + (p - k + 1 - base);     // the k-mer's symbol codes
    unsigned long long key = 0;
    if (cl.exact) for (int j = 0; j < k; ++j) key = (key << 3) | s[j];
    else          for (int j = 0; j < k; ++j) key = key * kKfBase + (s[j] + 1u);
    uint32_t aa;
    const unsigned long long h = kf_hash(key, cl.salt, aa);
    unsigned long long slot = kf_slot(h, a.f.table_mask);
    for (;;) {
        const KfEntry* e = a.f.table + slot;
        const uint32_t cnt = e->loci_cnt;
        if (cnt == 0) return;                                          // empty slot: not a keyword
        if (e->key == key && e->cls == (uint32_t)c) {
            bool same = true;
            if (!cl.exact) {
                const uint8_t* __restrict__ t = a.f.text + e->text_off;
                for (int j = 0; j < k && same; ++j) same = (s[j] == t[j]);
            }
            if (same) {
                // the read that owns position p: the last one with seq_off[r] <= p
                int lo = a.tile_first[tile], hi = a.tile_first[tile + 1];
                while (lo < hi) {
                    const int mid = (lo + hi + 1) >> 1;
                    if (a.seq_off[mid] <= p) lo = mid; else hi = mid - 1;
                }
                if (p - k + 1 >= a.seq_off[lo]) {                      // else it starts in the previous read
                    const uint32_t off = e->loci_off;
                    for (uint32_t j = 0; j < cnt; ++j) kf_count(a, lo, a.f.loci[off + j]);
                }
                return;                                                // (keyword, class) is unique in the table
            }
        }
        slot = (slot + 1) & a.f.table_mask;
    }
}

// Filter hits are rare per lane but not per warp; resolving them where they occur would run the whole
// warp through the lookup for one lane's sake.  They are queued per warp (bit j of `hits`: end position
// p_first + j) and resolved together, one hit per lane, when the tile is done.
__device__ __forceinline__ void kf_enqueue(const KfScanArgs& a, int c, uint32_t hits, int64_t p_first, int64_t tile0,
                                           uint32_t* q, uint32_t* q_cnt, const unsigned char* sm, int64_t base, int tile)
{
    for (; hits; hits &= hits - 1) {
        const int64_t p = p_first + (__ffs(hits) - 1);
        const uint32_t i = atomicAdd(q_cnt, 1u);
        if (i < (uint32_t)kKfQueue) q[i] = (uint32_t)(p - tile0) | ((uint32_t)c << 16);
        else kf_resolve(a, c, p, sm, base, tile);                      // queue full: resolve on the spot
    }
}

__device__ __forceinline__ void fence_proxy_async_smem()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// Persistent, one CTA per SM, and every WARP is its own pipeline: it walks the warp tiles
// w, w + W, w + 2W, ... (W = warps in the grid), staging tile + halo with its own TMA bulk copy into
// its own pair of buffers (the next tile is in flight while the current one is scanned) and never
// meets a CTA-wide barrier after start-up -- a slow path taken by one warp delays nobody else.
// Dynamic shared memory: [first-level bitmap kKfL0Bytes][warp 0: buffer 0, buffer 1][warp 1: ...][hit queues]
__global__ void __launch_bounds__(kKfMaxWarps * 32, 1) kfilter_scan_kernel(const __grid_constant__ KfScanArgs a)
{
    extern __shared__ __align__(128) unsigned char kf_smem[];
    __shared__ uint64_t bars[2 * kKfMaxWarps + 1];             // [2w], [2w+1]: buffers of warp w; last: bitmap
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_warps = blockDim.x >> 5;
    const int halo = a.f.halo;
    const uint32_t buf_bytes = (uint32_t)halo + kKfTile;
    const uint32_t* l0 = reinterpret_cast<const uint32_t*>(kf_smem);
    unsigned char* my_bufs = kf_smem + kKfL0Bytes + (size_t)warp * 2 * buf_bytes;
    uint64_t* my_bars = bars + 2 * warp;
    __shared__ uint32_t q_cnts[kKfMaxWarps];
    uint32_t* q = reinterpret_cast<uint32_t*>(kf_smem + kKfL0Bytes + (size_t)n_warps * 2 * buf_bytes) + warp * kKfQueue;
    uint32_t* q_cnt = &q_cnts[warp];
    if (lane == 0) *q_cnt = 0;
    uint64_t* l0_bar = bars + 2 * kKfMaxWarps;
    const int64_t n16 = (a.n_bases + 15) & ~(int64_t)15;

    auto stage = [&](int t, int b) {                            // lane 0: tile t -> buffer b
        const int64_t tile0 = (int64_t)t * kKfTile;
        const int64_t base = tile0 - halo;
        const int64_t lo = base < 0 ? 0 : base;
        const int64_t hi = tile0 + kKfTile < n16 ? tile0 + kKfTile : n16;
        const uint32_t bytes = (uint32_t)(hi - lo);
        mbar_expect_tx(&my_bars[b], bytes);
        tma_bulk_g2s(my_bufs + (size_t)b * buf_bytes + (lo - base), a.seqs + lo, bytes, &my_bars[b]);
    };
    if (tid == 0) {
        for (int i = 0; i < 2 * kKfMaxWarps + 1; ++i) mbar_init(&bars[i], 1);
        mbar_expect_tx(l0_bar, (uint32_t)kKfL0Bytes);
        for (uint32_t o = 0; o < (uint32_t)kKfL0Bytes; o += 32768u)
            tma_bulk_g2s(kf_smem + o, reinterpret_cast<const unsigned char*>(a.f.l0) + o, 32768u, l0_bar);
    }
    __syncthreads();                                            // barriers initialised
    const int stride = gridDim.x * n_warps;
    int tile = blockIdx.x * n_warps + warp;
    if (lane == 0 && tile < a.n_tiles) stage(tile, 0);
    mbar_wait(l0_bar, 0);
    const uint32_t* __restrict__ bloom = a.f.bloom;
    const uint32_t bshift = a.f.bloom_shift;

    for (int it = 0; tile < a.n_tiles; tile += stride, ++it) {
        const int b = it & 1;
        if (lane == 0 && tile + stride < a.n_tiles) {
            fence_proxy_async_smem();                           // buffer b^1 was last touched by generic loads/stores
            stage(tile + stride, b ^ 1);
        }
        mbar_wait(&my_bars[b], (it >> 1) & 1);
        const int64_t tile0 = (int64_t)tile * kKfTile;
        const int64_t base = tile0 - halo;                      // text position of buf[0]
        const int64_t lo = base < 0 ? 0 : base;
        const int64_t hi = tile0 + kKfTile < n16 ? tile0 + kKfTile : n16;
        const uint32_t bytes = (uint32_t)(hi - lo);
        unsigned char* buf = my_bufs + (size_t)b * buf_bytes;
        {   // ASCII -> symbol codes, in place; consecutive lanes take consecutive 16-byte chunks
            uint4* v = reinterpret_cast<uint4*>(buf + (lo - base));
            for (uint32_t i = lane; i < bytes / 16; i += 32) {
                uint4 w = v[i];
                w.x = kf_codes4(w.x); w.y = kf_codes4(w.y); w.z = kf_codes4(w.z); w.w = kf_codes4(w.w);
                v[i] = w;
            }
        }
        __syncwarp();
        const int64_t p0 = tile0 + (int64_t)lane * kKfPerThread;    // first end position of this lane
        if (p0 < a.n_bases) {
            const unsigned char* __restrict__ mine = buf + halo + lane * kKfPerThread;
            for (int c = 0; c < a.f.n_classes; ++c) {
                const int k = a.f.cls[c].k;
                const unsigned long long salt = a.f.cls[c].salt;
                int back = k - 1;                                   // symbols before p0 that enter the first k-mer
                if (back > p0) back = (int)p0;
                if (a.f.cls[c].exact) {
                    const unsigned long long mask = a.f.cls[c].key_mask;
                    unsigned long long key = 0;
                    for (int j = -back; j < 0; ++j) key = (key << 3) | mine[j];
#pragma unroll 1
                    for (int ch = 0; ch < kKfPerThread / 16; ++ch) {
                        const uint4 w4 = *reinterpret_cast<const uint4*>(mine + ch * 16);
                        const uint32_t w[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
                        for (int g = 0; g < 16 / kKfGroup; ++g) {
                            uint32_t as[kKfGroup], bs[kKfGroup];
                            uint32_t pass = 0;
#pragma unroll
                            for (int j = 0; j < kKfGroup; ++j) {      // level 0: shared-memory bitmap
                                const int bb = g * kKfGroup + j;
                                const uint32_t code = (w[bb >> 2] >> ((bb & 3) * 8)) & 0xffu;
                                key = ((key << 3) | code) & mask;
                                bs[j] = kf_bloom_bit(kf_hash(key, salt, as[j]));
                                const uint32_t bit = as[j] & (kKfL0Bits - 1);
                                pass |= ((l0[bit >> 5] >> (bit & 31)) & 1u) << j;
                            }
                            if (pass) {                               // level 1: one-bit Bloom filter in L2
                                uint32_t hits = 0;
#pragma unroll
                                for (int j = 0; j < kKfGroup; ++j) {
                                    const uint32_t word = (pass >> j & 1u) ? __ldg(bloom + (as[j] >> bshift)) : 0u;
                                    hits |= ((word >> bs[j]) & 1u) << j;
                                }
                                if (hits) kf_enqueue(a, c, hits, p0 + ch * 16 + g * kKfGroup, tile0, q, q_cnt, buf, base, tile);
                            }
                        }
                    }
                } else {
                    const unsigned long long bk = a.f.cls[c].bk;
                    unsigned long long key = 0;
                    int fed = 0;
                    for (int j = -back; j < 0; ++j, ++fed) key = key * kKfBase + (mine[j] + 1u);
#pragma unroll 1
                    for (int j = 0; j < kKfPerThread; ++j) {
                        key = key * kKfBase + (mine[j] + 1u);
                        if (fed == k) key -= (mine[j - k] + 1u) * bk; else ++fed;
                        uint32_t aa;
                        const uint32_t bb = kf_bloom_bit(kf_hash(key, salt, aa));
                        const uint32_t bit = aa & (kKfL0Bits - 1);
                        if ((l0[bit >> 5] >> (bit & 31)) & 1u)
                            if ((__ldg(bloom + (aa >> bshift)) >> bb) & 1u) kf_enqueue(a, c, 1u, p0 + j, tile0, q, q_cnt, buf, base, tile);
                    }
                }
            }
        }
        __syncwarp();
        {   // resolve the queued hits of this tile, one per lane
            const uint32_t n = min(*q_cnt, (uint32_t)kKfQueue);
            for (uint32_t i = lane; i < n; i += 32) {
                const uint32_t e = q[i];
                kf_resolve(a, (int)(e >> 16), tile0 + (int64_t)(e & 0xffffu), buf, base, tile);
            }
            __syncwarp();
            if (lane == 0) *q_cnt = 0;
            __syncwarp();                                       // every lane is done with buffer b
        }
    }
}

// tile_first[t] = the read that owns text position t * kKfTile (the last r with seq_off[r] <= p);
// tile_first[n_tiles] = n_reads - 1
__global__ void __launch_bounds__(256) kfilter_tile_index_kernel(const int64_t* __restrict__ seq_off, int n_reads,
                                                                 int64_t n_tiles, int32_t* __restrict__ tile_first)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t > n_tiles) return;
    int lo = 0, hi = n_reads - 1;
    if (t < n_tiles) {
        const int64_t p = t * kKfTile;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (seq_off[mid] <= p) lo = mid; else hi = mid - 1;
        }
    } else {
        lo = hi;
    }
    tile_first[t] = lo;
}

struct KfCompactArgs {
    const unsigned long long* cnt_keys;
    const uint32_t* cnt_vals;
    unsigned long long cnt_size;
    int32_t min_matches;
    int32_t* hit_read;
    int32_t* hit_locus;
    int32_t* hit_count;
    long long hit_cap;
    unsigned long long* n_hits;
};

__global__ void __launch_bounds__(256) kfilter_compact_kernel(const KfCompactArgs a)
{
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.cnt_size) return;
    const unsigned long long key = a.cnt_keys[i];
    if (key == ~0ull) return;
    const uint32_t c = a.cnt_vals[i];
    if ((int32_t)c < a.min_matches) return;
    const unsigned long long o = atomicAdd(a.n_hits, 1ull);
    if ((long long)o < a.hit_cap) {
        a.hit_read[o] = (int32_t)(key >> 32);
        a.hit_locus[o] = (int32_t)(key & 0xffffffffu);
        a.hit_count[o] = (int32_t)c;
    }
}

}  // namespace
