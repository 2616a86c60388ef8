// K2+K3 (partitioned): canonical k-mer counting with every random access on-chip.
//
// Measured on B200 (tools/atomics_bench.cu): a dependent "load slot, then atomic on it" sustains only
// ~1.5e10/s when the table lives in HBM (the v1 kernel, spk_count.cu, sits exactly there) and ~6e10/s
// even when the table is L2-resident, while shared-memory atomics run > 3e11/s.  So the chromosome is
// split into P = 2^pbits hash partitions small enough that one partition's table fits in the shared
// memory of a CTA; HBM only sees streaming traffic:
//
//   k_hist1         one traversal (TMA-staged tiles): bucket histogram (top b1 bits) privatised in smem
//   k_scatter_l1    4-tile super-tiles -> 2^b1 buckets: smem histogram, scan, one global reservation per
//                   bucket, counting-sort in smem, coalesced runs; REDs the P-bin partition histogram
//   k_scan_*        exclusive scan of the P sizes -> partition extents, cursors
//   k_scatter_l2    16384-entry chunks of a bucket -> its 2^b2 final partitions (same smem-staged scheme),
//                   writing the 32-bit remainder r
//   k_part_count32  one partition per CTA iteration: stream r (software-pipelined through registers),
//                   CAS+add into a 32-bit smem table, one reservation per partition, one sweep that
//                   writes the partition's dump contiguously (pindex) and clears the table
//   (inputs below 16.7 M bases or remainders wider than 32 bits: k_part_pass<hist|scatter> one-level
//    scatter with cursor atomics; remainders wider than 31 bits: k_part_count<ENT64> with 64-bit slots)
//
// DRAM traffic per k-mer: 2 x 0.375 B (sequence, read twice) + 2 x (4 B write + 4 B read) = 16.75 B.
// Partitioning uses a bijective mixer f on the 2k-bit canonical word: partition = top pbits of f(u),
// remainder r = the rest; the dump applies f^-1, so keys are exact and the result is bit-identical to
// the v1 kernel / the jellyfish semantics (only the dump ORDER differs, which is arbitrary anyway).
#include <stdlib.h>
#include <type_traits>
#include "spk_common.cuh"
#include "spk_tile.cuh"
#include "spk_mixer.cuh"

namespace {

constexpr int PC_SLOTS = 8192;          // smem table slots per CTA (64 KB as u64, 96 KB for wide keys)
constexpr int PC_TARGET = 4096;         // planned max mean entries per partition (all distinct -> load 0.5)
constexpr int PC_MAX_PBITS = 22;
constexpr int PC_THREADS = 256;
struct PcPlan {
    int k, pbits, ent64;
    uint64_t P;
    Mixer mx;
    uint64_t n_tiles;
    int two_level, b1, b2;        // two-level scatter: 2^b1 buckets x 2^b2 sub-partitions (b1 + b2 = pbits)
    int v3;                       // descriptor pipeline (k_v3_*): no histogram passes, no global atomics
    uint32_t n_st;                // v3: super-tiles (16384 bases)
    size_t off_buf, off_buf1, off_psize, off_pstart, off_cursor, off_segs, off_cur1, off_ustart, off_out, total;
    size_t off_desc, off_off2, off_chunks, off_btot, off_cb;     // v3
};

constexpr int SC_TILES = 4;                         // tiles per super-tile of the two-level scatter
constexpr int SC_CHUNK = SC_TILES * SPK_TILE_BASES; // 16384 entries staged in shared memory at a time
constexpr int SC_MAX_BINS = 2048;
constexpr int V3_THREADS = 1024;
constexpr int V3_CH = SC_CHUNK;                      // 16384: entries of a level-1 super-tile
// Level-2 chunk: 12288 entries (12 per thread).  With 2^9 sub-partitions a run then averages 24 entries, so that a
// run longer than the 32 entries one warp-wide load of the counter covers is rare (~5 %; at 16384 it is 45 % of the
// runs, and every such run costs the counter an un-prefetched dependent load).
constexpr int V3_E2 = 12;
constexpr int V3_CH2 = V3_THREADS * V3_E2;
constexpr int V3_ST1 = SC_CHUNK + 3 * 1024;          // slots of a level-1 super-tile: 16384 entries + <= 3 padding slots per bucket
static_assert(V3_THREADS * SPK_KMERS_PER_THREAD == V3_CH && SC_TILES * SPK_TILE_THREADS == V3_THREADS, "v3 geometry");
static_assert(V3_E2 % 4 == 0 && V3_CH2 <= 65535, "v3 chunk geometry");

inline size_t al256(size_t x) { return (x + 255) / 256 * 256; }

int auto_pbits(uint64_t n_bases, int k) {
    int pbits = 2;
    while (pbits < PC_MAX_PBITS && (n_bases >> pbits) > (uint64_t)PC_TARGET) pbits++;
    if (pbits > 2 * k) pbits = 2 * k;  // tiny k: at most 4^k distinct words
    return pbits;
}

// pbits_req > 0: the caller fixes the number of partition bits (>= the automatic choice, so that partitions
// only get smaller) — every chromosome of a genome is then split by the same function of the k-mer.
int make_plan(uint64_t n_bases, int k, int pbits_req, PcPlan* pl) {
    if (k < 1 || k > 32) return SPK_EINVAL;
    if (n_bases >= 0xffffffffull) return SPK_EINVAL;  // 32-bit partition offsets
    pl->k = k;
    int pbits = auto_pbits(n_bases, k);
    if (pbits_req > 0) {
        if (pbits_req < pbits || pbits_req > PC_MAX_PBITS || pbits_req > 2 * k) return SPK_EINVAL;
        pbits = pbits_req;
    }
    pl->pbits = pbits;
    pl->P = 1ull << pbits;
    pl->ent64 = (2 * k - pbits > 32) ? 1 : 0;
    pl->mx.mask = (k == 32) ? ~0ull : ((1ull << (2 * k)) - 1);
    pl->mx.s = k;
    pl->mx.rbits = 2 * k - pbits;
    pl->n_tiles = (n_bases + SPK_TILE_BASES - 1) / SPK_TILE_BASES;
    // two-level scatter when it pays (enough partitions) and the level-1 entry still fits 32 bits
    pl->b1 = (pbits + 1) / 2;
    pl->b2 = pbits - pl->b1;
    const char* e2 = getenv("SPK_PCOUNT_TWO_LEVEL");
    pl->two_level = (!pl->ent64 && pbits >= 12 && 2 * k - pl->b1 <= 32 && (1 << pl->b1) <= SC_MAX_BINS &&
                     !(e2 && e2[0] == '0')) ? 1 : 0;
    // v3 descriptor pipeline: 2^9 sub-partitions per bucket (runs of ~32 entries per chunk for the counter),
    // 2^(pbits-9) buckets (runs of >= 16 entries per super-tile for level 2), 32-bit entries and offsets
    const char* e3 = getenv("SPK_PCOUNT_PIPE");
    pl->v3 = 0;
    pl->n_st = (uint32_t)((pl->n_tiles + SC_TILES - 1) / SC_TILES);
    if (!pl->ent64 && pbits >= 15 && pbits <= 19 && 2 * k - pbits <= 31 && 2 * k - (pbits - 9) <= 31 &&
        n_bases < 3400000000ull && !(e3 && e3[0] == 'v' && e3[1] == '2')) {
        pl->v3 = 1;
        pl->two_level = 0;
        pl->b2 = 9;
        pl->b1 = pbits - 9;
    }
    size_t off = 0;
    if (pl->v3) {
        const size_t nb1 = (size_t)1 << pl->b1, nb2 = (size_t)1 << pl->b2;
        const size_t max_chunks = ((size_t)pl->n_st * V3_ST1 + V3_CH2 - 1) / V3_CH2 + nb1 + 1;   // (slots: entries + padding)
        pl->off_buf = off;                                   // chunk-sorted stream (level-2 output)
        off += al256(max_chunks * V3_CH2 * 4);
        pl->off_buf1 = off;                                  // super-tile-sorted stream (level-1 output)
        off += al256(((size_t)pl->n_st + 1) * V3_ST1 * 4);
        pl->off_desc = off;
        off += al256(nb1 * (size_t)pl->n_st * 4);
        pl->off_off2 = off;
        off += al256(max_chunks * (nb2 + 1) * 2);
        pl->off_chunks = off;
        off += al256(max_chunks * 16);
        pl->off_btot = off;
        off += al256(nb1 * 4);
        pl->off_cb = off;
        off += al256((nb1 + 1) * 4);
        pl->off_psize = pl->off_pstart = pl->off_cursor = pl->off_segs = pl->off_cur1 = pl->off_ustart = 0;
        pl->off_out = off;
        off += 256;
        pl->total = off;
        return SPK_OK;
    }
    pl->off_buf = off;
    off += al256((size_t)(n_bases + 64) * (pl->ent64 ? 8 : 4));
    pl->off_buf1 = off;
    if (pl->two_level) off += al256((size_t)(n_bases + 64) * 4);
    pl->off_cur1 = off;
    off += al256((size_t)(SC_MAX_BINS + 1) * 4);
    pl->off_ustart = off;
    off += al256((size_t)(SC_MAX_BINS + 2) * 4);
    pl->off_psize = off;
    off += al256((size_t)(pl->P + 1) * 4);
    pl->off_pstart = off;
    off += al256((size_t)(pl->P + 1) * 4);
    pl->off_cursor = off;
    off += al256((size_t)(pl->P + 1) * 4);
    pl->off_segs = off;
    off += al256((size_t)(pl->P / 1024 + 2) * 4);
    pl->off_out = off;
    off += 256;
    pl->total = off;
    return SPK_OK;
}

// ---- phases 0 and 1: one traversal of the packed sequence -----------------------------------------------
template <bool SCATTER, bool ENT64>
__global__ void __launch_bounds__(SPK_TILE_THREADS, 4)
k_part_pass(const uint8_t* __restrict__ packed, const uint8_t* __restrict__ valid, uint64_t n_tiles, int k,
            Mixer mx, uint32_t* __restrict__ psize, uint32_t* __restrict__ cursor, void* __restrict__ buf,
            uint64_t* __restrict__ stats) {
    __shared__ SpkTileSmem sm;
    const int tid = threadIdx.x;
    const SpkKmerParams kp = spk_kmer_params(k);
    spk_tile_init(sm);
    uint64_t n_valid = 0;
    uint64_t tile = blockIdx.x;
    if (tid == 0 && tile < n_tiles) spk_tile_issue(sm, packed, valid, tile, 0);
    const uint64_t rmask = (mx.rbits >= 64) ? ~0ull : ((1ull << mx.rbits) - 1);
    for (uint32_t it = 0; tile < n_tiles; it++, tile += gridDim.x) {
        const int b = it & 1;
        __syncthreads();
        if (tid == 0 && tile + gridDim.x < n_tiles) spk_tile_issue(sm, packed, valid, tile + gridDim.x, b ^ 1);
        spk_mbar_wait(&sm.bar[b], (it >> 1) & 1);
        uint64_t key[SPK_KMERS_PER_THREAD];
        uint32_t okmask;
        spk_tile_kmers(sm, b, kp, key, okmask);
        n_valid += __popc(okmask);
        if (!SCATTER) {
#pragma unroll
            for (int j = 0; j < SPK_KMERS_PER_THREAD; j++)
                if ((okmask >> j) & 1u) atomicAdd(&psize[mx.fwd(key[j]) >> mx.rbits], 1u);
        } else {
            // issue all 16 cursor atomics first, then the 16 stores
            uint32_t pos[SPK_KMERS_PER_THREAD];
#pragma unroll
            for (int j = 0; j < SPK_KMERS_PER_THREAD; j++) {
                key[j] = mx.fwd(key[j]);
                pos[j] = ((okmask >> j) & 1u) ? atomicAdd(&cursor[key[j] >> mx.rbits], 1u) : 0u;
            }
#pragma unroll
            for (int j = 0; j < SPK_KMERS_PER_THREAD; j++) {
                if ((okmask >> j) & 1u) {
                    const uint64_t r = key[j] & rmask;
                    if (ENT64) ((uint64_t*)buf)[pos[j]] = r;
                    else ((uint32_t*)buf)[pos[j]] = (uint32_t)r;
                }
            }
        }
    }
    if (!SCATTER) {
        n_valid = spk_warp_sum_u64(n_valid);
        if ((tid & 31) == 0 && n_valid) atomicAdd((unsigned long long*)&stats[0], (unsigned long long)n_valid);
    }
}

// ---- exclusive scan of the P partition sizes (three passes over 1024-element segments) --------------------
__device__ __forceinline__ uint32_t block_scan_1024(uint32_t v, uint32_t* s_warp, uint32_t* total) {
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if ((threadIdx.x & 31) >= o) incl += t;
    }
    if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = incl;
    __syncthreads();
    uint32_t prefix = 0, tot = 0;
    for (int w = 0; w < 32; w++) {
        if (w < (int)(threadIdx.x >> 5)) prefix += s_warp[w];
        tot += s_warp[w];
    }
    __syncthreads();
    *total = tot;
    return prefix + incl - v;
}

__global__ void __launch_bounds__(1024) k_scan_seg_totals(const uint32_t* __restrict__ v, uint64_t n,
                                                           uint32_t* __restrict__ seg) {
    __shared__ uint32_t s_warp[32];
    const uint64_t i = (uint64_t)blockIdx.x * 1024 + threadIdx.x;
    uint32_t tot;
    block_scan_1024(i < n ? v[i] : 0u, s_warp, &tot);
    if (threadIdx.x == 0) seg[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(1024) k_scan_segs(uint32_t* seg, uint64_t nseg) {
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (uint64_t base = 0; base < nseg; base += 1024) {
        const uint64_t i = base + threadIdx.x;
        uint32_t tot;
        const uint32_t excl = block_scan_1024(i < nseg ? seg[i] : 0u, s_warp, &tot);
        if (i < nseg) seg[i] = s_carry + excl;
        __syncthreads();
        if (threadIdx.x == 0) s_carry += tot;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(1024) k_scan_apply(const uint32_t* __restrict__ v, uint64_t n,
                                                      const uint32_t* __restrict__ seg,
                                                      uint32_t* __restrict__ pstart,
                                                      uint32_t* __restrict__ cursor) {
    __shared__ uint32_t s_warp[32];
    const uint64_t i = (uint64_t)blockIdx.x * 1024 + threadIdx.x;
    const uint32_t x = i < n ? v[i] : 0u;
    uint32_t tot;
    const uint32_t e = seg[blockIdx.x] + block_scan_1024(x, s_warp, &tot);
    if (i < n) {
        pstart[i] = e;
        cursor[i] = e;
        if (i == n - 1) pstart[n] = e + x;
    }
}

// ---- two-level scatter: shared-memory staged, coalesced runs, no per-k-mer global atomic --------------------
// Level 1 splits the k-mers of a 4-tile super-tile into 2^b1 buckets, level 2 splits 16384-entry chunks of
// a bucket into its 2^b2 final partitions.  Both levels: histogram in smem, exclusive scan, ONE global
// atomicAdd per non-empty bin to reserve a run, counting-sort placement in smem, coalesced write of runs.
struct ScatterSmem {
    uint32_t* cnt;     // [nb]   histogram, then running cursor
    uint32_t* off;     // [nb]   exclusive prefix inside the chunk
    uint32_t* gbase;   // [nb]   (global run start) - off  (mod 2^32)
    uint32_t* sorted;  // [SC_CHUNK]
    uint16_t* bin;     // [SC_CHUNK]
};

__device__ __forceinline__ ScatterSmem carve_scatter(uint8_t* base, int nb) {
    ScatterSmem s;
    s.cnt = (uint32_t*)base;
    s.off = s.cnt + nb;
    s.gbase = s.off + nb;
    s.sorted = s.gbase + nb;
    s.bin = (uint16_t*)(s.sorted + SC_CHUNK);
    return s;
}
inline size_t scatter_smem_bytes(int nb) { return (size_t)nb * 12 + (size_t)SC_CHUNK * 6; }

// exclusive scan of cnt[nb] into off[nb] by a 256-thread CTA (nb a multiple of 256 or smaller); returns total
template <int T = SPK_TILE_THREADS>
__device__ __forceinline__ uint32_t bins_scan(const uint32_t* cnt, uint32_t* off, int nb, uint32_t* s_warp) {
    const int per = (nb + T - 1) / T;
    const int b0 = threadIdx.x * per;
    uint32_t sum = 0;
    for (int q = 0; q < per; q++)
        if (b0 + q < nb) sum += cnt[b0 + q];
    uint32_t incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if ((threadIdx.x & 31) >= o) incl += t;
    }
    if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = incl;
    __syncthreads();
    uint32_t prefix = 0, total = 0;
    for (int w = 0; w < T / 32; w++) {
        if (w < (int)(threadIdx.x >> 5)) prefix += s_warp[w];
        total += s_warp[w];
    }
    uint32_t run = prefix + incl - sum;
    for (int q = 0; q < per; q++)
        if (b0 + q < nb) {
            off[b0 + q] = run;
            run += cnt[b0 + q];
        }
    __syncthreads();
    return total;
}

// reserve one global run per non-empty bin and turn cnt into the running cursor
template <int T = SPK_TILE_THREADS>
__device__ __forceinline__ void bins_reserve(ScatterSmem& s, int nb, uint32_t* __restrict__ gcursor) {
    for (int b = threadIdx.x; b < nb; b += T) {
        const uint32_t c = s.cnt[b];
        if (c) s.gbase[b] = atomicAdd(&gcursor[b], c) - s.off[b];
        s.cnt[b] = s.off[b];
    }
    __syncthreads();
}

// Bucket-level histogram (2^b1 bins, privatised in shared memory) — all the two-level scatter needs before
// level 1 can reserve its runs.  The P-bin histogram of the final partitions is accumulated by level 1
// itself (its REDs to L2 hide under that kernel's shared-memory work), so the separate full-histogram
// traversal (L2-atomic bound, as long as level 1 itself) is gone.
__global__ void __launch_bounds__(SPK_TILE_THREADS, 4)
k_hist1(const uint8_t* __restrict__ packed, const uint8_t* __restrict__ valid, uint64_t n_tiles, int k, Mixer mx,
        int b1, uint32_t* __restrict__ bsize, uint64_t* __restrict__ stats, uint32_t* __restrict__ psize,
        int red_split) {
    __shared__ SpkTileSmem sm;
    __shared__ uint32_t s_bins[SC_MAX_BINS];
    const int tid = threadIdx.x;
    const int nb = 1 << b1;
    const int sh1 = 2 * k - b1;
    const SpkKmerParams kp = spk_kmer_params(k);
    for (int b = tid; b < nb; b += SPK_TILE_THREADS) s_bins[b] = 0;
    spk_tile_init(sm);
    uint64_t n_valid = 0;
    uint64_t tile = blockIdx.x;
    if (tid == 0 && tile < n_tiles) spk_tile_issue(sm, packed, valid, tile, 0);
    for (uint32_t it = 0; tile < n_tiles; it++, tile += gridDim.x) {
        const int b = it & 1;
        __syncthreads();
        if (tid == 0 && tile + gridDim.x < n_tiles) spk_tile_issue(sm, packed, valid, tile + gridDim.x, b ^ 1);
        spk_mbar_wait(&sm.bar[b], (it >> 1) & 1);
        uint64_t key[SPK_KMERS_PER_THREAD];
        uint32_t okmask;
        spk_tile_kmers(sm, b, kp, key, okmask);
        n_valid += __popc(okmask);
#pragma unroll
        for (int j = 0; j < SPK_KMERS_PER_THREAD; j++)
            if ((okmask >> j) & 1u) {
                const uint64_t h = mx.fwd(key[j]);
                atomicAdd(&s_bins[h >> sh1], 1u);
                // this kernel is ALU-bound with the L2 atomic units idle, level 1 is co-limited by them: the
                // P-bin partition histogram is split between the two (k-mers j < red_split of every thread here)
                if (j < red_split) atomicAdd(&psize[h >> mx.rbits], 1u);
            }
    }
    __syncthreads();
    for (int b = tid; b < nb; b += SPK_TILE_THREADS)
        if (s_bins[b]) atomicAdd(&bsize[b], s_bins[b]);
    n_valid = spk_warp_sum_u64(n_valid);
    if ((tid & 31) == 0 && n_valid) atomicAdd((unsigned long long*)&stats[0], (unsigned long long)n_valid);
}

// exclusive scan of the <= SC_MAX_BINS bucket sizes -> level-1 run cursors (single CTA)
__global__ void __launch_bounds__(1024) k_bucket_scan(const uint32_t* __restrict__ bsize, int nb,
                                                       uint32_t* __restrict__ cur1) {
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < nb; base += 1024) {
        const int b = base + threadIdx.x;
        uint32_t tot;
        const uint32_t excl = block_scan_1024(b < nb ? bsize[b] : 0u, s_warp, &tot);
        if (b < nb) cur1[b] = s_carry + excl;
        __syncthreads();
        if (threadIdx.x == 0) s_carry += tot;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(SPK_TILE_THREADS, 2)
k_scatter_l1(const uint8_t* __restrict__ packed, const uint8_t* __restrict__ valid, uint64_t n_tiles, int k,
             Mixer mx, int b1, uint32_t* __restrict__ cur1, uint32_t* __restrict__ buf1,
             uint32_t* __restrict__ psize, int red_split) {
    extern __shared__ __align__(16) uint8_t s_raw[];
    constexpr int PKW = (SC_TILES * SPK_TILE_PACKED_BYTES + SPK_HALO_PACKED_BYTES) / 4;   // 1028
    constexpr int VDW = (SC_TILES * SPK_TILE_VALID_BYTES + SPK_HALO_VALID_BYTES) / 4;     // 516
    uint32_t* pk = (uint32_t*)s_raw;
    uint32_t* vd = pk + PKW;
    uint64_t* bar = (uint64_t*)(vd + VDW + (VDW & 1));
    const int nb = 1 << b1;
    ScatterSmem sc = carve_scatter((uint8_t*)(bar + 2), nb);
    __shared__ uint32_t s_warp[SPK_TILE_THREADS / 32];
    const int tid = threadIdx.x;
    const SpkKmerParams kp = spk_kmer_params(k);
    const int sh1 = 2 * k - b1;                                   // bucket = h >> sh1
    const uint64_t m1 = (sh1 >= 64) ? ~0ull : ((1ull << sh1) - 1); // level-1 entry = h & m1  (<= 32 bits)
    if (tid == 0) {
        spk_mbar_init(bar, 1);
        spk_fence_mbar_init();
    }
    __syncthreads();
    const uint64_t n_super = (n_tiles + SC_TILES - 1) / SC_TILES;
    uint32_t it = 0;
    for (uint64_t st = blockIdx.x; st < n_super; st += gridDim.x, it++) {
        const uint64_t t0 = st * SC_TILES;
        const int ntl = (int)min((uint64_t)SC_TILES, n_tiles - t0);
        if (tid == 0) {
            const uint32_t pb = ntl * SPK_TILE_PACKED_BYTES + SPK_HALO_PACKED_BYTES;
            const uint32_t vb = ntl * SPK_TILE_VALID_BYTES + SPK_HALO_VALID_BYTES;
            spk_mbar_expect_tx(bar, pb + vb);
            spk_bulk_g2s(pk, packed + t0 * SPK_TILE_PACKED_BYTES, pb, bar);
            spk_bulk_g2s(vd, valid + t0 * SPK_TILE_VALID_BYTES, vb, bar);
        }
        for (int b = tid; b < nb; b += SPK_TILE_THREADS) sc.cnt[b] = 0;
        __syncthreads();
        spk_mbar_wait(bar, it & 1);
        // pass A: histogram of bucket ids
        for (int t = 0; t < ntl; t++) {
            uint64_t key[SPK_KMERS_PER_THREAD];
            uint32_t okmask;
            spk_kmers_from(pk + t * (SPK_TILE_PACKED_BYTES / 4), vd + t * (SPK_TILE_VALID_BYTES / 4), kp, key, okmask);
#pragma unroll
            for (int j = 0; j < SPK_KMERS_PER_THREAD; j++)
                if ((okmask >> j) & 1u) atomicAdd(&sc.cnt[mx.fwd(key[j]) >> sh1], 1u);
        }
        __syncthreads();
        const uint32_t n_e = bins_scan(sc.cnt, sc.off, nb, s_warp);
        bins_reserve(sc, nb, cur1);
        // pass B: counting-sort placement
        for (int t = 0; t < ntl; t++) {
            uint64_t key[SPK_KMERS_PER_THREAD];
            uint32_t okmask;
            spk_kmers_from(pk + t * (SPK_TILE_PACKED_BYTES / 4), vd + t * (SPK_TILE_VALID_BYTES / 4), kp, key, okmask);
#pragma unroll
            for (int j = 0; j < SPK_KMERS_PER_THREAD; j++)
                if ((okmask >> j) & 1u) {
                    const uint64_t h = mx.fwd(key[j]);
                    const uint32_t b = (uint32_t)(h >> sh1);
                    if (j >= red_split) atomicAdd(&psize[h >> mx.rbits], 1u);   // final-partition histogram (RED to L2)
                    const uint32_t p = atomicAdd(&sc.cnt[b], 1u);
                    sc.sorted[p] = (uint32_t)(h & m1);
                    sc.bin[p] = (uint16_t)b;
                }
        }
        __syncthreads();
        for (uint32_t i = tid; i < n_e; i += SPK_TILE_THREADS) buf1[sc.gbase[sc.bin[i]] + i] = sc.sorted[i];
        __syncthreads();   // smem is reused by the next super-tile (and by the next bulk copy)
    }
}

// cursors of the level-1 buckets and the unit table of level 2 (unit = one 16384-entry chunk of a bucket)
__global__ void __launch_bounds__(1024)
k_scatter_prepare(const uint32_t* __restrict__ pstart, int b1, int b2, uint32_t* __restrict__ cur1,
                  uint32_t* __restrict__ ustart) {
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_carry;
    const int nb = 1 << b1;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < nb; base += 1024) {
        const int b = base + threadIdx.x;
        uint32_t units = 0;
        if (b < nb) {
            const uint32_t bs = pstart[(uint64_t)b << b2], be = pstart[((uint64_t)b + 1) << b2];
            cur1[b] = bs;
            units = (be - bs + SC_CHUNK - 1) / SC_CHUNK;
        }
        uint32_t tot;
        const uint32_t excl = block_scan_1024(units, s_warp, &tot);
        if (b < nb) ustart[b] = s_carry + excl;
        __syncthreads();
        if (threadIdx.x == 0) s_carry += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) ustart[nb] = s_carry;
}

constexpr int SC2_THREADS = 512;     // level 2 is latency-bound on its chunk loads: twice the warps of level 1

__global__ void __launch_bounds__(SC2_THREADS, 2)
k_scatter_l2(const uint32_t* __restrict__ buf1, const uint32_t* __restrict__ pstart,
             const uint32_t* __restrict__ ustart, int b1, int b2, int rbits, uint32_t* __restrict__ cursor,
             uint32_t* __restrict__ buf) {
    extern __shared__ __align__(16) uint8_t s_raw[];
    const int nb = 1 << b2;
    ScatterSmem sc = carve_scatter(s_raw, nb);
    __shared__ uint32_t s_warp[SC2_THREADS / 32];
    __shared__ uint32_t s_unit[3];   // bucket, begin, end
    const int tid = threadIdx.x;
    const int nb1 = 1 << b1;
    const uint32_t n_units = ustart[nb1];
    const uint32_t rmask = (rbits >= 32) ? 0xffffffffu : ((1u << rbits) - 1u);
    for (uint32_t u = blockIdx.x; u < n_units; u += gridDim.x) {
        if (tid == 0) {
            int lo = 0, hi = nb1;              // last bucket with ustart[b] <= u
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (ustart[mid] <= u) lo = mid;
                else hi = mid;
            }
            const uint32_t bs = pstart[(uint64_t)lo << b2], be = pstart[((uint64_t)lo + 1) << b2];
            const uint32_t beg = bs + (u - ustart[lo]) * SC_CHUNK;
            s_unit[0] = lo;
            s_unit[1] = beg;
            s_unit[2] = min(beg + (uint32_t)SC_CHUNK, be);
        }
        for (int b = tid; b < nb; b += SC2_THREADS) sc.cnt[b] = 0;
        __syncthreads();
        const uint32_t bucket = s_unit[0], beg = s_unit[1], end = s_unit[2];
        for (uint32_t base = beg; base < end; base += 4 * SC2_THREADS) {
            uint32_t v[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const uint32_t i = base + u * SC2_THREADS + tid;
                v[u] = i < end ? __ldg(buf1 + i) : 0u;
            }
#pragma unroll
            for (int u = 0; u < 4; u++)
                if (base + u * SC2_THREADS + tid < end) atomicAdd(&sc.cnt[v[u] >> rbits], 1u);
        }
        __syncthreads();
        const uint32_t n_e = bins_scan<SC2_THREADS>(sc.cnt, sc.off, nb, s_warp);
        bins_reserve<SC2_THREADS>(sc, nb, cursor + ((uint64_t)bucket << b2));
        for (uint32_t base = beg; base < end; base += 4 * SC2_THREADS) {
            uint32_t v[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const uint32_t i = base + u * SC2_THREADS + tid;
                v[u] = i < end ? __ldg(buf1 + i) : 0u;     // second read of the chunk: L1/L2 hit
            }
#pragma unroll
            for (int u = 0; u < 4; u++)
                if (base + u * SC2_THREADS + tid < end) {
                    const uint32_t sub = v[u] >> rbits;
                    const uint32_t p = atomicAdd(&sc.cnt[sub], 1u);
                    sc.sorted[p] = v[u] & rmask;
                    sc.bin[p] = (uint16_t)sub;
                }
        }
        __syncthreads();
        for (uint32_t i = tid; i < n_e; i += SC2_THREADS) buf[sc.gbase[sc.bin[i]] + i] = sc.sorted[i];
        __syncthreads();
    }
}


// ---- v3: descriptor pipeline -----------------------------------------------------------------------------------
// The two-level scatter above needs the bucket sizes before level 1 can write (k_hist1: a whole extra traversal)
// and the P partition sizes before level 2 can write (one RED to L2 per k-mer, the hidden limiter of level 1),
// and both levels generate every k-mer twice (histogram pass, placement pass).  v3 removes all three:
//   k_v3_l1      one traversal, 1024 threads x 16 k-mers held in registers: rank = smem atomicAdd on the bucket
//                counter, block scan, placement into a smem buffer, ONE bulk (TMA) store of the locally sorted
//                super-tile at a fixed stride.  Every bucket's run starts on a 16-byte boundary and is padded to
//                a multiple of 4 entries with a sentinel, so that level 2 can copy it with 16-byte cp.async.  The
//                per-(bucket, super-tile) run descriptors (start | slots) go to descT[bucket][super-tile].
//                No global reservation, no histogram pass, no recompute.
//   k_v3_plan / k_v3_chunks   bucket totals -> chunk table: bucket b's virtual stream (its runs in super-tile
//                order) is cut into 12288-slot chunks; every chunk records the run it starts in.
//   k_v3_l2      one chunk at a time: gather its runs (cp.async 16 B, 8 lanes per run), rank by sub-partition in
//                registers, scan, place, bulk store of the chunk sorted by sub-partition + the u16 offsets of the
//                2^b2 sub-partitions inside the chunk.  No partition histogram, no global atomics.
//   k_part_count32<true>   partition (b, sub) = one short run per chunk of bucket b, gathered (warp per run,
//                software-pipelined one partition ahead through registers) into the same smem table.
// Everything is a deterministic function of the input (no atomic decides an output position).

constexpr uint32_t V3_SENT = 0xffffffffu;            // padding slot (entries are < 2^31: see make_plan)

__device__ __forceinline__ void spk_bulk_s2g(void* gdst, const void* ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(spk_smem_u32(ssrc)),
                 "r"(bytes)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void spk_bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void spk_bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void spk_cp_async16(uint32_t sdst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sdst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void spk_cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// exclusive scan over the 1024 threads of a CTA (two barriers, ~30 instructions per thread); s_warp: 32 words
__device__ __forceinline__ uint32_t v3_scan(uint32_t v, uint32_t* s_warp, uint32_t& total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    const uint32_t wt = s_warp[lane];
    uint32_t wi = wt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, wi, o);
        if (lane >= o) wi += t;
    }
    total = __shfl_sync(0xffffffffu, wi, 31);
    const uint32_t prefix = __shfl_sync(0xffffffffu, wi - wt, warp);
    __syncthreads();
    return prefix + incl - v;
}

struct V3Chunk {          // one 12288-slot chunk of a bucket's virtual stream
    uint32_t st0;         // super-tile whose run contains the chunk's first slot
    uint32_t skip;        // slots of that run that belong to the previous chunk
    uint32_t bucket;
    uint32_t n;           // slots in the chunk (12288 except the last chunk of a bucket)
};

__global__ void __launch_bounds__(V3_THREADS, 1)
k_v3_l1(const uint8_t* __restrict__ packed, const uint8_t* __restrict__ valid, uint64_t n_tiles, int k, Mixer mx,
        int b1, uint32_t* __restrict__ buf1, uint32_t* __restrict__ descT, uint32_t n_st,
        uint32_t* __restrict__ btot, uint64_t* __restrict__ stats) {
    extern __shared__ __align__(16) uint8_t s_raw[];
    constexpr int PKW = (SC_TILES * SPK_TILE_PACKED_BYTES + SPK_HALO_PACKED_BYTES) / 4;   // 1028
    constexpr int VDW = (SC_TILES * SPK_TILE_VALID_BYTES + SPK_HALO_VALID_BYTES) / 4;     // 516
    uint32_t* sorted = (uint32_t*)s_raw;                 // [V3_ST1]
    uint32_t* pk0 = sorted + V3_ST1;                     // [2][PKW]
    uint32_t* vd0 = pk0 + 2 * PKW;                       // [2][VDW]
    uint64_t* bar = (uint64_t*)(vd0 + 2 * VDW);          // [2]   (8-byte aligned: V3_ST1, 2*PKW + 2*VDW are even)
    uint32_t* cnt = (uint32_t*)(bar + 2);                // [1024]
    uint32_t* off = cnt + V3_THREADS;                    // [1024]
    __shared__ uint32_t s_warp[32];
    const int nb = 1 << b1;                              // <= 1024: bucket q is scanned by thread q
    const int tid = threadIdx.x;
    const int tile = tid >> 8, t256 = tid & 255;
    const SpkKmerParams kp = spk_kmer_params(k);
    const int sh1 = 2 * k - b1;                                    // bucket = h >> sh1
    const uint64_t m1 = (sh1 >= 64) ? ~0ull : ((1ull << sh1) - 1);  // entry = h & m1  (<= 31 bits)
    if (tid == 0) {
        spk_mbar_init(&bar[0], 1);
        spk_mbar_init(&bar[1], 1);
        spk_fence_mbar_init();
    }
    cnt[tid] = 0;
    __syncthreads();
    auto issue = [&](uint64_t st, int buf) {
        const uint64_t t0 = st * SC_TILES;
        const int ntl = (int)min((uint64_t)SC_TILES, n_tiles - t0);
        const uint32_t pb = ntl * SPK_TILE_PACKED_BYTES + SPK_HALO_PACKED_BYTES;
        const uint32_t vb = ntl * SPK_TILE_VALID_BYTES + SPK_HALO_VALID_BYTES;
        spk_mbar_expect_tx(&bar[buf], pb + vb);
        spk_bulk_g2s(pk0 + buf * PKW, packed + t0 * SPK_TILE_PACKED_BYTES, pb, &bar[buf]);
        spk_bulk_g2s(vd0 + buf * VDW, valid + t0 * SPK_TILE_VALID_BYTES, vb, &bar[buf]);
    };
    uint64_t n_valid = 0;
    uint64_t st = blockIdx.x;
    if (tid == 0 && st < n_st) issue(st, 0);
    for (uint32_t it = 0; st < n_st; it++, st += gridDim.x) {
        const int b = it & 1;
        // (every thread passed the barrier that follows the placement of the previous super-tile, so nobody still
        //  reads the tile buffer b^1 that is refilled here)
        if (tid == 0 && st + gridDim.x < n_st) issue(st + gridDim.x, b ^ 1);
        spk_mbar_wait(&bar[b], (it >> 1) & 1);
        const int ntl = (int)min((uint64_t)SC_TILES, n_tiles - st * SC_TILES);
        uint32_t e[SPK_KMERS_PER_THREAD], meta[SPK_KMERS_PER_THREAD];
        {
            uint64_t key[SPK_KMERS_PER_THREAD];
            uint32_t okmask;
            spk_kmers_from_t(pk0 + b * PKW + tile * (SPK_TILE_PACKED_BYTES / 4),
                             vd0 + b * VDW + tile * (SPK_TILE_VALID_BYTES / 4), t256, kp, key, okmask);
            if (tile >= ntl) okmask = 0;                    // tiles past the end were not loaded
#pragma unroll
            for (int j = 0; j < SPK_KMERS_PER_THREAD; j++) {
                meta[j] = 0xffffffffu;
                e[j] = 0;
                if ((okmask >> j) & 1u) {
                    const uint64_t h = mx.fwd(key[j]);
                    const uint32_t bk = (uint32_t)(h >> sh1);
                    e[j] = (uint32_t)(h & m1);
                    meta[j] = (bk << 16) | atomicAdd(&cnt[bk], 1u);     // rank inside the bucket (< 16384)
                }
            }
        }
        __syncthreads();
        // bucket q: c entries in (c + 3) & ~3 slots starting at a multiple of 4
        const uint32_t c = cnt[tid];
        const uint32_t slots = (c + 3u) & ~3u;
        uint32_t n_slots;
        const uint32_t o = v3_scan(slots, s_warp, n_slots);
        off[tid] = o;
        cnt[tid] = 0;
        if (tid < nb) {
            descT[(size_t)tid * n_st + st] = o | (slots << 16);      // start < 2^16, slots <= 16384
            if (slots) atomicAdd(&btot[tid], slots);
        }
        if (tid == 0) {
            spk_bulk_wait_read();          // the previous super-tile's bulk store has finished reading `sorted`
        }
        n_valid += c;
        __syncthreads();
        for (uint32_t i = c; i < slots; i++) sorted[o + i] = V3_SENT;
#pragma unroll
        for (int j = 0; j < SPK_KMERS_PER_THREAD; j++)
            if (meta[j] != 0xffffffffu) sorted[off[meta[j] >> 16] + (meta[j] & 0xffffu)] = e[j];
        spk_fence_proxy_async();           // generic-proxy smem writes -> visible to the bulk-copy (async) proxy
        __syncthreads();
        if (tid == 0 && n_slots) spk_bulk_s2g(buf1 + (size_t)st * V3_ST1, sorted, n_slots * 4);
    }
    if (tid == 0) spk_bulk_wait_all();
    n_valid = spk_warp_sum_u64(n_valid);
    if ((tid & 31) == 0 && n_valid) atomicAdd((unsigned long long*)&stats[0], (unsigned long long)n_valid);
}

// bucket totals (slots) -> first chunk of every bucket (single CTA; nb <= 1024)
__global__ void __launch_bounds__(1024) k_v3_plan(const uint32_t* __restrict__ btot, int nb, uint32_t* __restrict__ cb) {
    __shared__ uint32_t s_warp[32];
    const int b = threadIdx.x;
    const uint32_t nc = b < nb ? (btot[b] + V3_CH2 - 1) / V3_CH2 : 0u;
    uint32_t tot;
    const uint32_t excl = v3_scan(nc, s_warp, tot);
    if (b < nb) cb[b] = excl;
    if (b == 0) cb[nb] = tot;
}

// one CTA per bucket: walk its run sizes in super-tile order, note where every chunk starts
__global__ void __launch_bounds__(1024)
k_v3_chunks(const uint32_t* __restrict__ descT, uint32_t n_st, const uint32_t* __restrict__ btot,
            const uint32_t* __restrict__ cb, V3Chunk* __restrict__ chunks) {
    __shared__ uint32_t s_warp[32];
    const uint32_t b = blockIdx.x;
    const uint32_t total = btot[b], g0 = cb[b];
    uint32_t carry = 0;
    for (uint32_t base = 0; base < n_st; base += 1024) {
        const uint32_t st = base + threadIdx.x;
        const uint32_t c = st < n_st ? (descT[(size_t)b * n_st + st] >> 16) : 0u;
        uint32_t tot;
        const uint32_t excl = carry + v3_scan(c, s_warp, tot);
        // chunk boundaries inside this run (a run of up to 16384 slots can hold two)
        for (uint32_t m = (excl + V3_CH2 - 1) / V3_CH2; c && (uint64_t)m * V3_CH2 < (uint64_t)excl + c; m++) {
            const uint64_t at = (uint64_t)m * V3_CH2;
            V3Chunk ch;
            ch.st0 = st;
            ch.skip = (uint32_t)(at - excl);
            ch.bucket = b;
            ch.n = (uint32_t)min((uint64_t)V3_CH2, (uint64_t)total - at);
            chunks[g0 + m] = ch;
        }
        carry += tot;
    }
}

__global__ void __launch_bounds__(V3_THREADS, 1)
k_v3_l2(const uint32_t* __restrict__ buf1, const uint32_t* __restrict__ descT, uint32_t n_st,
        const V3Chunk* __restrict__ chunks, const uint32_t* __restrict__ n_chunks_p, int b2, int rbits,
        uint32_t* __restrict__ buf2, uint16_t* __restrict__ off2) {
    extern __shared__ __align__(16) uint8_t s_raw[];
    uint32_t* sorted = (uint32_t*)s_raw;                 // [V3_CH2]
    uint32_t* stage = sorted + V3_CH2;                   // [V3_CH2]
    uint32_t* r_src = stage + V3_CH2;                    // [1024] run list of one batch: first slot in buf1 / 4
    uint32_t* r_dl = r_src + V3_THREADS;                 // [1024] dst | len << 16 (slots)
    uint32_t* cnt = r_dl + V3_THREADS;                   // [1024]
    uint32_t* off = cnt + V3_THREADS;                    // [1024]
    __shared__ uint32_t s_warp[32];
    const int nb = 1 << b2;                              // <= 1024
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t n_chunks = *n_chunks_p;
    const uint32_t rmask = (rbits >= 32) ? 0xffffffffu : ((1u << rbits) - 1u);
    const uint32_t stage_s = spk_smem_u32(stage);
    cnt[tid] = 0;
    __syncthreads();

    // gather the runs of a chunk into `stage` (asynchronous 16-byte copies; completed by cp.async.wait_all + barrier).
    // The chunk record and this thread's first run descriptor were loaded one iteration earlier: their latency —
    // two dependent global loads — would otherwise be exposed to all 1024 threads of the only CTA on the SM.
    auto gather_issue = [&](const V3Chunk ch, const uint32_t d_first) {
        const uint32_t* drow = descT + (size_t)ch.bucket * n_st;
        uint32_t acc = 0;
        for (uint32_t base = ch.st0; base < n_st && acc < ch.n; base += V3_THREADS) {     // block-uniform
            const uint32_t st = base + tid;
            uint32_t c = 0, start = 0;
            if (st < n_st) {
                const uint32_t d = (base == ch.st0) ? d_first : drow[st];
                start = d & 0xffffu;
                c = d >> 16;
                if (st == ch.st0) { start += ch.skip; c -= ch.skip; }
            }
            uint32_t tot;
            const uint32_t dst = acc + v3_scan(c, s_warp, tot);
            uint32_t len = 0;
            if (c && dst < ch.n) len = min(c, ch.n - dst);
            r_src[tid] = (uint32_t)(((uint64_t)st * V3_ST1 + start) >> 2);
            r_dl[tid] = len ? (dst | (len << 16)) : 0u;  // len <= 12288, dst < 12288 (dst of a run past the chunk's end can be anything)
            acc += tot;
            // runs of the batch that contribute: [0, n_run) (the slots are consecutive, so they are a prefix... except
            // empty runs in between, which simply have len 0)
            const uint32_t n_run = min((uint32_t)V3_THREADS, n_st - base);
            __syncthreads();
            // 8 lanes per run, 16 bytes per lane and step
            for (uint32_t r = 4 * warp + (lane >> 3); r < n_run; r += 4 * (V3_THREADS / 32)) {
                const uint32_t dl = r_dl[r];
                const uint32_t ln = dl >> 16, d0 = dl & 0xffffu;
                const uint4* src = reinterpret_cast<const uint4*>(buf1) + r_src[r];
                for (uint32_t i = (lane & 7); 4 * i < ln; i += 8) spk_cp_async16(stage_s + 4 * (d0 + 4 * i), src + i);
            }
            __syncthreads();
        }
    };
    // (record of chunk g + 3G and first descriptor of chunk g + 2G are loaded per iteration: two independent loads)
    auto load_rec = [&](uint32_t g, V3Chunk& ch) {
        ch.st0 = ch.skip = ch.bucket = ch.n = 0;
        if (g < n_chunks) ch = chunks[g];
    };
    auto load_desc = [&](uint32_t g, const V3Chunk& ch) -> uint32_t {
        return (g < n_chunks && ch.st0 + tid < n_st) ? descT[(size_t)ch.bucket * n_st + ch.st0 + tid] : 0u;
    };
    uint32_t g = blockIdx.x;
    const uint32_t G = gridDim.x;
    V3Chunk cur_ch, nxt_ch, nn_ch;
    uint32_t nxt_d;
    load_rec(g, cur_ch);
    load_rec(g + G, nxt_ch);
    load_rec(g + 2 * G, nn_ch);
    if (g < n_chunks) gather_issue(cur_ch, load_desc(g, cur_ch));
    nxt_d = load_desc(g + G, nxt_ch);
    for (; g < n_chunks; g += G) {
        const uint32_t n_s = cur_ch.n;                    // slots (entries + padding)
        const uint32_t nn_d = load_desc(g + 2 * G, nn_ch);
        V3Chunk n3_ch;
        load_rec(g + 3 * G, n3_ch);
        spk_cp_async_wait_all();
        __syncthreads();
        uint32_t e[V3_E2], meta[V3_E2];
#pragma unroll
        for (int u = 0; u < V3_E2 / 4; u++) {
            const uint32_t i0 = 4u * (u * V3_THREADS + tid);
            uint4 v = make_uint4(V3_SENT, V3_SENT, V3_SENT, V3_SENT);
            if (i0 < n_s) v = *reinterpret_cast<const uint4*>(stage + i0);       // n_s is a multiple of 4
            const uint32_t vs[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int q = 0; q < 4; q++) {
                meta[4 * u + q] = 0xffffffffu;
                e[4 * u + q] = vs[q] & rmask;
                if (vs[q] != V3_SENT) {
                    const uint32_t sub = vs[q] >> rbits;
                    meta[4 * u + q] = (sub << 16) | atomicAdd(&cnt[sub], 1u);
                }
            }
        }
        __syncthreads();                                  // `stage` is consumed: the next chunk may stream in
        if (g + G < n_chunks) gather_issue(nxt_ch, nxt_d);
        uint32_t n_e;
        const uint32_t c = cnt[tid];
        const uint32_t o = v3_scan(c, s_warp, n_e);
        off[tid] = o;
        cnt[tid] = 0;
        uint16_t* orow = off2 + (size_t)g * (nb + 1);
        if (tid < nb) orow[tid] = (uint16_t)o;
        if (tid == 0) {
            orow[nb] = (uint16_t)n_e;
            spk_bulk_wait_read();
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < V3_E2; j++)
            if (meta[j] != 0xffffffffu) sorted[off[meta[j] >> 16] + (meta[j] & 0xffffu)] = e[j];
        spk_fence_proxy_async();
        __syncthreads();
        if (tid == 0 && n_e) spk_bulk_s2g(buf2 + (size_t)g * V3_CH2, sorted, (n_e * 4 + 15) & ~15u);
        cur_ch = nxt_ch;
        nxt_ch = nn_ch;
        nxt_d = nn_d;
        nn_ch = n3_ch;
    }
    if (tid == 0) spk_bulk_wait_all();
}

// ---- phase 2: one partition per CTA iteration, table in shared memory ---------------------------------------
__device__ __forceinline__ uint32_t fmix32(uint32_t h) {
    h ^= h >> 16;
    h *= 0x85ebca6bu;
    h ^= h >> 13;
    h *= 0xc2b2ae35u;
    h ^= h >> 16;
    return h;
}

struct CountOut {
    uint64_t* keys;
    uint32_t* counts;
    uint64_t cap;
    uint64_t* cursor;
    uint64_t* stats;
    uint64_t* histo;
    uint32_t histo_len;
    uint32_t lower;
    uint32_t* pindex;   // optional [2P]: first dump index / number of dumped entries of every partition
};

// ENT32: slot u64 = (r << 32) | count, empty = 0 (count >= 1 once occupied).
// ENT64: key u64 (empty = ~0) and count u32 in separate arrays.
// The table is cleared once per CTA; every insert that creates a key appends its slot to a list, so the
// dump visits (and re-clears) only the occupied slots — the zero + full-scan loops of a naive version
// cost more instructions than the inserts themselves.
// The dump of a partition is contiguous (one reservation per partition, recorded in pindex): the
// partitioned union/filter (spk_pmatrix.cu) merges the same partition of every chromosome on-chip.
constexpr int PC_LIST = 4096;   // occupied-slot list capacity; beyond it a partition falls back to a full scan

template <bool ENT64>
__global__ void __launch_bounds__(PC_THREADS, 3)
k_part_count(const void* __restrict__ buf, const uint32_t* __restrict__ pstart, uint64_t P, Mixer mx,
             CountOut o) {
    extern __shared__ __align__(16) uint8_t s_raw[];
    uint64_t* s_key = (uint64_t*)s_raw;
    uint16_t* s_list = (uint16_t*)(s_raw + (size_t)PC_SLOTS * 8);
    uint32_t* s_cnt = (uint32_t*)(s_raw + (size_t)PC_SLOTS * 8 + (size_t)PC_LIST * 2);  // ENT64 only
    __shared__ uint32_t s_hist[256];
    __shared__ uint64_t s_red[2][PC_THREADS / 32];
    __shared__ uint32_t s_nocc, s_nkeep, s_wr;
    __shared__ uint64_t s_base;
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    uint64_t distinct = 0, nge = 0, sumge = 0, sumall = 0, n_fail = 0;   // distinct/sumall: thread 0 only
    uint32_t h1 = 0, h2 = 0;
    if (o.histo) s_hist[tid] = 0;
    const uint64_t EMPTY = ENT64 ? SPK_EMPTY_KEY : 0ull;
    constexpr uint32_t TMASK = PC_SLOTS - 1;
    for (uint32_t i = tid; i < PC_SLOTS; i += PC_THREADS) {
        s_key[i] = EMPTY;
        if (ENT64) s_cnt[i] = 0;
    }
    if (tid == 0) {
        s_nocc = 0;
        s_nkeep = 0;
        s_wr = 0;
    }
    __syncthreads();

    auto slot_count = [&](uint32_t slot, uint64_t& r) -> uint64_t {   // 0: empty
        const uint64_t v = s_key[slot];
        if (v == EMPTY) return 0;
        r = ENT64 ? v : (v >> 32);
        return ENT64 ? (uint64_t)s_cnt[slot] : (v & 0xffffffffull);
    };
    // one slot of the table -> histogram / dump; clears the slot
    auto visit = [&](uint32_t slot, bool active, uint64_t p) {
        uint64_t r = 0, cnt = 0;
        if (active) {
            cnt = slot_count(slot, r);
            if (cnt) {
                s_key[slot] = EMPTY;
                if (ENT64) s_cnt[slot] = 0;
                if (o.histo) {
                    const uint64_t b = cnt < (uint64_t)(o.histo_len - 1) ? cnt : (uint64_t)(o.histo_len - 1);
                    if (b == 1) h1++;
                    else if (b == 2) h2++;
                    else if (b < 256) atomicAdd(&s_hist[b], 1u);
                    else atomicAdd((unsigned long long*)&o.histo[b], 1ull);
                }
            }
        }
        const bool keep = cnt >= o.lower && cnt > 0;
        const uint32_t ballot = __ballot_sync(0xffffffffu, keep);
        if (ballot) {
            uint32_t wbase = 0;
            if (lane == 0) wbase = atomicAdd(&s_wr, (uint32_t)__popc(ballot));
            wbase = __shfl_sync(0xffffffffu, wbase, 0);
            if (keep) {
                nge++;
                sumge += cnt;
                const uint64_t at = s_base + wbase + __popc(ballot & ((1u << lane) - 1));
                if (at < o.cap) {
                    o.keys[at] = mx.inv((p << mx.rbits) | r);
                    o.counts[at] = (uint32_t)cnt;
                }
            }
        }
    };

    // insert one remainder into the shared-memory table
    auto insert = [&](uint64_t r) {
        uint32_t s = (ENT64 ? (uint32_t)spk_hash64(r) : fmix32((uint32_t)r)) & TMASK;
        bool done = false;
        for (uint32_t probes = 0; probes < PC_SLOTS; probes++) {
            uint64_t c = s_key[s];
            if (c == EMPTY) {
                const uint64_t fresh = ENT64 ? r : ((r << 32) | 1ull);
                const uint64_t old = atomicCAS((unsigned long long*)&s_key[s], (unsigned long long)EMPTY,
                                               (unsigned long long)fresh);
                if (old == EMPTY) {
                    if (ENT64) atomicAdd(&s_cnt[s], 1u);
                    const uint32_t li = atomicAdd(&s_nocc, 1u);
                    if (li < PC_LIST) s_list[li] = (uint16_t)s;
                    done = true;
                    break;
                }
                c = old;
            }
            if (ENT64 ? (c == r) : ((c >> 32) == r)) {
                if (ENT64) atomicAdd(&s_cnt[s], 1u);
                else atomicAdd((unsigned int*)&s_key[s], 1u);   // low word = count (little endian)
                done = true;
                break;
            }
            s = (s + 1) & TMASK;
        }
        if (!done) n_fail++;
    };
    using RegT = typename std::conditional<ENT64, uint64_t, uint32_t>::type;
    auto load_r = [&](uint32_t i) -> RegT { return __ldcs((const RegT*)buf + i); };

    // The partition stream is software-pipelined through registers: while partition p is inserted, the
    // first PC_PF * 256 entries of the CTA's next partition (and the extents of the one after) are already
    // in flight, so the global-load latency (the top stall of the unpipelined version) is hidden.
    constexpr int PC_PF = ENT64 ? 8 : 16;
    const uint64_t G = gridDim.x;
    uint64_t p = blockIdx.x;
    uint32_t beg = 0, end = 0, nbeg = 0, nend = 0;
    if (p < P) { beg = pstart[p]; end = pstart[p + 1]; }
    if (p + G < P) { nbeg = pstart[p + G]; nend = pstart[p + G + 1]; }
    RegT cur[PC_PF];
#pragma unroll
    for (int u = 0; u < PC_PF; u++) {
        const uint32_t i = beg + u * PC_THREADS + tid;
        cur[u] = i < end ? load_r(i) : 0;
    }
    for (; p < P; p += G) {
        RegT nxt[PC_PF];
#pragma unroll
        for (int u = 0; u < PC_PF; u++) {
            const uint32_t i = nbeg + u * PC_THREADS + tid;
            nxt[u] = i < nend ? load_r(i) : 0;
        }
        uint32_t nnbeg = 0, nnend = 0;
        if (p + 2 * G < P) { nnbeg = pstart[p + 2 * G]; nnend = pstart[p + 2 * G + 1]; }
        // ---- insert ----
#pragma unroll
        for (int u = 0; u < PC_PF; u++)
            if (beg + u * PC_THREADS + tid < end) insert(cur[u]);
        for (uint32_t base = beg + PC_PF * PC_THREADS; base < end; base += 4 * PC_THREADS) {   // oversized partition
            RegT t[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const uint32_t i = base + u * PC_THREADS + tid;
                t[u] = i < end ? load_r(i) : 0;
            }
#pragma unroll
            for (int u = 0; u < 4; u++)
                if (base + u * PC_THREADS + tid < end) insert(t[u]);
        }
        __syncthreads();
        // ---- dump pass 1: how many entries does this partition dump? (one reservation per partition) ----
        const uint32_t nocc = s_nocc;
        const bool listed = nocc <= PC_LIST;
        const uint32_t nvis = listed ? nocc : (uint32_t)PC_SLOTS;
        uint32_t myk = 0;
        for (uint32_t i = tid; i < nvis; i += PC_THREADS) {
            uint64_t r;
            const uint64_t cnt = slot_count(listed ? (uint32_t)s_list[i] : i, r);
            myk += (cnt >= o.lower && cnt > 0) ? 1u : 0u;
        }
        myk = spk_warp_sum_u32(myk);
        if (lane == 0 && myk) atomicAdd(&s_nkeep, myk);
        __syncthreads();
        if (tid == 0) {
            const uint32_t nk = s_nkeep;
            const uint64_t base = nk ? atomicAdd((unsigned long long*)o.cursor, (unsigned long long)nk) : 0ull;
            s_base = base;
            if (o.pindex) {
                o.pindex[2 * p] = (uint32_t)base;
                o.pindex[2 * p + 1] = nk;
            }
            distinct += nocc;
            sumall += end - beg;
        }
        __syncthreads();
        // ---- dump pass 2: histogram, write, clear ----
        for (uint32_t base = 0; base < nvis; base += PC_THREADS) {
            const uint32_t i = base + tid;
            visit(listed ? (i < nocc ? (uint32_t)s_list[i] : 0u) : i, i < nvis, p);
        }
        __syncthreads();
        if (tid == 0) {
            s_nocc = 0;
            s_nkeep = 0;
            s_wr = 0;
        }
        __syncthreads();
#pragma unroll
        for (int u = 0; u < PC_PF; u++) cur[u] = nxt[u];
        beg = nbeg; end = nend; nbeg = nnbeg; nend = nnend;
    }
    // ---- per-CTA totals ----
    __syncthreads();
    if (o.histo) {
        h1 = spk_warp_sum_u32(h1);
        h2 = spk_warp_sum_u32(h2);
        if (lane == 0) {
            if (h1) atomicAdd(&s_hist[1], h1);
            if (h2) atomicAdd(&s_hist[2], h2);
        }
        __syncthreads();
        if ((uint32_t)tid < o.histo_len && s_hist[tid])
            atomicAdd((unsigned long long*)&o.histo[tid], (unsigned long long)s_hist[tid]);
    }
    uint64_t v[2] = {nge, sumge};
#pragma unroll
    for (int q = 0; q < 2; q++) {
        v[q] = spk_warp_sum_u64(v[q]);
        if (lane == 0) s_red[q][tid >> 5] = v[q];
    }
    n_fail = spk_warp_sum_u64(n_fail);
    if (lane == 0 && n_fail) atomicAdd((unsigned long long*)&o.stats[1], (unsigned long long)n_fail);
    __syncthreads();
    if (tid < 2) {
        uint64_t s = 0;
        for (int w = 0; w < PC_THREADS / 32; w++) s += s_red[tid][w];
        if (s) atomicAdd((unsigned long long*)&o.stats[5 + tid], (unsigned long long)s);
    }
    if (tid == 0) {
        if (distinct) atomicAdd((unsigned long long*)&o.stats[4], (unsigned long long)distinct);
        if (sumall) atomicAdd((unsigned long long*)&o.stats[7], (unsigned long long)sumall);
    }
}

// ---- phase 2, 32-bit remainders (rbits <= 31: every practical k / chromosome size) --------------------------
// Same contract as k_part_count<false>, leaner instruction stream (the 64-bit-slot kernel spent ~5 warp
// instructions per k-mer, half of them in divergent probe loops, and stalled on instruction fetch):
//  * keys and counts are separate 32-bit arrays; the FIRST probe of every entry is straight-line code —
//    `old = CAS(key[s], EMPTY, r); if (old == EMPTY || old == r) c = count[s]++` — identical for a new key and
//    a hit, no loop, no divergence; the ~20 % of entries whose home slot holds another key go to a
//    shared-memory retry queue and are probed afterwards by all lanes together;
//  * distinct keys and dump size are counted where they happen (old == EMPTY; the add that makes a count
//    reach lower_count), so there is no sizing pass: one reservation per partition, then ONE 128-bit sweep
//    over the 8192 slots that histograms, writes the kept entries (contiguous per partition) and clears.
constexpr uint32_t PC_EMPTY32 = 0xffffffffu;
constexpr int PC_RETRY = 2048;          // retry-queue capacity (entries beyond it probe inline)

// full linear probing from the slot after home (out of line: rare in the first-probe loop, dense in the drain)
// -> bit 0: created the key, bit 1: its count reached `lower` (bits 16..: slot + 1), bit 2: table full
__device__ __noinline__ uint32_t pc32_probe_rest(uint32_t* s_key, uint32_t* s_cnt, uint32_t r, uint32_t lower) {
    constexpr uint32_t TMASK = PC_SLOTS - 1;
    uint32_t s = ((r & TMASK) + 1) & TMASK;
    for (uint32_t probes = 1; probes < PC_SLOTS; probes++) {
        const uint32_t old = atomicCAS(&s_key[s], PC_EMPTY32, r);
        if (old == PC_EMPTY32 || old == r) {
            const uint32_t c = atomicAdd(&s_cnt[s], 1u);
            return ((old == PC_EMPTY32) ? 1u : 0u) | ((c + 1 == lower) ? (2u | ((s + 1) << 16)) : 0u);
        }
        s = (s + 1) & TMASK;
    }
    return 4u;
}

// GATHER (v3 pipeline): partition p = (bucket b, sub-partition) is not contiguous; it is one run per chunk of bucket
// b: run r = [off2[g][sub], off2[g][sub + 1]) of chunk g = cb[b] + r.  Warp w owns the runs w, w + 8, ...; lane l
// loads the descriptor of run w + 8 l two partitions ahead, the first 32 entries of the warp's first 16 runs are
// loaded one partition ahead (lane i = entry i of the run), longer / later runs are read when they are inserted.
// ---- versioned table (VER): no clearing, no sweep ---------------------------------------------------------------
// A slot is 64 bits: high word = (generation << rbits) | remainder, low word = count.  The CTA bumps its generation
// for every partition, so a slot written for an earlier partition is simply "free" — the table is never cleared —
// and the k-mers to dump (count >= lower) are collected WHEN their count reaches `lower` (that add returns lower - 1
// exactly once per key) in a small slot list, so the 8192-slot sweep that found them is gone too.  Claiming a free
// slot is one 64-bit CAS that installs key and count = 1 together; a hit is a 32-bit add on the low word.
// (A count histogram (`-histo`), or more kept keys than the list holds, falls back to scanning the table.)
constexpr int PC_KEEP = 2048;            // kept-slot list capacity (u16 slot ids)
constexpr int PC_RETRY_V = 1024;         // retry queue of the versioned kernel

// -> bit 0: created the key, bit 1: table full, bits 16..: slot + 1 when the count reached `lower` there
__device__ __noinline__ uint32_t pcv_probe_rest(unsigned long long* s_tab, uint32_t r, uint32_t want, uint32_t gen,
                                                int rbits, uint32_t lower) {
    constexpr uint32_t TMASK = PC_SLOTS - 1;
    uint32_t s = ((r & TMASK) + 1) & TMASK;
    for (uint32_t probes = 1; probes < PC_SLOTS; probes++) {
        unsigned long long v = s_tab[s];
        while (true) {
            const uint32_t hi = (uint32_t)(v >> 32);
            if (hi == want) {
                const uint32_t c = atomicAdd((unsigned int*)&s_tab[s], 1u);        // low word (little endian)
                return (c + 1 == lower) ? ((s + 1) << 16) : 0u;
            }
            if ((hi >> rbits) == gen) break;                                       // another key of this partition
            const unsigned long long old = atomicCAS(&s_tab[s], v, ((unsigned long long)want << 32) | 1ull);
            if (old == v) return 1u | ((lower == 1) ? ((s + 1) << 16) : 0u);
            v = old;                                                               // lost the race: look again
        }
        s = (s + 1) & TMASK;
    }
    return 2u;
}

struct GatherIn {
    const uint32_t* cb;       // [2^b1 + 1] first chunk of every bucket
    const uint16_t* off2;     // [(2^b2 + 1) per chunk]
    int b2;
};

// LIST (default when no count histogram is wanted): the 32-bit key / count arrays of the sweep kernel, but the slots
// to dump are collected where a count reaches `lower` (u16 list, as in the versioned table), read into registers
// before the reservation, and the table is then cleared with stores only — the sweep's 2 x 64 KB of shared-memory
// loads, its per-slot tests and two of the six barriers per partition are gone.  More kept keys than the list holds
// (a partition of nothing but repeats, lower_count 1): that partition is swept as before.
constexpr int PC_RETRY_L = 1792;         // retry queue of the list kernel (same 72 KB per CTA as the sweep kernel)
constexpr int PC_KEEP_L = 512;           // kept-slot list capacity (2 per thread)

template <bool GATHER, bool VER, bool LIST = false>
__global__ void __launch_bounds__(PC_THREADS, 3)
k_part_count32(const uint32_t* __restrict__ buf, const uint32_t* __restrict__ pstart, uint64_t P, Mixer mx,
               CountOut o, uint32_t retry_cap, GatherIn gi) {
    static_assert(!(VER && LIST), "one table variant");
    extern __shared__ __align__(16) uint8_t s_raw[];
    constexpr int QN = VER ? PC_RETRY_V : (LIST ? PC_RETRY_L : PC_RETRY);   // words of s_q
    constexpr uint32_t KEEPN = VER ? PC_KEEP : PC_KEEP_L;
    uint32_t* s_key = (uint32_t*)s_raw;
    uint32_t* s_cnt = s_key + PC_SLOTS;
    uint32_t* s_q = s_cnt + PC_SLOTS;                    // [QN]; retry_cap <= QN entries are used as queue
    unsigned long long* s_tab = (unsigned long long*)s_raw;           // VER: [PC_SLOTS] 64-bit slots (same bytes)
    uint16_t* s_keep = (uint16_t*)(s_q + QN);            // VER / LIST: [KEEPN]
    __shared__ uint32_t s_nlist;
    __shared__ uint32_t s_hist[256];
    __shared__ uint64_t s_red[4][PC_THREADS / 32];
    __shared__ uint32_t s_nq, s_nkeep, s_ndist, s_wr;
    __shared__ uint64_t s_base;
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    uint64_t distinct = 0, nge = 0, sumge = 0, sumall = 0, n_fail = 0;
    uint32_t h1 = 0, h2 = 0;
    if (o.histo) s_hist[tid] = 0;
    constexpr uint32_t TMASK = PC_SLOTS - 1;
    constexpr int SWEEP = PC_SLOTS / (PC_THREADS * 4);   // uint4 loads per thread per sweep (8)
    for (uint32_t i = tid; i < PC_SLOTS; i += PC_THREADS) {
        s_key[i] = VER ? 0u : PC_EMPTY32;                 // VER: generation 0 = never used
        s_cnt[i] = 0;
    }
    if (tid == 0) {
        s_nq = 0;
        s_nkeep = 0;
        s_ndist = 0;
        s_wr = 0;
        s_nlist = 0;
    }
    __syncthreads();
    const uint32_t lower = o.lower;
    uint32_t my_new = 0, my_keep = 0;    // per partition: keys created / counts that reached `lower` by this thread
    const int rbits = mx.rbits;                                      // VER: <= 31
    const uint32_t gmax = VER ? ((rbits >= 32) ? 0u : (0xffffffffu >> rbits)) : 0u;
    uint32_t gen = 1;                                                // VER: generation of the current partition
    auto keep_slot = [&](uint32_t slot) {                            // VER / LIST: this key is dumped: remember where it lives
        const uint32_t at = atomicAdd(&s_nlist, 1u);
        if (at < KEEPN) s_keep[at] = (uint16_t)slot;
    };

    // first probe only; false: the home slot belongs to another key
    auto try_home = [&](uint32_t r) -> bool {
        const uint32_t s = r & TMASK;                    // r = low bits of the mixer output: already uniform
        if constexpr (VER) {
            const uint32_t want = (gen << rbits) | r;
            unsigned long long v = s_tab[s];
            while (true) {
                const uint32_t hi = (uint32_t)(v >> 32);
                if (hi == want) {
                    const uint32_t c = atomicAdd((unsigned int*)&s_tab[s], 1u);    // low word = count
                    if (c + 1 == lower) keep_slot(s);
                    return true;
                }
                if ((hi >> rbits) == gen) return false;                            // another key of this partition
                const unsigned long long old = atomicCAS(&s_tab[s], v, ((unsigned long long)want << 32) | 1ull);
                if (old == v) {
                    my_new++;
                    if (lower == 1) keep_slot(s);
                    return true;
                }
                v = old;
            }
        } else {
            const uint32_t old = atomicCAS(&s_key[s], PC_EMPTY32, r);
            if (old == PC_EMPTY32 || old == r) {
                const uint32_t c = atomicAdd(&s_cnt[s], 1u);
                my_new += (old == PC_EMPTY32) ? 1u : 0u;
                if constexpr (LIST) {
                    if (c + 1 == lower) keep_slot(s);
                } else {
                    my_keep += (c + 1 == lower) ? 1u : 0u;
                }
                return true;
            }
            return false;
        }
    };
    auto probe_rest = [&](uint32_t r) {
        if constexpr (VER) {
            const uint32_t f = pcv_probe_rest(s_tab, r, (gen << rbits) | r, gen, rbits, lower);
            my_new += f & 1u;
            n_fail += (f >> 1) & 1u;
            if (f >> 16) keep_slot((f >> 16) - 1);
        } else {
            const uint32_t f = pc32_probe_rest(s_key, s_cnt, r, lower);
            my_new += f & 1u;
            if constexpr (LIST) {
                if (f >> 16) keep_slot((f >> 16) - 1);
            } else {
                my_keep += (f >> 1) & 1u;
            }
            n_fail += (f >> 2) & 1u;
        }
    };
    auto insert = [&](uint32_t r) {
        if (!try_home(r)) {
            const uint32_t qi = atomicAdd(&s_nq, 1u);
            if (qi < retry_cap) s_q[qi] = r;
            else probe_rest(r);
        }
    };

    constexpr int PC_PF = 16;
    constexpr int NW = PC_THREADS / 32;
    const uint64_t G = gridDim.x;
    uint64_t p = blockIdx.x;
    // contiguous input: [beg, end) of the partition stream; gathered input: this lane's run descriptor (src, len)
    // and the number of runs (chunks of the bucket) for the current / next / next-next partition of the CTA
    uint32_t beg = 0, end = 0, nbeg = 0, nend = 0, cnc = 0, nnc = 0;
    // two dependent loads (chunk range of the bucket, then the run offsets): split over two loop iterations, so
    // that neither is waited for where it is issued
    auto range_load = [&](uint64_t pp, uint32_t& g0, uint32_t& g1) {
        g0 = g1 = 0;
        if (pp < P) {
            const uint32_t b = (uint32_t)(pp >> gi.b2);
            g0 = __ldg(gi.cb + b);
            g1 = __ldg(gi.cb + b + 1);
        }
    };
    auto desc_load = [&](uint64_t pp, uint32_t g0, uint32_t g1, uint32_t& src, uint32_t& len, uint32_t& nc) {
        src = 0; len = 0;
        nc = g1 - g0;
        const uint32_t r = warp + NW * lane;
        if (pp < P && r < nc) {
            const uint32_t nb2 = 1u << gi.b2;
            const uint32_t sub = (uint32_t)pp & (nb2 - 1);
            const uint16_t* q = gi.off2 + (size_t)(g0 + r) * (nb2 + 1) + sub;
            const uint32_t a = __ldg(q), z = __ldg(q + 1);
            src = (g0 + r) * (uint32_t)V3_CH2 + a;
            len = z - a;
        }
    };
    uint32_t rg0 = 0, rg1 = 0;       // chunk range of the bucket of partition p + 3G
    uint32_t cur[PC_PF];
    if constexpr (GATHER) {
        range_load(p, rg0, rg1);
        desc_load(p, rg0, rg1, beg, end, cnc);
        range_load(p + G, rg0, rg1);
        desc_load(p + G, rg0, rg1, nbeg, nend, nnc);
        range_load(p + 2 * G, rg0, rg1);
#pragma unroll
        for (int u = 0; u < PC_PF; u++) {
            const uint32_t sr = __shfl_sync(0xffffffffu, beg, u), ln = __shfl_sync(0xffffffffu, end, u);
            cur[u] = (uint32_t)lane < ln ? __ldcs(buf + sr + lane) : 0;
        }
    } else {
        if (p < P) { beg = pstart[p]; end = pstart[p + 1]; }
        if (p + G < P) { nbeg = pstart[p + G]; nend = pstart[p + G + 1]; }
#pragma unroll
        for (int u = 0; u < PC_PF; u++) {
            const uint32_t i = beg + u * PC_THREADS + tid;
            cur[u] = i < end ? __ldcs(buf + i) : 0;
        }
    }
    for (; p < P; p += G) {
        uint32_t nxt[PC_PF];
        uint32_t nnbeg = 0, nnend = 0, nnnc = 0;
        if constexpr (GATHER) {
#pragma unroll
            for (int u = 0; u < PC_PF; u++) {
                const uint32_t sr = __shfl_sync(0xffffffffu, nbeg, u), ln = __shfl_sync(0xffffffffu, nend, u);
                nxt[u] = (uint32_t)lane < ln ? __ldcs(buf + sr + lane) : 0;
            }
            desc_load(p + 2 * G, rg0, rg1, nnbeg, nnend, nnnc);
            range_load(p + 3 * G, rg0, rg1);
        } else {
#pragma unroll
            for (int u = 0; u < PC_PF; u++) {
                const uint32_t i = nbeg + u * PC_THREADS + tid;
                nxt[u] = i < nend ? __ldcs(buf + i) : 0;
            }
            if (p + 2 * G < P) { nnbeg = pstart[p + 2 * G]; nnend = pstart[p + 2 * G + 1]; }
        }
        // gathered input: what the prefetch did not cover (entries 32.. of a run, the runs of lanes 16..31) is loaded
        // here, before the first probes, so that its latency hides under them (first NLO such runs; the rest below)
        constexpr int NLO = 2;
        uint32_t lo[NLO], lov = 0, todo = 0, todo_long = 0;
        if constexpr (GATHER) {
            todo = __ballot_sync(0xffffffffu, end > (lane < PC_PF ? 32u : 0u));
#pragma unroll
            for (int q = 0; q < NLO; q++) {
                lo[q] = 0;
                if (todo) {                                   // warp-uniform
                    const int u = __ffs(todo) - 1;
                    todo &= todo - 1;
                    const uint32_t sr = __shfl_sync(0xffffffffu, beg, u), ln = __shfl_sync(0xffffffffu, end, u);
                    const uint32_t i = (u < PC_PF ? 32u : 0u) + lane;
                    if (i < ln) {
                        lo[q] = __ldcs(buf + sr + i);
                        lov |= 1u << q;
                    }
                    if (ln > (u < PC_PF ? 64u : 32u)) todo_long |= 1u << u;
                }
            }
        }
        // ---- insert, first probes (straight-line); failures are queued with one reservation per thread ----
        uint32_t failm = 0;
#pragma unroll
        for (int u = 0; u < PC_PF; u++) {
            bool have;
            if constexpr (GATHER) have = (uint32_t)lane < __shfl_sync(0xffffffffu, end, u);
            else have = beg + u * PC_THREADS + tid < end;
            if (have && !try_home(cur[u])) failm |= 1u << u;
        }
        if (failm) {
            uint32_t qi = atomicAdd(&s_nq, (uint32_t)__popc(failm));
            if (qi + __popc(failm) <= retry_cap) {
#pragma unroll
                for (int u = 0; u < PC_PF; u++)
                    if ((failm >> u) & 1u) s_q[qi++] = cur[u];
            } else {                                        // queue (nearly) full — a heavily colliding partition:
#pragma unroll                                              // fill what is left of the reservation, probe the rest inline
                for (int u = 0; u < PC_PF; u++)
                    if ((failm >> u) & 1u) {
                        if (qi < retry_cap) s_q[qi] = cur[u];     // every slot below min(s_nq, retry_cap) must be written
                        else probe_rest(cur[u]);
                        qi++;
                    }
            }
        }
        if constexpr (GATHER) {
            // what the prefetch did not cover: entries 32.. of the first 16 runs of the warp, the runs of lanes 16..31,
            // and (a bucket with more than 256 chunks) the runs beyond the descriptor registers
#pragma unroll
            for (int q = 0; q < NLO; q++)
                if ((lov >> q) & 1u) insert(lo[q]);
            while (todo | todo_long) {                        // rare: more than NLO such runs, or a run beyond 64 entries
                const bool lg = todo == 0;
                const uint32_t m = lg ? todo_long : todo;
                const int u = __ffs(m) - 1;
                if (lg) todo_long &= todo_long - 1;
                else todo &= todo - 1;
                const uint32_t sr = __shfl_sync(0xffffffffu, beg, u), ln = __shfl_sync(0xffffffffu, end, u);
                for (uint32_t i = (u < PC_PF ? 32u : 0u) + (lg ? 32u : 0u) + lane; i < ln; i += 32)
                    insert(__ldcs(buf + sr + i));
            }
            if (cnc > 32 * NW) {
                const uint32_t nb2 = 1u << gi.b2;
                const uint32_t b = (uint32_t)(p >> gi.b2), sub = (uint32_t)p & (nb2 - 1);
                const uint32_t g0 = __ldg(gi.cb + b);
                for (uint32_t r = 32 * NW + warp; r < cnc; r += NW) {
                    const uint16_t* q = gi.off2 + (size_t)(g0 + r) * (nb2 + 1) + sub;
                    const uint32_t a = __ldg(q), z = __ldg(q + 1);
                    const uint32_t sr = (g0 + r) * (uint32_t)V3_CH2 + a;
                    for (uint32_t i = a + lane; i < z; i += 32) insert(__ldcs(buf + sr + (i - a)));
                    if (lane == 0) sumall += z - a;
                }
            }
            sumall += end;                      // this lane's run (every run is owned by exactly one lane)
        } else {
            for (uint32_t base = beg + PC_PF * PC_THREADS; base < end; base += 4 * PC_THREADS) {   // oversized partition
                uint32_t t[4];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const uint32_t i = base + u * PC_THREADS + tid;
                    t[u] = i < end ? __ldcs(buf + i) : 0;
                }
#pragma unroll
                for (int u = 0; u < 4; u++)
                    if (base + u * PC_THREADS + tid < end) insert(t[u]);
            }
        }
        __syncthreads();
        // ---- insert, queued entries: all lanes probe together ----
        const uint32_t nq = min(s_nq, retry_cap);
        for (uint32_t i = tid; i < nq; i += PC_THREADS) probe_rest(s_q[i]);
        // ---- partition totals -> one reservation ----
        my_new = spk_warp_sum_u32(my_new);
        my_keep = spk_warp_sum_u32(my_keep);
        if (lane == 0) {
            if (my_new) atomicAdd(&s_ndist, my_new);
            if (my_keep) atomicAdd(&s_nkeep, my_keep);
        }
        my_new = my_keep = 0;
        __syncthreads();
        // LIST: the kept entries go to registers now (the table is cleared by other threads after the next barrier)
        constexpr int KR = PC_KEEP_L / PC_THREADS;
        uint32_t lk[KR], lc[KR];
        const uint32_t nl_list = LIST ? s_nlist : 0u;
        const bool list_ok = LIST && nl_list <= KEEPN && !o.histo;          // (block-uniform)
        if constexpr (LIST) {
            if (list_ok) {
#pragma unroll
                for (int q = 0; q < KR; q++) {
                    const uint32_t i = (uint32_t)tid + (uint32_t)q * PC_THREADS;
                    lk[q] = lc[q] = 0;
                    if (i < nl_list) {
                        const uint32_t sl = s_keep[i];
                        lk[q] = s_key[sl];
                        lc[q] = s_cnt[sl];
                    }
                }
            }
        }
        if (tid == 0) {
            const uint32_t total = (VER || LIST) ? s_nlist : s_nkeep;
            const uint64_t base = total ? atomicAdd((unsigned long long*)o.cursor, (unsigned long long)total) : 0ull;
            s_base = base;
            if (o.pindex) {
                o.pindex[2 * p] = (uint32_t)base;
                o.pindex[2 * p + 1] = total;
            }
            distinct += s_ndist;
            if constexpr (!GATHER) sumall += end - beg;
        }
        __syncthreads();
        const uint64_t pbase = s_base;
        if constexpr (VER) {
            // ---- dump from the kept-slot list (no sweep, nothing to clear) ----
            const uint32_t nl = s_nlist;
            const uint32_t rmask = (rbits >= 32) ? 0xffffffffu : ((1u << rbits) - 1u);
            if (nl <= PC_KEEP && !o.histo) {
                for (uint32_t i = tid; i < nl; i += PC_THREADS) {
                    const unsigned long long v = s_tab[s_keep[i]];
                    const uint32_t cnt = (uint32_t)v;
                    nge++;
                    sumge += cnt;
                    if (pbase + i < o.cap) {
                        o.keys[pbase + i] = mx.inv((p << mx.rbits) | (uint64_t)((uint32_t)(v >> 32) & rmask));
                        o.counts[pbase + i] = cnt;
                    }
                }
            } else {
                // histogram wanted, or more kept keys than the list holds: scan the slots of this generation
                for (uint32_t sl = tid; sl < PC_SLOTS; sl += PC_THREADS) {
                    const unsigned long long v = s_tab[sl];
                    const uint32_t hi = (uint32_t)(v >> 32), cnt = (uint32_t)v;
                    if ((hi >> rbits) != gen) continue;
                    if (o.histo) {
                        const uint32_t b = cnt < o.histo_len - 1 ? cnt : o.histo_len - 1;
                        if (b == 1) h1++;
                        else if (b == 2) h2++;
                        else if (b < 256) atomicAdd(&s_hist[b], 1u);
                        else atomicAdd((unsigned long long*)&o.histo[b], 1ull);
                    }
                    if (cnt >= lower) {
                        const uint32_t pos = atomicAdd(&s_wr, 1u);
                        nge++;
                        sumge += cnt;
                        if (pbase + pos < o.cap) {
                            o.keys[pbase + pos] = mx.inv((p << mx.rbits) | (uint64_t)(hi & rmask));
                            o.counts[pbase + pos] = cnt;
                        }
                    }
                }
            }
            __syncthreads();
            if (gen == gmax) {                            // generation space used up: start over with a clean table
                for (uint32_t i = tid; i < PC_SLOTS; i += PC_THREADS) s_tab[i] = 0ull;
                gen = 0;
            }
            gen++;
            if (tid == 0) s_nlist = 0;
        }
        if constexpr (LIST) {
            if (list_ok) {
                // ---- write the kept entries from registers, clear the table with stores only; no further barrier ----
#pragma unroll
                for (int q = 0; q < KR; q++) {
                    const uint32_t i = (uint32_t)tid + (uint32_t)q * PC_THREADS;
                    if (i < nl_list) {
                        nge++;
                        sumge += lc[q];
                        if (pbase + i < o.cap) {
                            o.keys[pbase + i] = mx.inv((p << mx.rbits) | (uint64_t)lk[q]);
                            o.counts[pbase + i] = lc[q];
                        }
                    }
                }
#pragma unroll
                for (int i = 0; i < SWEEP; i++) {
                    const uint32_t idx = (i * PC_THREADS + tid) * 4;
                    *reinterpret_cast<uint4*>(s_key + idx) = make_uint4(PC_EMPTY32, PC_EMPTY32, PC_EMPTY32, PC_EMPTY32);
                    *reinterpret_cast<uint4*>(s_cnt + idx) = make_uint4(0u, 0u, 0u, 0u);
                }
                if (tid == 0) {
                    s_nq = 0;
                    s_nkeep = 0;
                    s_ndist = 0;
                    s_nlist = 0;
                }
                __syncthreads();
#pragma unroll
                for (int u = 0; u < PC_PF; u++) cur[u] = nxt[u];
                beg = nbeg; end = nend; nbeg = nnbeg; nend = nnend;
                cnc = nnc; nnc = nnnc;
                continue;
            }
        }
        // ---- sweep: histogram, write the dump, clear ----
#pragma unroll 1
        for (int i = 0; i < (VER ? 0 : SWEEP); i++) {
            const uint32_t idx = (i * PC_THREADS + tid) * 4;
            const uint4 c = *reinterpret_cast<const uint4*>(s_cnt + idx);
            const bool any = (c.x | c.y | c.z | c.w) != 0;
            if (__ballot_sync(0xffffffffu, any) == 0) continue;          // warp-uniform: 128 empty slots
            if (!any) continue;
            const uint4 kk = *reinterpret_cast<const uint4*>(s_key + idx);
            *reinterpret_cast<uint4*>(s_key + idx) = make_uint4(PC_EMPTY32, PC_EMPTY32, PC_EMPTY32, PC_EMPTY32);
            *reinterpret_cast<uint4*>(s_cnt + idx) = make_uint4(0u, 0u, 0u, 0u);
            const uint32_t cs[4] = {c.x, c.y, c.z, c.w}, ks[4] = {kk.x, kk.y, kk.z, kk.w};
            if (o.histo) {
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    const uint32_t cnt = cs[e];
                    if (cnt == 0) continue;
                    const uint32_t b = cnt < o.histo_len - 1 ? cnt : o.histo_len - 1;
                    if (b == 1) h1++;
                    else if (b == 2) h2++;
                    else if (b < 256) atomicAdd(&s_hist[b], 1u);
                    else atomicAdd((unsigned long long*)&o.histo[b], 1ull);
                }
            }
            // kept entries are rare (a few per cent of the distinct keys): per-thread reservation in smem
            const uint32_t km = (c.x >= lower ? 1u : 0u) | (c.y >= lower ? 2u : 0u) | (c.z >= lower ? 4u : 0u) |
                                (c.w >= lower ? 8u : 0u);                 // lower >= 1: implies occupied
            if (km) {
                uint32_t pos = atomicAdd(&s_wr, (uint32_t)__popc(km));
#pragma unroll
                for (int e = 0; e < 4; e++)
                    if ((km >> e) & 1u) {
                        if (pos < (uint32_t)(QN / 2)) {      // staged (the retry queue is free by now): written below by all lanes
                            s_q[2 * pos] = ks[e];
                            s_q[2 * pos + 1] = cs[e];
                        } else {
                            nge++;
                            sumge += cs[e];
                            if (pbase + pos < o.cap) {
                                o.keys[pbase + pos] = mx.inv((p << mx.rbits) | (uint64_t)ks[e]);
                                o.counts[pbase + pos] = cs[e];
                            }
                        }
                        pos++;
                    }
            }
        }
        if constexpr (!VER) __syncthreads();
        if constexpr (!VER)
        {   // kept entries: f^-1 and the global stores with every lane busy (a lane-at-a-time version cost a third of the kernel)
            const uint32_t nst = min(s_wr, (uint32_t)(QN / 2));
            for (uint32_t i = tid; i < nst; i += PC_THREADS) {
                const uint32_t r = s_q[2 * i], cnt = s_q[2 * i + 1];
                nge++;
                sumge += cnt;
                if (pbase + i < o.cap) {
                    o.keys[pbase + i] = mx.inv((p << mx.rbits) | (uint64_t)r);
                    o.counts[pbase + i] = cnt;
                }
            }
        }
        __syncthreads();
        if (tid == 0) {
            s_nq = 0;
            s_nkeep = 0;
            s_ndist = 0;
            s_wr = 0;
            if constexpr (LIST) s_nlist = 0;
        }
        __syncthreads();
#pragma unroll
        for (int u = 0; u < PC_PF; u++) cur[u] = nxt[u];
        beg = nbeg; end = nend; nbeg = nnbeg; nend = nnend;
        cnc = nnc; nnc = nnnc;
    }
    // ---- per-CTA totals ----
    if (o.histo) {
        h1 = spk_warp_sum_u32(h1);
        h2 = spk_warp_sum_u32(h2);
        if (lane == 0) {
            if (h1) atomicAdd(&s_hist[1], h1);
            if (h2) atomicAdd(&s_hist[2], h2);
        }
        __syncthreads();
        if ((uint32_t)tid < o.histo_len && s_hist[tid])
            atomicAdd((unsigned long long*)&o.histo[tid], (unsigned long long)s_hist[tid]);
    }
    uint64_t v[4] = {distinct, nge, sumge, sumall};
#pragma unroll
    for (int q = 0; q < 4; q++) {
        v[q] = spk_warp_sum_u64(v[q]);
        if (lane == 0) s_red[q][warp] = v[q];
    }
    n_fail = spk_warp_sum_u64(n_fail);
    if (lane == 0 && n_fail) atomicAdd((unsigned long long*)&o.stats[1], (unsigned long long)n_fail);
    __syncthreads();
    if (tid < 4) {
        uint64_t t = 0;
        for (int w = 0; w < PC_THREADS / 32; w++) t += s_red[tid][w];
        if (t) atomicAdd((unsigned long long*)&o.stats[4 + tid], (unsigned long long)t);
    }
}

template <bool ENT64>
int run_plan(const PcPlan& pl, const uint8_t* pk, const uint8_t* vl, uint32_t lower, char* ws, uint64_t* d_keys,
             uint32_t* d_counts, uint64_t cap, uint64_t* d_stats, uint64_t* d_histo, uint32_t histo_len,
             uint32_t* d_pindex, cudaStream_t st) {
    void* buf = ws + pl.off_buf;
    uint64_t* out_cursor = (uint64_t*)(ws + pl.off_out);
    SPK_CUDA(cudaMemsetAsync(out_cursor, 0, 256, st));
    if (pl.n_tiles == 0) return SPK_OK;
    const int sms = spk_num_sms();
    if (pl.v3 && !ENT64) {
        uint32_t* buf1 = (uint32_t*)(ws + pl.off_buf1);
        uint32_t* descT = (uint32_t*)(ws + pl.off_desc);
        uint16_t* off2 = (uint16_t*)(ws + pl.off_off2);
        V3Chunk* chunks = (V3Chunk*)(ws + pl.off_chunks);
        uint32_t* btot = (uint32_t*)(ws + pl.off_btot);
        uint32_t* cb = (uint32_t*)(ws + pl.off_cb);
        const int nb1 = 1 << pl.b1, nb2 = 1 << pl.b2;
        SPK_CUDA(cudaMemsetAsync(btot, 0, (size_t)nb1 * 4, st));
        constexpr size_t TILE_BYTES = 2 * ((SC_TILES * SPK_TILE_PACKED_BYTES + SPK_HALO_PACKED_BYTES) +
                                           (SC_TILES * SPK_TILE_VALID_BYTES + SPK_HALO_VALID_BYTES)) + 16;
        const size_t smem1 = (size_t)V3_ST1 * 4 + TILE_BYTES + (size_t)V3_THREADS * 8;
        const size_t smem2 = (size_t)V3_CH2 * 8 + (size_t)V3_THREADS * 16;
        SPK_CUDA(cudaFuncSetAttribute(k_v3_l1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
        SPK_CUDA(cudaFuncSetAttribute(k_v3_l2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
        k_v3_l1<<<(unsigned)min((uint32_t)sms, pl.n_st), V3_THREADS, smem1, st>>>(pk, vl, pl.n_tiles, pl.k, pl.mx, pl.b1, buf1,
                                                                                  descT, pl.n_st, btot, d_stats);
        SPK_LAUNCH_CHECK();
        k_v3_plan<<<1, 1024, 0, st>>>(btot, nb1, cb);
        SPK_LAUNCH_CHECK();
        k_v3_chunks<<<nb1, 1024, 0, st>>>(descT, pl.n_st, btot, cb, chunks);
        SPK_LAUNCH_CHECK();
        k_v3_l2<<<(unsigned)min((uint32_t)sms, 2 * pl.n_st + (uint32_t)nb1 + 1u), V3_THREADS, smem2, st>>>(
            buf1, descT, pl.n_st, chunks, cb + nb1, pl.b2, pl.mx.rbits, (uint32_t*)buf, off2);
        SPK_LAUNCH_CHECK();
        // "versioned": the no-clear / no-sweep table variant.  Measured slower on B200 (7.0 vs 5.8 ms on a 676-Mb
        // chromosome: its 64-bit shared-memory loads + CAS cost more than the sweep they save), kept for comparison.
        const char* ev = getenv("SPK_PCOUNT_TABLE");
        const bool ver = (ev && ev[0] == 'v');
        const bool list = !ver && !(ev && ev[0] == 's') && !d_histo;       // "sweep": the sweep kernel for every call
        uint32_t retry_cap = ver ? PC_RETRY_V : (list ? PC_RETRY_L : PC_RETRY);
        if (const char* e = getenv("SPK_PCOUNT_RETRY_CAP")) {       // test hook: exercise the queue-overflow paths
            const long v = atol(e);
            if (v >= 0 && v < (long)retry_cap) retry_cap = (uint32_t)v;
        }
        CountOut o{d_keys, d_counts, cap, out_cursor, d_stats, d_histo, histo_len, lower, d_pindex};
        GatherIn gi{cb, off2, pl.b2};
        const unsigned cgrid3 = (unsigned)min((uint64_t)sms * 3, pl.P);
        if (ver) {
            const size_t smv = (size_t)PC_SLOTS * 8 + PC_RETRY_V * 4 + PC_KEEP * 2;
            SPK_CUDA(cudaFuncSetAttribute(k_part_count32<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smv));
            k_part_count32<true, true><<<cgrid3, PC_THREADS, smv, st>>>((const uint32_t*)buf, nullptr, pl.P, pl.mx, o,
                                                                       retry_cap, gi);
        } else if (list) {
            const size_t sml = (size_t)PC_SLOTS * 8 + PC_RETRY_L * 4 + PC_KEEP_L * 2;
            SPK_CUDA(cudaFuncSetAttribute(k_part_count32<true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sml));
            k_part_count32<true, false, true><<<cgrid3, PC_THREADS, sml, st>>>((const uint32_t*)buf, nullptr, pl.P, pl.mx, o,
                                                                              retry_cap, gi);
        } else {
            SPK_CUDA(cudaFuncSetAttribute(k_part_count32<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          PC_SLOTS * 8 + PC_RETRY * 4));
            k_part_count32<true, false><<<cgrid3, PC_THREADS, PC_SLOTS * 8 + PC_RETRY * 4, st>>>(
                (const uint32_t*)buf, nullptr, pl.P, pl.mx, o, retry_cap, gi);
        }
        SPK_LAUNCH_CHECK();
        return SPK_OK;
    }
    uint32_t* psize = (uint32_t*)(ws + pl.off_psize);
    uint32_t* pstart = (uint32_t*)(ws + pl.off_pstart);
    uint32_t* cursor = (uint32_t*)(ws + pl.off_cursor);
    uint32_t* segs = (uint32_t*)(ws + pl.off_segs);
    SPK_CUDA(cudaMemsetAsync(psize, 0, (pl.P + 1) * 4, st));
    const unsigned pass_grid = (unsigned)min((uint64_t)sms * 4, pl.n_tiles);
    const uint64_t nseg = (pl.P + 1023) / 1024;
    if (pl.two_level && !ENT64) {
        uint32_t* buf1 = (uint32_t*)(ws + pl.off_buf1);
        uint32_t* cur1 = (uint32_t*)(ws + pl.off_cur1);
        uint32_t* ustart = (uint32_t*)(ws + pl.off_ustart);
        uint32_t* bsize = ustart;                      // bucket sizes live in ustart until k_scatter_prepare
        const int nb1 = 1 << pl.b1;
        SPK_CUDA(cudaMemsetAsync(bsize, 0, (size_t)nb1 * 4, st));
        int red_split = 6;                             // of 16: share of the partition-histogram REDs issued by k_hist1
        if (const char* e = getenv("SPK_PCOUNT_SPLIT")) red_split = atoi(e);
        k_hist1<<<pass_grid, SPK_TILE_THREADS, 0, st>>>(pk, vl, pl.n_tiles, pl.k, pl.mx, pl.b1, bsize, d_stats, psize,
                                                        red_split);
        SPK_LAUNCH_CHECK();
        k_bucket_scan<<<1, 1024, 0, st>>>(bsize, nb1, cur1);
        SPK_LAUNCH_CHECK();
        const size_t tile_bytes = ((SC_TILES * SPK_TILE_PACKED_BYTES + SPK_HALO_PACKED_BYTES) +
                                   (SC_TILES * SPK_TILE_VALID_BYTES + SPK_HALO_VALID_BYTES) + 8 + 16 + 15) / 16 * 16;
        const size_t smem1 = tile_bytes + scatter_smem_bytes(1 << pl.b1);
        const size_t smem2 = scatter_smem_bytes(1 << pl.b2);
        // (the attribute is per device: set it on every call, not once per process)
        SPK_CUDA(cudaFuncSetAttribute(k_scatter_l1, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
        SPK_CUDA(cudaFuncSetAttribute(k_scatter_l2, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
        const uint64_t n_super = (pl.n_tiles + SC_TILES - 1) / SC_TILES;
        k_scatter_l1<<<(unsigned)min((uint64_t)sms * 2, n_super), SPK_TILE_THREADS, smem1, st>>>(
            pk, vl, pl.n_tiles, pl.k, pl.mx, pl.b1, cur1, buf1, psize, red_split);
        SPK_LAUNCH_CHECK();
        k_scan_seg_totals<<<(unsigned)nseg, 1024, 0, st>>>(psize, pl.P, segs);
        SPK_LAUNCH_CHECK();
        k_scan_segs<<<1, 1024, 0, st>>>(segs, nseg);
        SPK_LAUNCH_CHECK();
        k_scan_apply<<<(unsigned)nseg, 1024, 0, st>>>(psize, pl.P, segs, pstart, cursor);
        SPK_LAUNCH_CHECK();
        k_scatter_prepare<<<1, 1024, 0, st>>>(pstart, pl.b1, pl.b2, cur1, ustart);
        SPK_LAUNCH_CHECK();
        k_scatter_l2<<<(unsigned)(sms * 2), SC2_THREADS, smem2, st>>>(buf1, pstart, ustart, pl.b1, pl.b2,
                                                                           pl.mx.rbits, cursor, (uint32_t*)buf);
        SPK_LAUNCH_CHECK();
    } else {
        k_part_pass<false, ENT64><<<pass_grid, SPK_TILE_THREADS, 0, st>>>(pk, vl, pl.n_tiles, pl.k, pl.mx, psize,
                                                                          nullptr, nullptr, d_stats);
        SPK_LAUNCH_CHECK();
        k_scan_seg_totals<<<(unsigned)nseg, 1024, 0, st>>>(psize, pl.P, segs);
        SPK_LAUNCH_CHECK();
        k_scan_segs<<<1, 1024, 0, st>>>(segs, nseg);
        SPK_LAUNCH_CHECK();
        k_scan_apply<<<(unsigned)nseg, 1024, 0, st>>>(psize, pl.P, segs, pstart, cursor);
        SPK_LAUNCH_CHECK();
        k_part_pass<true, ENT64><<<pass_grid, SPK_TILE_THREADS, 0, st>>>(pk, vl, pl.n_tiles, pl.k, pl.mx, psize,
                                                                         cursor, buf, d_stats);
        SPK_LAUNCH_CHECK();
    }
    const size_t smem = (size_t)PC_SLOTS * (ENT64 ? 12 : 8) + (size_t)PC_LIST * 2;
    SPK_CUDA(cudaFuncSetAttribute(k_part_count<ENT64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CountOut o{d_keys, d_counts, cap, out_cursor, d_stats, d_histo, histo_len, lower, d_pindex};
    const unsigned cgrid = (unsigned)min((uint64_t)sms * (ENT64 ? 2 : 3), pl.P);
    const char* v1 = getenv("SPK_PCOUNT_DUMP");        // "list": the 64-bit-slot kernel with the occupied-slot list
    if (!ENT64 && pl.mx.rbits <= 31 && lower >= 1 && !(v1 && v1[0] == 'l')) {
        SPK_CUDA(cudaFuncSetAttribute(k_part_count32<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      PC_SLOTS * 8 + PC_RETRY * 4));
        uint32_t retry_cap = PC_RETRY;
        if (const char* e = getenv("SPK_PCOUNT_RETRY_CAP")) {       // test hook: exercise the queue-overflow paths
            const long v = atol(e);
            if (v >= 0 && v < PC_RETRY) retry_cap = (uint32_t)v;
        }
        k_part_count32<false, false><<<cgrid, PC_THREADS, PC_SLOTS * 8 + PC_RETRY * 4, st>>>((const uint32_t*)buf, pstart, pl.P,
                                                                                     pl.mx, o, retry_cap, GatherIn{});
    } else {
        k_part_count<ENT64><<<cgrid, PC_THREADS, smem, st>>>(buf, pstart, pl.P, pl.mx, o);
    }
    SPK_LAUNCH_CHECK();
    return SPK_OK;
}

}  // namespace

extern "C" int spk_pcount_pbits(uint64_t n_bases, int k) {
    if (k < 1 || k > 32 || n_bases >= 0xffffffffull) return -1;
    return auto_pbits(n_bases, k);
}

extern "C" size_t spk_pcount_workspace_bytes_ex(uint64_t n_bases, int k, int pbits) {
    PcPlan pl;
    if (make_plan(n_bases, k, pbits, &pl) != SPK_OK) return 0;
    return pl.total;
}

extern "C" size_t spk_pcount_workspace_bytes(uint64_t n_bases, int k) {
    return spk_pcount_workspace_bytes_ex(n_bases, k, 0);
}

extern "C" int spk_pcount_canonical_ex(const uint32_t* d_packed, const uint32_t* d_valid, uint64_t n_bases, int k,
                                       uint32_t lower_count, void* d_ws, size_t ws_bytes, uint64_t* d_keys,
                                       uint32_t* d_counts, uint64_t cap, uint64_t* d_stats, uint64_t* d_histo,
                                       uint32_t histo_len, int pbits, uint32_t* d_pindex, void* stream) {
    SPK_CHECK_ARG(d_packed && d_valid && d_ws && d_stats, "null pointer");
    SPK_CHECK_ARG(cap == 0 || (d_keys && d_counts), "null output");
    SPK_CHECK_ARG(k >= 1 && k <= 32, "k must be in [1, 32]");
    SPK_CHECK_ARG(n_bases < 0xffffffffull, "partitioned counter handles < 2^32 bases per chromosome");
    SPK_CHECK_ARG(!d_histo || histo_len >= 2, "histo_len must be >= 2");
    SPK_CHECK_ARG(((uintptr_t)d_packed & 15) == 0 && ((uintptr_t)d_valid & 15) == 0 && ((uintptr_t)d_ws & 255) == 0,
                  "buffers must be aligned (16 B sequence, 256 B workspace)");
    PcPlan pl;
    if (make_plan(n_bases, k, pbits, &pl) != SPK_OK) {
        spk_set_error("spk_pcount_canonical: cannot plan (pbits=%d must be 0 or >= spk_pcount_pbits and <= min(2k, %d))",
                      pbits, PC_MAX_PBITS);
        return SPK_EINVAL;
    }
    if (ws_bytes < pl.total) {
        spk_set_error("spk_pcount_canonical: workspace %zu < %zu bytes", ws_bytes, pl.total);
        return SPK_ECAP;
    }
    cudaStream_t st = (cudaStream_t)stream;
    SPK_CUDA(cudaMemsetAsync(d_stats, 0, 8 * sizeof(uint64_t), st));
    if (n_bases < (uint64_t)k) {
        if (d_pindex) SPK_CUDA(cudaMemsetAsync(d_pindex, 0, (size_t)pl.P * 8, st));
        return SPK_OK;
    }
    if (pl.ent64)
        return run_plan<true>(pl, (const uint8_t*)d_packed, (const uint8_t*)d_valid, lower_count, (char*)d_ws, d_keys,
                              d_counts, cap, d_stats, d_histo, histo_len, d_pindex, st);
    return run_plan<false>(pl, (const uint8_t*)d_packed, (const uint8_t*)d_valid, lower_count, (char*)d_ws, d_keys,
                           d_counts, cap, d_stats, d_histo, histo_len, d_pindex, st);
}

extern "C" int spk_pcount_canonical(const uint32_t* d_packed, const uint32_t* d_valid, uint64_t n_bases, int k,
                                    uint32_t lower_count, void* d_ws, size_t ws_bytes, uint64_t* d_keys,
                                    uint32_t* d_counts, uint64_t cap, uint64_t* d_stats, uint64_t* d_histo,
                                    uint32_t histo_len, void* stream) {
    return spk_pcount_canonical_ex(d_packed, d_valid, n_bases, k, lower_count, d_ws, ws_bytes, d_keys, d_counts, cap,
                                   d_stats, d_histo, histo_len, 0, nullptr, stream);
}
