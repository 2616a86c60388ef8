// K2+K3 (partitioned): canonical k-mer counting with every random access on-chip.
//
// Measured on B200 (tools/atomics_bench.cu): a dependent "load slot, then atomic on it" sustains only
// ~1.5e10/s when the table lives in HBM (the v1 kernel, spk_count.cu, sits exactly there) and ~6e10/s
// even when the table is L2-resident, while shared-memory atomics run > 3e11/s.  So the chromosome is
// split into P = 2^pbits hash partitions small enough that one partition's table fits in the shared
// memory of a CTA; HBM only sees streaming traffic:
//
//   phase 0  k_part_pass<hist>     recompute k-mers tile by tile (TMA-staged), RED partition sizes
//            k_scan_*              exclusive scan of the P sizes -> partition extents, cursors
//   phase 1  k_part_pass<scatter>  same traversal, append the 32-bit remainder r to its partition
//   phase 2  k_part_count          one CTA per partition: stream r, insert into an smem table
//                                  (CAS on new keys, 32-bit add on hits), then scan the table:
//                                  stats, histogram, dump of count >= L
//
// DRAM traffic per k-mer: 2 x 0.375 B (sequence, read twice) + 4 B write + 4 B read.
// Partitioning uses a bijective mixer f on the 2k-bit canonical word: partition = top pbits of f(u),
// remainder r = the rest; the dump applies f^-1, so keys are exact and the result is bit-identical to
// the v1 kernel / the jellyfish semantics (only the dump ORDER differs, which is arbitrary anyway).
#include <stdlib.h>
#include "spk_common.cuh"
#include "spk_tile.cuh"

namespace {

constexpr int PC_SLOTS = 8192;          // smem table slots per CTA (64 KB as u64, 96 KB for wide keys)
constexpr int PC_TARGET = 3072;         // planned mean entries per partition (all distinct -> load 0.375)
constexpr int PC_MAX_PBITS = 22;
constexpr int PC_THREADS = 256;
constexpr uint64_t PC_C1 = 0xff51afd7ed558ccdULL;
constexpr uint64_t PC_C2 = 0xc4ceb9fe1a85ec53ULL;

constexpr uint64_t inv64(uint64_t a) {  // multiplicative inverse of odd a modulo 2^64 (Newton)
    uint64_t x = a;
    for (int i = 0; i < 6; i++) x *= 2 - a * x;
    return x;
}
constexpr uint64_t PC_C1_INV = inv64(PC_C1);
constexpr uint64_t PC_C2_INV = inv64(PC_C2);
static_assert(PC_C1 * PC_C1_INV == 1ull && PC_C2 * PC_C2_INV == 1ull, "inverse constants");

struct Mixer {
    uint64_t mask;  // 2k low bits
    int s;          // xorshift distance k = (2k)/2: x ^= x >> s is an involution on 2k-bit words
    int rbits;      // remainder bits = 2k - pbits
    __host__ __device__ uint64_t fwd(uint64_t u) const {
        uint64_t x = u;
        x ^= x >> s;
        x = (x * PC_C1) & mask;
        x ^= x >> s;
        x = (x * PC_C2) & mask;
        x ^= x >> s;
        return x;
    }
    __host__ __device__ uint64_t inv(uint64_t x) const {
        x ^= x >> s;
        x = (x * PC_C2_INV) & mask;
        x ^= x >> s;
        x = (x * PC_C1_INV) & mask;
        x ^= x >> s;
        return x;
    }
};

struct PcPlan {
    int k, pbits, ent64;
    uint64_t P;
    Mixer mx;
    uint64_t n_tiles;
    size_t off_buf, off_psize, off_pstart, off_cursor, off_segs, off_out, total;
};

inline size_t al256(size_t x) { return (x + 255) / 256 * 256; }

int make_plan(uint64_t n_bases, int k, PcPlan* pl) {
    if (k < 1 || k > 32) return SPK_EINVAL;
    if (n_bases >= 0xffffffffull) return SPK_EINVAL;  // 32-bit partition offsets
    pl->k = k;
    int pbits = 2;
    while (pbits < PC_MAX_PBITS && (n_bases >> pbits) > (uint64_t)PC_TARGET) pbits++;
    if (pbits > 2 * k) pbits = 2 * k;  // tiny k: at most 4^k distinct words
    pl->pbits = pbits;
    pl->P = 1ull << pbits;
    pl->ent64 = (2 * k - pbits > 32) ? 1 : 0;
    pl->mx.mask = (k == 32) ? ~0ull : ((1ull << (2 * k)) - 1);
    pl->mx.s = k;
    pl->mx.rbits = 2 * k - pbits;
    pl->n_tiles = (n_bases + SPK_TILE_BASES - 1) / SPK_TILE_BASES;
    size_t off = 0;
    pl->off_buf = off;
    off += al256((size_t)(n_bases + 64) * (pl->ent64 ? 8 : 4));
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
};

// ENT32: slot u64 = (r << 32) | count, empty = 0 (count >= 1 once occupied).
// ENT64: key u64 (empty = ~0) and count u32 in separate arrays.
template <bool ENT64>
__global__ void __launch_bounds__(PC_THREADS, 3)
k_part_count(const void* __restrict__ buf, const uint32_t* __restrict__ pstart, uint64_t P, Mixer mx,
             CountOut o) {
    extern __shared__ __align__(16) uint8_t s_raw[];
    uint64_t* s_key = (uint64_t*)s_raw;
    uint32_t* s_cnt = (uint32_t*)(s_raw + (size_t)PC_SLOTS * 8);  // ENT64 only
    __shared__ uint32_t s_hist[256];
    __shared__ uint64_t s_red[4][PC_THREADS / 32];
    const int tid = threadIdx.x;
    uint64_t distinct = 0, nge = 0, sumge = 0, sumall = 0, n_fail = 0;
    uint32_t h1 = 0, h2 = 0;
    if (o.histo) s_hist[tid] = 0;
    const uint64_t EMPTY = ENT64 ? SPK_EMPTY_KEY : 0ull;

    for (uint64_t p = blockIdx.x; p < P; p += gridDim.x) {
        const uint32_t beg = pstart[p], end = pstart[p + 1];
        const uint32_t n_p = end - beg;
        if (n_p == 0) continue;
        // table size: power of two >= 2 * entries, at most PC_SLOTS
        uint32_t tsz = 64;
        while (tsz < PC_SLOTS && tsz < 2 * n_p) tsz <<= 1;
        const uint32_t tmask = tsz - 1;
        __syncthreads();  // previous partition's scan is finished
        for (uint32_t i = tid; i < tsz; i += PC_THREADS) {
            s_key[i] = EMPTY;
            if (ENT64) s_cnt[i] = 0;
        }
        __syncthreads();
        // ---- insert ----
        for (uint32_t i = beg + tid; i < end; i += PC_THREADS) {
            const uint64_t r = ENT64 ? __ldcs((const uint64_t*)buf + i) : (uint64_t)__ldcs((const uint32_t*)buf + i);
            uint32_t s = (ENT64 ? (uint32_t)spk_hash64(r) : fmix32((uint32_t)r)) & tmask;
            bool done = false;
            for (uint32_t probes = 0; probes < tsz; probes++) {
                uint64_t c = s_key[s];
                if (c == EMPTY) {
                    const uint64_t fresh = ENT64 ? r : ((r << 32) | 1ull);
                    const uint64_t old = atomicCAS((unsigned long long*)&s_key[s], (unsigned long long)EMPTY,
                                                   (unsigned long long)fresh);
                    if (old == EMPTY) {
                        if (ENT64) atomicAdd(&s_cnt[s], 1u);
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
                s = (s + 1) & tmask;
            }
            if (!done) n_fail++;
        }
        __syncthreads();
        // ---- scan: stats, histogram, dump ----
        for (uint32_t base = 0; base < tsz; base += PC_THREADS) {
            const uint32_t i = base + tid;
            const uint64_t v = (i < tsz) ? s_key[i] : EMPTY;   // tables smaller than the CTA exist
            const bool occ = v != EMPTY;
            uint64_t r = 0, cnt = 0;
            if (occ) {
                r = ENT64 ? v : (v >> 32);
                cnt = ENT64 ? (uint64_t)s_cnt[i] : (v & 0xffffffffull);
                distinct++;
                sumall += cnt;
                if (o.histo) {
                    const uint64_t b = cnt < (uint64_t)(o.histo_len - 1) ? cnt : (uint64_t)(o.histo_len - 1);
                    if (b == 1) h1++;
                    else if (b == 2) h2++;
                    else if (b < 256) atomicAdd(&s_hist[b], 1u);
                    else atomicAdd((unsigned long long*)&o.histo[b], 1ull);
                }
            }
            const bool keep = occ && cnt >= o.lower;
            if (keep) {
                nge++;
                sumge += cnt;
            }
            const uint32_t ballot = __ballot_sync(0xffffffffu, keep);
            if (ballot) {
                const int lane = tid & 31;
                uint64_t wbase = 0;
                if (lane == 0) wbase = atomicAdd((unsigned long long*)o.cursor, (unsigned long long)__popc(ballot));
                wbase = __shfl_sync(0xffffffffu, wbase, 0);
                if (keep) {
                    const uint64_t at = wbase + __popc(ballot & ((1u << lane) - 1));
                    if (at < o.cap) {
                        o.keys[at] = mx.inv((p << mx.rbits) | r);
                        o.counts[at] = (uint32_t)cnt;
                    }
                }
            }
        }
    }
    // ---- per-CTA totals ----
    __syncthreads();
    if (o.histo) {
        h1 = spk_warp_sum_u32(h1);
        h2 = spk_warp_sum_u32(h2);
        if ((tid & 31) == 0) {
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
        if ((tid & 31) == 0) s_red[q][tid >> 5] = v[q];
    }
    n_fail = spk_warp_sum_u64(n_fail);
    if ((tid & 31) == 0 && n_fail) atomicAdd((unsigned long long*)&o.stats[1], (unsigned long long)n_fail);
    __syncthreads();
    if (tid < 4) {
        uint64_t s = 0;
        for (int w = 0; w < PC_THREADS / 32; w++) s += s_red[tid][w];
        if (s) atomicAdd((unsigned long long*)&o.stats[4 + tid], (unsigned long long)s);
    }
}

template <bool ENT64>
int run_plan(const PcPlan& pl, const uint8_t* pk, const uint8_t* vl, uint32_t lower, char* ws, uint64_t* d_keys,
             uint32_t* d_counts, uint64_t cap, uint64_t* d_stats, uint64_t* d_histo, uint32_t histo_len,
             cudaStream_t st) {
    void* buf = ws + pl.off_buf;
    uint32_t* psize = (uint32_t*)(ws + pl.off_psize);
    uint32_t* pstart = (uint32_t*)(ws + pl.off_pstart);
    uint32_t* cursor = (uint32_t*)(ws + pl.off_cursor);
    uint32_t* segs = (uint32_t*)(ws + pl.off_segs);
    uint64_t* out_cursor = (uint64_t*)(ws + pl.off_out);
    SPK_CUDA(cudaMemsetAsync(psize, 0, (pl.P + 1) * 4, st));
    SPK_CUDA(cudaMemsetAsync(out_cursor, 0, 256, st));
    if (pl.n_tiles == 0) return SPK_OK;
    const int sms = spk_num_sms();
    const unsigned pass_grid = (unsigned)min((uint64_t)sms * 4, pl.n_tiles);
    k_part_pass<false, ENT64><<<pass_grid, SPK_TILE_THREADS, 0, st>>>(pk, vl, pl.n_tiles, pl.k, pl.mx, psize,
                                                                      nullptr, nullptr, d_stats);
    SPK_LAUNCH_CHECK();
    const uint64_t nseg = (pl.P + 1023) / 1024;
    k_scan_seg_totals<<<(unsigned)nseg, 1024, 0, st>>>(psize, pl.P, segs);
    SPK_LAUNCH_CHECK();
    k_scan_segs<<<1, 1024, 0, st>>>(segs, nseg);
    SPK_LAUNCH_CHECK();
    k_scan_apply<<<(unsigned)nseg, 1024, 0, st>>>(psize, pl.P, segs, pstart, cursor);
    SPK_LAUNCH_CHECK();
    k_part_pass<true, ENT64><<<pass_grid, SPK_TILE_THREADS, 0, st>>>(pk, vl, pl.n_tiles, pl.k, pl.mx, psize,
                                                                     cursor, buf, d_stats);
    SPK_LAUNCH_CHECK();
    const size_t smem = (size_t)PC_SLOTS * (ENT64 ? 12 : 8);
    static bool attr_set[2] = {false, false};
    if (!attr_set[ENT64 ? 1 : 0]) {
        SPK_CUDA(cudaFuncSetAttribute(k_part_count<ENT64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set[ENT64 ? 1 : 0] = true;
    }
    CountOut o{d_keys, d_counts, cap, out_cursor, d_stats, d_histo, histo_len, lower};
    const unsigned cgrid = (unsigned)min((uint64_t)sms * (ENT64 ? 2 : 3), pl.P);
    k_part_count<ENT64><<<cgrid, PC_THREADS, smem, st>>>(buf, pstart, pl.P, pl.mx, o);
    SPK_LAUNCH_CHECK();
    return SPK_OK;
}

}  // namespace

extern "C" size_t spk_pcount_workspace_bytes(uint64_t n_bases, int k) {
    PcPlan pl;
    if (make_plan(n_bases, k, &pl) != SPK_OK) return 0;
    return pl.total;
}

extern "C" int spk_pcount_canonical(const uint32_t* d_packed, const uint32_t* d_valid, uint64_t n_bases, int k,
                                    uint32_t lower_count, void* d_ws, size_t ws_bytes, uint64_t* d_keys,
                                    uint32_t* d_counts, uint64_t cap, uint64_t* d_stats, uint64_t* d_histo,
                                    uint32_t histo_len, void* stream) {
    SPK_CHECK_ARG(d_packed && d_valid && d_ws && d_stats, "null pointer");
    SPK_CHECK_ARG(cap == 0 || (d_keys && d_counts), "null output");
    SPK_CHECK_ARG(k >= 1 && k <= 32, "k must be in [1, 32]");
    SPK_CHECK_ARG(n_bases < 0xffffffffull, "partitioned counter handles < 2^32 bases per chromosome");
    SPK_CHECK_ARG(!d_histo || histo_len >= 2, "histo_len must be >= 2");
    SPK_CHECK_ARG(((uintptr_t)d_packed & 15) == 0 && ((uintptr_t)d_valid & 15) == 0 && ((uintptr_t)d_ws & 255) == 0,
                  "buffers must be aligned (16 B sequence, 256 B workspace)");
    PcPlan pl;
    if (make_plan(n_bases, k, &pl) != SPK_OK) {
        spk_set_error("spk_pcount_canonical: cannot plan");
        return SPK_EINVAL;
    }
    if (ws_bytes < pl.total) {
        spk_set_error("spk_pcount_canonical: workspace %zu < %zu bytes", ws_bytes, pl.total);
        return SPK_ECAP;
    }
    cudaStream_t st = (cudaStream_t)stream;
    SPK_CUDA(cudaMemsetAsync(d_stats, 0, 8 * sizeof(uint64_t), st));
    if (n_bases < (uint64_t)k) return SPK_OK;
    if (pl.ent64)
        return run_plan<true>(pl, (const uint8_t*)d_packed, (const uint8_t*)d_valid, lower_count, (char*)d_ws, d_keys,
                              d_counts, cap, d_stats, d_histo, histo_len, st);
    return run_plan<false>(pl, (const uint8_t*)d_packed, (const uint8_t*)d_valid, lower_count, (char*)d_ws, d_keys,
                           d_counts, cap, d_stats, d_histo, histo_len, st);
}
