// K2 (v2): partitioned canonical k-mer counting — the table never leaves L2.
//
// The v1 kernel (spk_count.cu) inserts straight into a multi-GB table: every k-mer is a random DRAM
// read-modify-write and the chip saturates at ~1e10 inserts/s (GUPS-bound; ncu: 111 B DRAM read per
// k-mer, 27 % L2 hit, 29 % DRAM throughput).  Here the chromosome is first split into P hash
// partitions, streamed to HBM as 4-byte remainders, and each partition is then counted in an
// open-addressed table small enough (<= ~16 MB) to stay resident in the 126 MB L2:
//
//   phase 0  k_part_hist     per 64-tile chunk: histogram of partition ids (smem atomics)
//            k_part_offsets  column scans -> exact write offset of every (chunk, partition) run
//   phase 1  k_part_scatter  recompute the k-mers, append remainder r to its partition (smem cursors)
//   phase 2  k_part_count    G partitions at a time: coalesced read of r, probe + CAS/RED in L2
//            k_part_extract  scan the G tables (L2 hits): stats, dump count >= L, clear for reuse
//
// DRAM traffic per k-mer: 0.375 B (sequence, twice) + 4 B write + 4 B read, all streaming.
// Partitioning uses a bijective mixer f on the 2k-bit canonical word: partition = top bits of f(u),
// remainder r = low bits; the dump inverts f, so keys are exact.  Semantics identical to v1 / jellyfish.
#include <stdlib.h>
#include "spk_common.cuh"
#include "spk_tile.cuh"

namespace {

constexpr int PC_CHUNK_TILES = 64;
constexpr int PC_MAX_P = 1024;
constexpr int PC_BATCH = 8;
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
    int s;          // xorshift distance >= ceil(2k/2): x ^= x >> s is an involution on 2k-bit words
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
    int k, pbits, P, ent64;       // ent64: remainders do not fit 32 bits
    Mixer mx;
    uint64_t n_tiles, n_chunks;
    uint64_t T;                   // table slots per partition
    int G;                        // partitions counted concurrently
    // workspace offsets (bytes)
    size_t off_buf, off_hist, off_psize, off_pstart, off_tables, off_cursor, total;
};

inline size_t al256(size_t x) { return (x + 255) / 256 * 256; }

int env_int(const char* name, int dflt) {
    const char* e = getenv(name);
    return (e && *e) ? atoi(e) : dflt;
}

int make_plan(uint64_t n_bases, int k, PcPlan* pl) {
    if (k < 1 || k > 32) return SPK_EINVAL;
    if (n_bases >= 0xffffffffull) return SPK_EINVAL;  // 32-bit partition offsets
    pl->k = k;
    const double table_mb = (double)env_int("SPK_PCOUNT_TABLE_MB", 16);
    // smallest power of two P with (n/P)/0.7 * 8 B <= table_mb
    int pbits = 4;
    while (pbits < 10 && ((double)n_bases / (double)(1u << pbits)) / 0.7 * 8.0 > table_mb * 1e6) pbits++;
    if (2 * k - pbits > 32 && 2 * k - 10 <= 32) pbits = 2 * k - 32;  // keep remainders in 32 bits if possible
    if (pbits > 2 * k) pbits = 2 * k;                                 // tiny k: at most 4^k partitions
    pl->pbits = pbits;
    pl->P = 1 << pbits;
    pl->ent64 = (2 * k - pbits > 32) ? 1 : 0;
    pl->mx.mask = (k == 32) ? ~0ull : ((1ull << (2 * k)) - 1);
    pl->mx.s = k;  // = 2k/2
    pl->mx.rbits = 2 * k - pbits;
    pl->n_tiles = (n_bases + SPK_TILE_BASES - 1) / SPK_TILE_BASES;
    pl->n_chunks = (pl->n_tiles + PC_CHUNK_TILES - 1) / PC_CHUNK_TILES;
    // distinct keys spread uniformly over partitions whatever their multiplicities
    uint64_t per_part = n_bases / pl->P + 1;
    if (pl->mx.rbits < 40 && (1ull << pl->mx.rbits) < per_part) per_part = 1ull << pl->mx.rbits;
    pl->T = (uint64_t)((double)per_part / 0.7) + 4096;
    pl->G = env_int("SPK_PCOUNT_G", 4);
    if (pl->G > pl->P) pl->G = pl->P;
    if (pl->G < 1) pl->G = 1;
    size_t off = 0;
    pl->off_buf = off;
    off += al256((size_t)(n_bases + 64) * (pl->ent64 ? 8 : 4));
    pl->off_hist = off;
    off += al256((size_t)(pl->n_chunks + 1) * pl->P * 4);
    pl->off_psize = off;
    off += al256((size_t)pl->P * 4);
    pl->off_pstart = off;
    off += al256((size_t)(pl->P + 1) * 4);
    pl->off_tables = off;
    off += al256((size_t)pl->G * pl->T * (pl->ent64 ? 12 : 8));
    pl->off_cursor = off;
    off += 256;
    pl->total = off;
    return SPK_OK;
}

// ---- phases 0 and 1 share the traversal ---------------------------------------------------------------
template <bool SCATTER, bool ENT64>
__global__ void __launch_bounds__(SPK_TILE_THREADS, 4)
k_part_pass(const uint8_t* __restrict__ packed, const uint8_t* __restrict__ valid, uint64_t n_tiles,
            int k, Mixer mx, int P, uint32_t* __restrict__ chunk_hist, const uint32_t* __restrict__ pstart,
            void* __restrict__ buf, uint64_t* __restrict__ stats) {
    __shared__ SpkTileSmem sm;
    __shared__ uint32_t s_bins[PC_MAX_P];
    const int tid = threadIdx.x;
    const SpkKmerParams kp = spk_kmer_params(k);
    const uint64_t chunk = blockIdx.x;
    const uint64_t t0 = chunk * PC_CHUNK_TILES;
    const uint64_t t1 = min(t0 + PC_CHUNK_TILES, n_tiles);
    spk_tile_init(sm);
    for (int p = tid; p < P; p += SPK_TILE_THREADS)
        s_bins[p] = SCATTER ? (pstart[p] + chunk_hist[chunk * P + p]) : 0u;
    uint64_t n_valid = 0;
    if (tid == 0 && t0 < t1) spk_tile_issue(sm, packed, valid, t0, 0);
    uint32_t it = 0;
    for (uint64_t tile = t0; tile < t1; tile++, it++) {
        const int b = it & 1;
        __syncthreads();  // buffer b^1 free; s_bins initialised (first iteration)
        if (tid == 0 && tile + 1 < t1) spk_tile_issue(sm, packed, valid, tile + 1, b ^ 1);
        spk_mbar_wait(&sm.bar[b], (it >> 1) & 1);
        uint64_t key[SPK_KMERS_PER_THREAD];
        uint32_t okmask;
        spk_tile_kmers(sm, b, kp, key, okmask);
        n_valid += __popc(okmask);
#pragma unroll
        for (int j = 0; j < SPK_KMERS_PER_THREAD; j++) {
            if ((okmask >> j) & 1u) {
                const uint64_t h = mx.fwd(key[j]);
                const uint32_t p = (uint32_t)(h >> mx.rbits);
                const uint32_t pos = atomicAdd(&s_bins[p], 1u);
                if (SCATTER) {
                    const uint64_t r = h & ((1ull << mx.rbits) - 1);
                    if (ENT64) ((uint64_t*)buf)[pos] = r;
                    else ((uint32_t*)buf)[pos] = (uint32_t)r;
                }
            }
        }
    }
    __syncthreads();
    if (!SCATTER) {
        for (int p = tid; p < P; p += SPK_TILE_THREADS) chunk_hist[chunk * P + p] = s_bins[p];
        n_valid = spk_warp_sum_u64(n_valid);
        if ((tid & 31) == 0 && n_valid) atomicAdd((unsigned long long*)&stats[0], (unsigned long long)n_valid);
    }
}

// Column scans of chunk_hist [n_chunks x P]: block = 32 partitions x 32 segments of chunks.
// chunk_hist[c][p] becomes the exclusive prefix over chunks; psize[p] the column total.
__global__ void __launch_bounds__(1024)
k_part_colscan(uint32_t* __restrict__ chunk_hist, uint64_t n_chunks, int P, uint32_t* __restrict__ psize) {
    __shared__ uint32_t s_seg[32][33];
    const int lane = threadIdx.x & 31, seg = threadIdx.x >> 5;
    const int p = blockIdx.x * 32 + lane;
    const uint64_t per = (n_chunks + 31) / 32;
    const uint64_t c0 = min((uint64_t)seg * per, n_chunks), c1 = min(c0 + per, n_chunks);
    uint32_t sum = 0;
    if (p < P)
        for (uint64_t c = c0; c < c1; c++) sum += chunk_hist[c * P + p];
    s_seg[seg][lane] = sum;
    __syncthreads();
    uint32_t base = 0;
    for (int s2 = 0; s2 < seg; s2++) base += s_seg[s2][lane];
    if (p < P) {
        uint32_t run = base;
        for (uint64_t c = c0; c < c1; c++) {
            const uint32_t v = chunk_hist[c * P + p];
            chunk_hist[c * P + p] = run;
            run += v;
        }
        if (seg == 31) psize[p] = run;
    }
}

// exclusive scan of the P partition sizes (P <= 1024) -> pstart[0..P]
__global__ void __launch_bounds__(1024) k_part_starts(const uint32_t* __restrict__ psize, int P,
                                                       uint32_t* __restrict__ pstart) {
    __shared__ uint32_t s_warp[32];
    const uint32_t v = ((int)threadIdx.x < P) ? psize[threadIdx.x] : 0u;
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if ((threadIdx.x & 31) >= o) incl += t;
    }
    if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = incl;
    __syncthreads();
    uint32_t prefix = 0;
    for (int w = 0; w < (int)(threadIdx.x >> 5); w++) prefix += s_warp[w];
    if ((int)threadIdx.x < P) pstart[threadIdx.x] = prefix + incl - v;
    if ((int)threadIdx.x == P - 1) pstart[P] = prefix + incl;
}

// ---- phase 2 ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t fmix32(uint32_t h) {
    h ^= h >> 16;
    h *= 0x85ebca6bu;
    h ^= h >> 13;
    h *= 0xc2b2ae35u;
    h ^= h >> 16;
    return h;
}

template <bool ENT64>
__device__ __forceinline__ uint64_t slot_of_r(uint64_t r, uint64_t T) {
    if (ENT64) return __umul64hi(spk_hash64(r), T);
    return (uint64_t)__umulhi(fmix32((uint32_t)r), (uint32_t)T);
}

// ENT32 slot: (r << 32) | count, empty = 0.  ENT64: keys[T] (empty = ~0) then counts[T].
template <bool ENT64>
__global__ void __launch_bounds__(256)
k_part_count(const void* __restrict__ buf, const uint32_t* __restrict__ pstart, int g0, int G, int P,
             int ctas_per_part, uint8_t* __restrict__ tables, uint64_t T, uint64_t* __restrict__ stats) {
    const int gslot = blockIdx.x / ctas_per_part;
    const int p = g0 + gslot;
    if (gslot >= G || p >= P) return;
    const int cta = blockIdx.x - gslot * ctas_per_part;
    const uint64_t beg = pstart[p], end = pstart[p + 1];
    uint64_t* tkeys = (uint64_t*)(tables + (size_t)gslot * T * (ENT64 ? 12 : 8));
    uint32_t* tcnt = ENT64 ? (uint32_t*)(tkeys + T) : nullptr;
    const uint64_t stride = (uint64_t)ctas_per_part * 256;
    uint64_t n_fail = 0;
    for (uint64_t base = beg + (uint64_t)cta * 256 + threadIdx.x; base < end; base += stride * PC_BATCH) {
        uint64_t r[PC_BATCH], slot[PC_BATCH], cur[PC_BATCH];
        bool ok[PC_BATCH];
#pragma unroll
        for (int j = 0; j < PC_BATCH; j++) {
            const uint64_t i = base + (uint64_t)j * stride;
            ok[j] = i < end;
            r[j] = 0;
            if (ok[j]) r[j] = ENT64 ? __ldcs((const uint64_t*)buf + i) : (uint64_t)__ldcs((const uint32_t*)buf + i);
        }
#pragma unroll
        for (int j = 0; j < PC_BATCH; j++) {
            slot[j] = slot_of_r<ENT64>(r[j], T);
            cur[j] = ok[j] ? __ldcg(tkeys + slot[j]) : 0;
        }
#pragma unroll
        for (int j = 0; j < PC_BATCH; j++) {
            if (!ok[j]) continue;
            uint64_t s = slot[j], c = cur[j];
            bool done = false;
            for (uint64_t probes = 0; probes < T; probes++) {
                if (!ENT64) {
                    if (c == 0) {
                        const uint64_t old = atomicCAS((unsigned long long*)(tkeys + s), 0ull,
                                                       (unsigned long long)((r[j] << 32) | 1ull));
                        if (old == 0) { done = true; break; }
                        c = old;
                    }
                    if ((c >> 32) == r[j]) {
                        atomicAdd((unsigned long long*)(tkeys + s), 1ull);
                        done = true;
                        break;
                    }
                } else {
                    if (c == SPK_EMPTY_KEY) {
                        const uint64_t old = atomicCAS((unsigned long long*)(tkeys + s),
                                                       (unsigned long long)SPK_EMPTY_KEY, (unsigned long long)r[j]);
                        c = (old == SPK_EMPTY_KEY) ? r[j] : old;
                    }
                    if (c == r[j]) {
                        atomicAdd(tcnt + s, 1u);
                        done = true;
                        break;
                    }
                }
                s++;
                if (s == T) s = 0;
                c = __ldcg(tkeys + s);
            }
            if (!done) n_fail++;
        }
    }
    if (n_fail) atomicAdd((unsigned long long*)&stats[1], (unsigned long long)n_fail);
}

template <bool ENT64>
__global__ void __launch_bounds__(256)
k_part_extract(int g0, int G, int P, int ctas_per_part, uint8_t* __restrict__ tables, uint64_t T, Mixer mx,
               uint32_t lower, uint64_t* __restrict__ out_keys, uint32_t* __restrict__ out_counts, uint64_t cap,
               uint64_t* __restrict__ cursor, uint64_t* __restrict__ stats, uint64_t* __restrict__ histo,
               uint32_t histo_len) {
    __shared__ uint64_t s_red[4][8];
    __shared__ uint32_t s_hist[256];
    const int gslot = blockIdx.x / ctas_per_part;
    const int p = g0 + gslot;
    const bool active = gslot < G && p < P;
    const int cta = blockIdx.x - gslot * ctas_per_part;
    uint64_t* tkeys = (uint64_t*)(tables + (size_t)gslot * T * (ENT64 ? 12 : 8));
    uint32_t* tcnt = ENT64 ? (uint32_t*)(tkeys + T) : nullptr;
    uint64_t distinct = 0, nge = 0, sumge = 0, sumall = 0;
    uint32_t h1 = 0, h2 = 0;
    if (histo) {
        s_hist[threadIdx.x] = 0;
        __syncthreads();
    }
    const uint64_t per = (T + ctas_per_part - 1) / ctas_per_part;
    const uint64_t beg = min((uint64_t)cta * per, T), end = active ? min(beg + per, T) : beg;
    for (uint64_t base = beg; base < end; base += 256) {
        const uint64_t i = base + threadIdx.x;
        uint64_t r = 0, cnt = 0;
        bool occ = false;
        if (i < end) {
            if (!ENT64) {
                const uint64_t v = tkeys[i];
                if (v != 0) {
                    occ = true;
                    r = v >> 32;
                    cnt = v & 0xffffffffull;
                    tkeys[i] = 0;
                }
            } else {
                const uint64_t v = tkeys[i];
                if (v != SPK_EMPTY_KEY) {
                    occ = true;
                    r = v;
                    cnt = tcnt[i];
                    tkeys[i] = SPK_EMPTY_KEY;
                    tcnt[i] = 0;
                }
            }
        }
        const bool keep = occ && cnt >= lower;
        if (occ) {
            distinct++;
            sumall += cnt;
            if (keep) {
                nge++;
                sumge += cnt;
            }
            if (histo) {
                const uint64_t b = cnt < (uint64_t)(histo_len - 1) ? cnt : (uint64_t)(histo_len - 1);
                if (b == 1) h1++;
                else if (b == 2) h2++;
                else if (b < 256) atomicAdd(&s_hist[b], 1u);
                else atomicAdd((unsigned long long*)&histo[b], 1ull);
            }
        }
        // warp-aggregated append
        const uint32_t ballot = __ballot_sync(0xffffffffu, keep);
        if (ballot) {
            const int lane = threadIdx.x & 31;
            uint64_t wbase = 0;
            if (lane == 0) wbase = atomicAdd((unsigned long long*)cursor, (unsigned long long)__popc(ballot));
            wbase = __shfl_sync(0xffffffffu, wbase, 0);
            if (keep) {
                const uint64_t o = wbase + __popc(ballot & ((1u << lane) - 1));
                if (o < cap) {
                    out_keys[o] = mx.inv(((uint64_t)p << mx.rbits) | r);
                    out_counts[o] = (uint32_t)cnt;
                }
            }
        }
    }
    if (histo) {
        h1 = spk_warp_sum_u32(h1);
        h2 = spk_warp_sum_u32(h2);
        if ((threadIdx.x & 31) == 0) {
            if (h1) atomicAdd(&s_hist[1], h1);
            if (h2) atomicAdd(&s_hist[2], h2);
        }
        __syncthreads();
        if (threadIdx.x < histo_len && s_hist[threadIdx.x])
            atomicAdd((unsigned long long*)&histo[threadIdx.x], (unsigned long long)s_hist[threadIdx.x]);
    }
    uint64_t v[4] = {distinct, nge, sumge, sumall};
#pragma unroll
    for (int q = 0; q < 4; q++) {
        v[q] = spk_warp_sum_u64(v[q]);
        if ((threadIdx.x & 31) == 0) s_red[q][threadIdx.x >> 5] = v[q];
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        uint64_t s = 0;
        for (int w = 0; w < 8; w++) s += s_red[threadIdx.x][w];
        if (s) atomicAdd((unsigned long long*)&stats[4 + threadIdx.x], (unsigned long long)s);
    }
}

template <bool ENT64>
int run_plan(const PcPlan& pl, const uint8_t* pk, const uint8_t* vl, uint32_t lower, char* ws, uint64_t* d_keys,
             uint32_t* d_counts, uint64_t cap, uint64_t* d_stats, uint64_t* d_histo, uint32_t histo_len,
             cudaStream_t st) {
    void* buf = ws + pl.off_buf;
    uint32_t* hist = (uint32_t*)(ws + pl.off_hist);
    uint32_t* psize = (uint32_t*)(ws + pl.off_psize);
    uint32_t* pstart = (uint32_t*)(ws + pl.off_pstart);
    uint8_t* tables = (uint8_t*)(ws + pl.off_tables);
    uint64_t* cursor = (uint64_t*)(ws + pl.off_cursor);
    SPK_CUDA(cudaMemsetAsync(cursor, 0, 256, st));
    // tables: ENT32 empty = 0; ENT64 keys empty = 0xFF.., counts 0 (k_part_extract restores this state)
    if (!ENT64) {
        SPK_CUDA(cudaMemsetAsync(tables, 0, (size_t)pl.G * pl.T * 8, st));
    } else {
        for (int g = 0; g < pl.G; g++) {
            SPK_CUDA(cudaMemsetAsync(tables + (size_t)g * pl.T * 12, 0xFF, pl.T * 8, st));
            SPK_CUDA(cudaMemsetAsync(tables + (size_t)g * pl.T * 12 + pl.T * 8, 0, pl.T * 4, st));
        }
    }
    if (pl.n_chunks == 0) return SPK_OK;
    k_part_pass<false, ENT64><<<(unsigned)pl.n_chunks, SPK_TILE_THREADS, 0, st>>>(
        pk, vl, pl.n_tiles, pl.k, pl.mx, pl.P, hist, nullptr, nullptr, d_stats);
    SPK_LAUNCH_CHECK();
    k_part_colscan<<<(pl.P + 31) / 32, 1024, 0, st>>>(hist, pl.n_chunks, pl.P, psize);
    SPK_LAUNCH_CHECK();
    k_part_starts<<<1, 1024, 0, st>>>(psize, pl.P, pstart);
    SPK_LAUNCH_CHECK();
    k_part_pass<true, ENT64><<<(unsigned)pl.n_chunks, SPK_TILE_THREADS, 0, st>>>(
        pk, vl, pl.n_tiles, pl.k, pl.mx, pl.P, hist, pstart, buf, d_stats);
    SPK_LAUNCH_CHECK();
    const int sms = spk_num_sms();
    int cpp = (sms * 6 + pl.G - 1) / pl.G;      // CTAs per partition in the count kernel
    if (cpp < 1) cpp = 1;
    int cpe = (sms * 4 + pl.G - 1) / pl.G;
    for (int g0 = 0; g0 < pl.P; g0 += pl.G) {
        k_part_count<ENT64><<<cpp * pl.G, 256, 0, st>>>(buf, pstart, g0, pl.G, pl.P, cpp, tables, pl.T, d_stats);
        SPK_LAUNCH_CHECK();
        k_part_extract<ENT64><<<cpe * pl.G, 256, 0, st>>>(g0, pl.G, pl.P, cpe, tables, pl.T, pl.mx, lower, d_keys,
                                                          d_counts, cap, cursor, d_stats, d_histo, histo_len);
        SPK_LAUNCH_CHECK();
    }
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
