// K9: per-position lookup of subgenome-specific k-mers, counted into (bin, chunk) lines.
//
// Replaces Seqs.map_kmer3 / map_kmer_each4 / _get_kmer (Seqs.py:74-119, 209-244): for every start
// position i, `sg = d_kmers[seq[i:i+k]]`; on a hit `d_bin[(i+offset)//bin_size][sg] += 1`.
// The reference dict holds each specific k-mer and its reverse complement (Cluster.py:174-175), so a
// forward-strand lookup hits exactly when the CANONICAL k-mer is in the matrix; the table here stores
// canonical keys only (half the size, stays L2-resident).  Shipped layout: the bucketed quotient table below
// (k_qt_build / k_map_bins_q, one 16/32-byte load per position); k_sig_build / k_map_bins (open addressing +
// one-hash bitmap) remain for keys wider than the quotient slots.
#include <stdlib.h>
#include "spk_common.cuh"
#include "spk_tile.cuh"
#include "spk_mixer.cuh"

namespace {

constexpr int MP_SMEM_LINES = 16;
constexpr int MP_MAX_S = 32;

__global__ void __launch_bounds__(256)
k_sig_build(const uint64_t* __restrict__ keys, const uint8_t* __restrict__ vals, uint64_t n,
            uint64_t* __restrict__ skeys, uint8_t* __restrict__ svals, uint64_t sslots,
            uint32_t* __restrict__ filter, uint64_t fmask, int pack_vals, uint64_t* __restrict__ fail) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t key = keys[i];
        const uint64_t h = spk_hash64(key);
        if (pack_vals) {
            // value rides in the top byte of the slot: one load answers "present?" and "which subgenome?"
            const uint64_t word = key | ((uint64_t)vals[i] << 56);
            uint64_t slot = spk_slot_of(h, sslots);
            bool done = false;
            for (uint64_t p = 0; p < sslots; p++) {
                const uint64_t old = atomicCAS((unsigned long long*)(skeys + slot),
                                               (unsigned long long)SPK_EMPTY_KEY, (unsigned long long)word);
                if (old == SPK_EMPTY_KEY || (old & 0x00ffffffffffffffull) == key) {
                    done = true;
                    break;
                }
                slot++;
                if (slot == sslots) slot = 0;
            }
            if (filter) {
                const uint64_t fb = h & fmask;
                atomicOr(&filter[fb >> 5], 1u << (fb & 31));
            }
            if (!done) atomicAdd((unsigned long long*)fail, 1ull);
            continue;
        }
        if (filter) {
            const uint64_t fb = h & fmask;
            atomicOr(&filter[fb >> 5], 1u << (fb & 31));
        }
        uint64_t slot = spk_slot_of(h, sslots);
        bool done = false;
        for (uint64_t p = 0; p < sslots; p++) {
            const uint64_t old = atomicCAS((unsigned long long*)(skeys + slot),
                                           (unsigned long long)SPK_EMPTY_KEY, (unsigned long long)key);
            if (old == SPK_EMPTY_KEY || old == key) {
                svals[slot] = vals[i];
                done = true;
                break;
            }
            slot++;
            if (slot == sslots) slot = 0;
        }
        if (!done) atomicAdd((unsigned long long*)fail, 1ull);
    }
}

struct MapArgs {
    const uint64_t* skeys;
    const uint8_t* svals;
    uint64_t sslots;
    const uint32_t* filter;   // one-hash Bloom bitmap over the keys (nullptr: none)
    uint64_t fmask;
    int pack_vals;            // subgenome id stored in the top byte of the key slot (k <= 28)
    int S;
    uint64_t bin_size;
    uint64_t chunk_size;
    uint32_t* line_counts;
    uint64_t n_lines;
    uint8_t* hit_flags;
    uint64_t* nhits;
    // multi-record FASTA (Seqs.py:121-153 treat every record separately): record r occupies packed positions
    // [rec_start[r], rec_start[r+1]) (incl. its separator base) and owns the lines from rec_line0[r]
    const uint64_t* rec_start;   // [n_rec + 1], ascending; nullptr: one record
    const uint64_t* rec_line0;   // [n_rec]
    uint32_t n_rec;
};

__device__ __forceinline__ uint64_t line_of(uint64_t pos, int k, uint64_t bin_size,
                                            uint64_t chunk_size) {
    uint64_t l = pos / bin_size;
    if (chunk_size) l += (pos + (uint64_t)(k - 1)) / chunk_size;
    return l;
}

__global__ void __launch_bounds__(SPK_TILE_THREADS, 3)
k_map_bins(const uint8_t* __restrict__ packed, const uint8_t* __restrict__ valid, uint64_t n_bases,
           int k, MapArgs a) {
    __shared__ SpkTileSmem sm;
    __shared__ uint32_t s_cnt[MP_SMEM_LINES * MP_MAX_S];
    __shared__ uint64_t s_line0;
    const uint64_t n_tiles = (n_bases + SPK_TILE_BASES - 1) / SPK_TILE_BASES;
    const int tid = threadIdx.x;
    const SpkKmerParams kp = spk_kmer_params(k);
    spk_tile_init(sm);
    uint64_t n_hit = 0;

    uint64_t tile = blockIdx.x;
    if (tid == 0 && tile < n_tiles) spk_tile_issue(sm, packed, valid, tile, 0);

    for (uint32_t it = 0; tile < n_tiles; it++, tile += gridDim.x) {
        const int buf = it & 1;
        const uint32_t parity = (it >> 1) & 1;
        for (int i = tid; i < MP_SMEM_LINES * MP_MAX_S; i += SPK_TILE_THREADS) s_cnt[i] = 0;
        if (tid == 0) s_line0 = line_of(tile * SPK_TILE_BASES, k, a.bin_size, a.chunk_size);
        __syncthreads();  // buffer buf^1 free, counters cleared, line0 visible
        if (tid == 0 && tile + gridDim.x < n_tiles)
            spk_tile_issue(sm, packed, valid, tile + gridDim.x, buf ^ 1);
        spk_mbar_wait(&sm.bar[buf], parity);

        uint64_t key[SPK_KMERS_PER_THREAD];
        uint32_t okmask;
        spk_tile_kmers(sm, buf, kp, key, okmask);

        // line bookkeeping for this thread's 16 consecutive positions
        const uint64_t pos0 = tile * SPK_TILE_BASES + (uint64_t)tid * SPK_KMERS_PER_THREAD;
        uint64_t bin = pos0 / a.bin_size;
        uint64_t brem = pos0 - bin * a.bin_size;
        uint64_t chk = 0, crem = 0;
        if (a.chunk_size) {
            chk = (pos0 + (uint64_t)(k - 1)) / a.chunk_size;
            crem = (pos0 + (uint64_t)(k - 1)) - chk * a.chunk_size;
        }
        const uint64_t line0 = s_line0;

        // stage 1: one-bit membership filter for all 16 positions (a miss costs a single 4-byte load;
        // ~97 % of the positions of a real chromosome are misses)
        uint64_t hsh[SPK_KMERS_PER_THREAD];
        uint32_t cand = okmask;
        if (a.filter) {
            uint32_t fw[SPK_KMERS_PER_THREAD];
#pragma unroll
            for (int j = 0; j < SPK_KMERS_PER_THREAD; j++) {
                hsh[j] = spk_hash64(key[j]);
                fw[j] = ((okmask >> j) & 1u) ? __ldg(a.filter + ((hsh[j] & a.fmask) >> 5)) : 0u;
            }
#pragma unroll
            for (int j = 0; j < SPK_KMERS_PER_THREAD; j++)
                if (!((fw[j] >> (hsh[j] & 31)) & 1u)) cand &= ~(1u << j);
        } else {
#pragma unroll
            for (int j = 0; j < SPK_KMERS_PER_THREAD; j++) hsh[j] = spk_hash64(key[j]);
        }
        // stage 2: table probes for the candidates
#pragma unroll
        for (int j = 0; j < SPK_KMERS_PER_THREAD; j++) {
            if ((cand >> j) & 1u) {
                uint64_t sl = spk_slot_of(hsh[j], a.sslots);
                uint64_t c = __ldg(a.skeys + sl);
                const uint64_t kk = key[j];
                const uint64_t kmask = a.pack_vals ? 0x00ffffffffffffffull : ~0ull;
                while (c != SPK_EMPTY_KEY && (c & kmask) != kk) {
                    sl++;
                    if (sl == a.sslots) sl = 0;
                    c = __ldg(a.skeys + sl);
                }
                if (c != SPK_EMPTY_KEY) {
                    const uint32_t sg = a.pack_vals ? (uint32_t)(c >> 56) : (uint32_t)a.svals[sl];
                    const uint64_t line = bin + chk;
                    const uint64_t rel = line - line0;
                    if (rel < MP_SMEM_LINES) atomicAdd(&s_cnt[rel * MP_MAX_S + sg], 1u);
                    else if (line < a.n_lines) atomicAdd(&a.line_counts[line * a.S + sg], 1u);
                    if (a.hit_flags && !a.hit_flags[sl]) a.hit_flags[sl] = 1;      // (orientation not tracked on this path)
                    n_hit++;
                }
            }
            // advance the (bin, chunk) cursors to the next position
            if (++brem == a.bin_size) { brem = 0; bin++; }
            if (a.chunk_size && ++crem == a.chunk_size) { crem = 0; chk++; }
        }
        __syncthreads();
        for (int i = tid; i < MP_SMEM_LINES * MP_MAX_S; i += SPK_TILE_THREADS) {
            const uint32_t c = s_cnt[i];
            if (c) {
                const uint64_t line = line0 + i / MP_MAX_S;
                const int sg = i % MP_MAX_S;
                if (line < a.n_lines && sg < a.S) atomicAdd(&a.line_counts[line * a.S + sg], c);
            }
        }
    }
    n_hit = spk_warp_sum_u64(n_hit);
    if ((tid & 31) == 0 && n_hit && a.nhits) atomicAdd((unsigned long long*)a.nhits, (unsigned long long)n_hit);
}

// ---- bucketed quotient table ---------------------------------------------------------------------------
// The per-position lookup is an L2 random-access problem (1.4e10 lookups for wheat), so the table is
// made as small and as one-touch as possible: f = bijective mixer on the 2k-bit canonical word, bucket =
// top `bbits` of f(u), and only the remainder (rbits = 2k - bbits bits) plus the subgenome id are stored:
//   slot = (remainder << sgbits) | sg        all-ones = empty (the sg field of a real entry is < S)
// A bucket is 8 slots = ONE 16-byte (u16 slots) or 32-byte (u32 slots) load: 7 entries + 1 overflow
// marker.  Membership is exact (f is a bijection); the ~1 % of keys whose bucket is full live in a small
// open-addressed stash that is probed only when the marker is set.  3.6e6 wheat k-mers: 16 MB, fully
// L2-resident, one 16-B load per position instead of a filter load + dependent 8-B probes of a 58-MB table.
constexpr int QT_SLOTS = 8;            // slots per bucket (7 entries + marker)

struct QtArgs {
    const void* buckets;
    int slot_bits;            // 16 or 32
    int bbits, sgbits;
    Mixer mx;
    const uint64_t* skeys;    // stash (open-addressed, full keys; value in the top byte or in svals)
    const uint8_t* svals;
    uint64_t sslots;
    int pack_vals;
};

template <typename SlotT>
__global__ void __launch_bounds__(256)
k_qt_build(const uint64_t* __restrict__ keys, const uint8_t* __restrict__ vals, uint64_t n, SlotT* __restrict__ buckets,
           int sgbits, Mixer mx, uint64_t* __restrict__ skeys, uint8_t* __restrict__ svals, uint64_t sslots,
           int pack_vals, uint64_t* __restrict__ fail) {
    const SlotT EMPTY = (SlotT)~(SlotT)0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t key = keys[i];
        const uint64_t h = mx.fwd_light(key);
        const uint64_t b = (mx.rbits >= 64) ? 0 : (h >> mx.rbits);
        const uint64_t rem = (mx.rbits >= 64) ? h : (h & ((1ull << mx.rbits) - 1));
        const SlotT word = (SlotT)((rem << sgbits) | (uint64_t)vals[i]);
        SlotT* bk = buckets + b * QT_SLOTS;
        bool done = false;
        for (int s = 0; s < QT_SLOTS - 1 && !done; s++) {
            const SlotT old = atomicCAS(bk + s, EMPTY, word);
            if (old == EMPTY || (old >> sgbits) == (word >> sgbits)) done = true;
        }
        if (done) continue;
        bk[QT_SLOTS - 1] = 0;            // marker: this bucket overflowed into the stash
        const uint64_t sw = pack_vals ? (key | ((uint64_t)vals[i] << 56)) : key;
        const uint64_t kmask = pack_vals ? 0x00ffffffffffffffull : ~0ull;
        uint64_t slot = spk_slot_of(spk_hash64(key), sslots);
        for (uint64_t p = 0; p < sslots; p++) {
            const uint64_t old = atomicCAS((unsigned long long*)(skeys + slot), (unsigned long long)SPK_EMPTY_KEY,
                                           (unsigned long long)sw);
            if (old == SPK_EMPTY_KEY || (old & kmask) == key) {
                if (!pack_vals) svals[slot] = vals[i];
                done = true;
                break;
            }
            slot++;
            if (slot == sslots) slot = 0;
        }
        if (!done) atomicAdd((unsigned long long*)fail, 1ull);
    }
}

// match a loaded bucket against q = remainder << sgbits: subgenome id or -1; `at` = slot of the hit,
// `over` = the bucket overflowed into the stash at build time
template <bool S16>
struct QtBucket;
template <>
struct QtBucket<true> {
    uint4 v;
    __device__ __forceinline__ void load(const void* buckets, uint32_t b) {
        v = __ldg(reinterpret_cast<const uint4*>(buckets) + b);
    }
    __device__ __forceinline__ void none() { v = make_uint4(~0u, ~0u, ~0u, ~0u); }
    __device__ __forceinline__ int match(uint32_t q, uint32_t S, int& at, bool& over) const {
        const uint32_t qq = q * 0x10001u;
        const uint32_t w[4] = {v.x ^ qq, v.y ^ qq, v.z ^ qq, v.w ^ qq};
        int sg = -1;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const uint32_t lo = w[i] & 0xffffu, hi = w[i] >> 16;
            if (lo < S) { sg = (int)lo; at = 2 * i; }
            if (i < 3 && hi < S) { sg = (int)hi; at = 2 * i + 1; }
        }
        over = (v.w >> 16) != 0xffffu;
        return sg;
    }
};
template <>
struct QtBucket<false> {
    uint4 v0, v1;
    __device__ __forceinline__ void load(const void* buckets, uint32_t b) {
        v0 = __ldg(reinterpret_cast<const uint4*>(buckets) + 2 * (uint64_t)b);
        v1 = __ldg(reinterpret_cast<const uint4*>(buckets) + 2 * (uint64_t)b + 1);
    }
    __device__ __forceinline__ void none() { v0 = v1 = make_uint4(~0u, ~0u, ~0u, ~0u); }
    __device__ __forceinline__ int match(uint32_t q, uint32_t S, int& at, bool& over) const {
        const uint32_t w[7] = {v0.x ^ q, v0.y ^ q, v0.z ^ q, v0.w ^ q, v1.x ^ q, v1.y ^ q, v1.z ^ q};
        int sg = -1;
#pragma unroll
        for (int i = 0; i < 7; i++)
            if (w[i] < S) { sg = (int)w[i]; at = i; }
        over = v1.w != 0xffffffffu;
        return sg;
    }
};

// rare: the bucket overflowed at build time -> the key may be in the stash
__device__ __noinline__ int qt_stash_lookup(const QtArgs& a, uint64_t key, uint64_t& sl_out) {
    const uint64_t kmask = a.pack_vals ? 0x00ffffffffffffffull : ~0ull;
    uint64_t sl = spk_slot_of(spk_hash64(key), a.sslots);
    uint64_t c = __ldg(a.skeys + sl);
    while (c != SPK_EMPTY_KEY && (c & kmask) != key) {
        sl++;
        if (sl == a.sslots) sl = 0;
        c = __ldg(a.skeys + sl);
    }
    if (c == SPK_EMPTY_KEY) return -1;
    sl_out = sl;
    return a.pack_vals ? (int)(c >> 56) : (int)a.svals[sl];
}

template <bool S16>
__global__ void __launch_bounds__(SPK_TILE_THREADS, 3)
k_map_bins_q(const uint8_t* __restrict__ packed, const uint8_t* __restrict__ valid, uint64_t n_bases, int k,
             MapArgs a, QtArgs qa) {
    __shared__ SpkTileSmem sm;
    const uint64_t n_tiles = (n_bases + SPK_TILE_BASES - 1) / SPK_TILE_BASES;
    const int tid = threadIdx.x;
    const SpkKmerParams kp = spk_kmer_params(k);
    spk_tile_init(sm);
    uint64_t n_hit = 0;
    const int S = a.S;
    const uint64_t rmask = (qa.mx.rbits >= 64) ? ~0ull : ((1ull << qa.mx.rbits) - 1);

    uint64_t tile = blockIdx.x;
    if (tid == 0 && tile < n_tiles) spk_tile_issue(sm, packed, valid, tile, 0);
    // (bin, chunk) cursors of this thread's first position: divided once, then advanced by the constant
    // grid stride (64-bit divisions per tile were ~10 % of the kernel's instructions)
    uint64_t pos0 = tile * SPK_TILE_BASES + (uint64_t)tid * SPK_KMERS_PER_THREAD;
    uint64_t bin = pos0 / a.bin_size;
    uint64_t brem = pos0 - bin * a.bin_size;
    uint64_t chk = 0, crem = 0;
    if (a.chunk_size) {
        chk = (pos0 + (uint64_t)(k - 1)) / a.chunk_size;
        crem = (pos0 + (uint64_t)(k - 1)) - chk * a.chunk_size;
    }
    const uint64_t stride = (uint64_t)gridDim.x * SPK_TILE_BASES;
    const uint64_t dbin = stride / a.bin_size, dbrem = stride - dbin * a.bin_size;
    const uint64_t dchk = a.chunk_size ? stride / a.chunk_size : 0, dcrem = a.chunk_size ? stride - dchk * a.chunk_size : 0;

    for (uint32_t it = 0; tile < n_tiles; it++, tile += gridDim.x) {
        const int buf = it & 1;
        const uint32_t parity = (it >> 1) & 1;
        __syncthreads();  // every warp is done with buffer buf^1 (the only CTA-wide barrier per tile)
        if (tid == 0 && tile + gridDim.x < n_tiles)
            spk_tile_issue(sm, packed, valid, tile + gridDim.x, buf ^ 1);
        spk_mbar_wait(&sm.bar[buf], parity);

        uint64_t key[SPK_KMERS_PER_THREAD];
        uint32_t okmask;
        spk_tile_kmers(sm, buf, kp, key, okmask);

        // ---- fast path (16-bit slots, the warp's 512 positions fall on one line, no per-entry hit flags wanted) ----
        // The 8 slots of a bucket are matched with SWAR arithmetic on the four 32-bit words of the 16-byte load (a slot
        // matches iff (slot ^ q) has no bit above the subgenome field: a zero-halfword test on the masked words), and a
        // hit goes straight into four packed 16-bit counters — no per-position hit masks, no slot-by-slot compares.
        // Only hits (and the rare bucket that overflowed into the stash) leave the straight-line code.
        if constexpr (S16) {
            const bool one_line_f = (brem + SPK_KMERS_PER_THREAD <= a.bin_size) &&
                                    (!a.chunk_size || crem + SPK_KMERS_PER_THREAD <= a.chunk_size);
            const uint64_t line_f = bin + chk;
            const uint32_t lo_f = (uint32_t)line_f, hi_f = (uint32_t)(line_f >> 32);
            const uint32_t lo_0 = __shfl_sync(0xffffffffu, lo_f, 0), hi_0 = __shfl_sync(0xffffffffu, hi_f, 0);
            const bool fast = __all_sync(0xffffffffu, one_line_f && lo_f == lo_0 && hi_f == hi_0) && S <= 4 &&
                              !a.rec_start && !a.hit_flags;
            if (fast) {
                const uint32_t hmask = (0xffffu << qa.sgbits) & 0xffffu;
                const uint32_t HM = hmask * 0x10001u;                   // bits above the subgenome field, both halfwords
                uint32_t c01 = 0, c23 = 0, rare = 0;
                constexpr int QB = 4;
#pragma unroll
                for (int j0 = 0; j0 < SPK_KMERS_PER_THREAD; j0 += QB) {
                    uint4 bv[QB];
                    uint32_t qq[QB];
#pragma unroll
                    for (int jj = 0; jj < QB; jj++) {
                        const int j = j0 + jj;
                        const uint64_t h = qa.mx.fwd_light(key[j]);
                        qq[jj] = (uint32_t)((h & rmask) << qa.sgbits) * 0x10001u;
                        bv[jj] = make_uint4(~0u, ~0u, ~0u, ~0u);
                        if ((okmask >> j) & 1u)
                            bv[jj] = __ldg(reinterpret_cast<const uint4*>(qa.buckets) + (uint32_t)(h >> qa.mx.rbits));
                    }
#pragma unroll
                    for (int jj = 0; jj < QB; jj++) {
                        const int j = j0 + jj;
                        const uint32_t w[4] = {bv[jj].x ^ qq[jj], bv[jj].y ^ qq[jj], bv[jj].z ^ qq[jj], bv[jj].w ^ qq[jj]};
                        uint32_t z = 0;
#pragma unroll
                        for (int i = 0; i < 4; i++) {
                            const uint32_t t = w[i] & HM;
                            const uint32_t zi = ~(((t & 0x7fff7fffu) + 0x7fff7fffu) | t) & 0x80008000u;   // exact zero-halfword flags
                            z |= (i == 3) ? (zi & 0x8000u) : zi;        // slot 7 (high half of the last word) is the marker
                        }
                        const bool okj = (okmask >> j) & 1u;
                        if (z && okj) {
                            int sg = -1;
#pragma unroll
                            for (int i = 0; i < 4; i++) {
                                const uint32_t lo = w[i] & 0xffffu, hi = w[i] >> 16;
                                if (lo < (uint32_t)S) sg = (int)lo;
                                if (i < 3 && hi < (uint32_t)S) sg = (int)hi;
                            }
                            if (sg >= 0) {
                                const uint32_t inc = 1u << ((sg & 1) * 16);
                                if (sg & 2) c23 += inc;
                                else c01 += inc;
                            } else if ((bv[jj].w >> 16) != 0xffffu) rare |= 1u << j;
                        } else if (okj && (bv[jj].w >> 16) != 0xffffu) rare |= 1u << j;
                    }
                }
                while (rare) {                                          // bucket overflowed at build time: the key may be in the stash
                    const int j = __ffs(rare) - 1;
                    rare &= rare - 1;
                    const uint64_t kj = spk_kmer_at(sm.packed[buf], tid * SPK_KMERS_PER_THREAD + j, kp);
                    uint64_t sl = 0;
                    const int sg = qt_stash_lookup(qa, kj, sl);
                    if (sg >= 0) {
                        const uint32_t inc = 1u << ((sg & 1) * 16);
                        if (sg & 2) c23 += inc;
                        else c01 += inc;
                    }
                }
                n_hit += (c01 & 0xffffu) + (c01 >> 16) + (c23 & 0xffffu) + (c23 >> 16);
                c01 = __reduce_add_sync(0xffffffffu, c01);
                c23 = __reduce_add_sync(0xffffffffu, c23);
                if ((tid & 31) == 0 && line_f < a.n_lines) {
                    const uint32_t c[4] = {c01 & 0xffffu, c01 >> 16, c23 & 0xffffu, c23 >> 16};
#pragma unroll
                    for (int q4 = 0; q4 < 4; q4++)
                        if (c[q4] && q4 < S) atomicAdd(&a.line_counts[line_f * a.S + q4], c[q4]);
                }
                pos0 += stride;
                bin += dbin;
                brem += dbrem;
                if (brem >= a.bin_size) { brem -= a.bin_size; bin++; }
                if (a.chunk_size) {
                    chk += dchk;
                    crem += dcrem;
                    if (crem >= a.chunk_size) { crem -= a.chunk_size; chk++; }
                }
                continue;
            }
        }
        // bucket lookups in batches of QB independent 16/32-byte loads (issued before any is consumed)
        uint32_t hm = 0;              // bit j: position j hit
        uint32_t rare = 0;            // bit j: position j needs the out-of-line path (stash probe / hit flag)
        uint64_t sgp[2] = {0, 0};     // subgenome id of hit j, 8 bits each
        constexpr int QB = 4;
#pragma unroll
        for (int j0 = 0; j0 < SPK_KMERS_PER_THREAD; j0 += QB) {
            QtBucket<S16> bk[QB];
            uint32_t qv[QB];
#pragma unroll
            for (int jj = 0; jj < QB; jj++) {
                const int j = j0 + jj;
                const uint64_t h = qa.mx.fwd_light(key[j]);
                qv[jj] = (uint32_t)((h & rmask) << qa.sgbits);
                if ((okmask >> j) & 1u) bk[jj].load(qa.buckets, (qa.mx.rbits >= 64) ? 0u : (uint32_t)(h >> qa.mx.rbits));
                else bk[jj].none();
            }
#pragma unroll
            for (int jj = 0; jj < QB; jj++) {
                const int j = j0 + jj;
                int at = 0;
                bool over;
                const int sg = bk[jj].match(qv[jj], (uint32_t)S, at, over);
                const bool okj = (okmask >> j) & 1u;
                if (okj && sg >= 0) {
                    hm |= 1u << j;
                    sgp[j >> 3] |= (uint64_t)sg << (8 * (j & 7));
                }
                if (okj && ((sg < 0 && over) || (sg >= 0 && a.hit_flags != nullptr))) rare |= 1u << j;
            }
        }
        // out-of-line (not unrolled): the bucket overflowed at build time and the key may sit in the stash, or the
        // caller wants to know WHICH table entries were hit (the "mapped k-mers" log line of Seqs.py:109-117).
        // The k-mer is recomputed from the tile instead of indexing key[] dynamically (no local memory).
        while (rare) {
            const int j = __ffs(rare) - 1;
            rare &= rare - 1;
            bool as_read;
            const uint64_t kj = spk_kmer_at(sm.packed[buf], tid * SPK_KMERS_PER_THREAD + j, kp, &as_read);
            const uint64_t h = qa.mx.fwd_light(kj);
            const uint64_t b = (qa.mx.rbits >= 64) ? 0ull : (h >> qa.mx.rbits);
            uint64_t where = 0;
            int sg = -1;
            if ((hm >> j) & 1u) {                       // bucket hit: find the slot again (flag bookkeeping only)
                QtBucket<S16> bq;
                bq.load(qa.buckets, (uint32_t)b);
                int at = 0;
                bool over;
                sg = bq.match((uint32_t)((h & rmask) << qa.sgbits), (uint32_t)S, at, over);
                where = b * QT_SLOTS + at;
            } else {
                uint64_t sl = 0;
                sg = qt_stash_lookup(qa, kj, sl);
                where = ((uint64_t)QT_SLOTS << qa.bbits) + sl;
                if (sg >= 0) {
                    hm |= 1u << j;
                    sgp[j >> 3] |= (uint64_t)sg << (8 * (j & 7));
                }
            }
            // bit 0: seen as stored (canonical orientation), bit 1: seen as its reverse complement — the reference's
            // set of mapped k-mer STRINGS (Seqs.py:109,114-117) holds the two orientations separately
            if (sg >= 0 && a.hit_flags) {
                const uint8_t bit = as_read ? 1 : 2;
                if (!(a.hit_flags[where] & bit)) atomicOr((unsigned int*)(a.hit_flags + (where & ~3ull)),
                                                          (unsigned int)bit << (8 * (where & 3)));
            }
        }
        n_hit += __popc(hm);

        // fast path: the whole warp (512 positions) falls on one line -> warp-reduce the per-subgenome hit
        // counts and let one lane add them to the line's counters (S REDs per warp and tile; nothing is staged
        // in shared memory, so the tile loop needs no second barrier)
        const bool one_line = (brem + SPK_KMERS_PER_THREAD <= a.bin_size) &&
                              (!a.chunk_size || crem + SPK_KMERS_PER_THREAD <= a.chunk_size);
        const uint64_t line_t = bin + chk;
        const uint32_t lo_t = (uint32_t)line_t, hi_t = (uint32_t)(line_t >> 32);
        const uint32_t lo_first = __shfl_sync(0xffffffffu, lo_t, 0), hi_first = __shfl_sync(0xffffffffu, hi_t, 0);
        const bool warp_one_line =
            __all_sync(0xffffffffu, one_line && lo_t == lo_first && hi_t == hi_first) && S <= 4 && !a.rec_start;
        if (a.rec_start) {
            // record mode (small inputs: LTRs, custom features): per hit, find the record and its local line
            uint32_t m = hm;
            while (m) {
                const int j = __ffs(m) - 1;
                m &= m - 1;
                const uint32_t sg = (uint32_t)(sgp[j >> 3] >> (8 * (j & 7))) & 0xffu;
                const uint64_t pos = pos0 + j;
                uint32_t lo = 0, hi = a.n_rec;               // last record with rec_start[r] <= pos
                while (hi - lo > 1) {
                    const uint32_t mid = (lo + hi) >> 1;
                    if (a.rec_start[mid] <= pos) lo = mid;
                    else hi = mid;
                }
                const uint64_t line = a.rec_line0[lo] + line_of(pos - a.rec_start[lo], k, a.bin_size, a.chunk_size);
                if (line < a.n_lines) atomicAdd(&a.line_counts[line * a.S + sg], 1u);
            }
        } else if (warp_one_line) {
            uint32_t c01 = 0, c23 = 0;   // four 16-bit counters (<= 16 per thread, <= 512 per warp)
#pragma unroll
            for (int j = 0; j < SPK_KMERS_PER_THREAD; j++) {
                const uint32_t sg = (uint32_t)(sgp[j >> 3] >> (8 * (j & 7))) & 0xffu;
                const uint32_t hit = (hm >> j) & 1u;
                const uint32_t inc = hit << ((sg & 1u) * 16);
                if (sg & 2u) c23 += inc;
                else c01 += inc;
            }
            c01 = __reduce_add_sync(0xffffffffu, c01);
            c23 = __reduce_add_sync(0xffffffffu, c23);
            if ((tid & 31) == 0 && line_t < a.n_lines) {
                const uint32_t c[4] = {c01 & 0xffffu, c01 >> 16, c23 & 0xffffu, c23 >> 16};
#pragma unroll
                for (int s = 0; s < 4; s++)
                    if (c[s] && s < S) atomicAdd(&a.line_counts[line_t * a.S + s], c[s]);
            }
        } else {
            // a bin / chunk border inside the warp's 512 positions (or S > 4): per-hit adds, line by division
            uint32_t m = hm;
            while (m) {
                const int j = __ffs(m) - 1;
                m &= m - 1;
                const uint32_t sg = (uint32_t)(sgp[j >> 3] >> (8 * (j & 7))) & 0xffu;
                const uint64_t line = line_of(pos0 + j, k, a.bin_size, a.chunk_size);
                if (line < a.n_lines) atomicAdd(&a.line_counts[line * a.S + sg], 1u);
            }
        }
        // advance the cursors to this thread's first position in the CTA's next tile
        pos0 += stride;
        bin += dbin;
        brem += dbrem;
        if (brem >= a.bin_size) { brem -= a.bin_size; bin++; }
        if (a.chunk_size) {
            chk += dchk;
            crem += dcrem;
            if (crem >= a.chunk_size) { crem -= a.chunk_size; chk++; }
        }
    }
    n_hit = spk_warp_sum_u64(n_hit);
    if ((tid & 31) == 0 && n_hit && a.nhits) atomicAdd((unsigned long long*)a.nhits, (unsigned long long)n_hit);
}

// ---- warp-private variant of k_map_bins_q<16-bit slots> ----------------------------------------------------------
// The tile pipeline above synchronises 8 warps per 4096 bases (TMA into shared memory, one CTA barrier per tile); with
// every position a dependent 16-byte gather from L2, the warps of a CTA finish a tile at different times and a third
// of the issue slots went to waiting at that barrier.  Here a warp owns 512 consecutive positions at a time and
// nothing is shared: lane l loads packed word l of the warp's 128-byte line (plus two tail words and 17 validity
// words), neighbours' words arrive by shuffle, the next line is loaded into registers before the lookups of the
// current one start.  No shared memory, no barrier.  Serves 16-bit slots, S <= 4, one record, no hit flags — the
// genome-scale calls; everything else stays with k_map_bins_q.
constexpr int MW_THREADS = 128;

__global__ void __launch_bounds__(MW_THREADS, 8)
k_map_bins_w(const uint32_t* __restrict__ packed, const uint32_t* __restrict__ valid, uint64_t n_bases, int k, MapArgs a,
             QtArgs qa) {
    constexpr uint32_t FULL = 0xffffffffu;
    constexpr int WT = 32 * SPK_KMERS_PER_THREAD;                 // 512 positions per warp step
    const int lane = threadIdx.x & 31;
    const uint64_t nw = (uint64_t)gridDim.x * (MW_THREADS / 32);
    uint64_t wt = (uint64_t)blockIdx.x * (MW_THREADS / 32) + (threadIdx.x >> 5);
    const uint64_t n_wt = (n_bases + WT - 1) / WT;
    const SpkKmerParams kp = spk_kmer_params(k);
    const int S = a.S;
    const uint64_t rmask = (qa.mx.rbits >= 64) ? ~0ull : ((1ull << qa.mx.rbits) - 1);
    const uint32_t hmask = (0xffffu << qa.sgbits) & 0xffffu;
    const uint32_t HM = hmask * 0x10001u;                         // bits above the subgenome field, both halfwords
    uint64_t n_hit = 0;

    uint64_t pos0 = wt * WT + (uint64_t)lane * SPK_KMERS_PER_THREAD;
    uint64_t bin = pos0 / a.bin_size;
    uint64_t brem = pos0 - bin * a.bin_size;
    uint64_t chk = 0, crem = 0;
    if (a.chunk_size) {
        chk = (pos0 + (uint64_t)(k - 1)) / a.chunk_size;
        crem = (pos0 + (uint64_t)(k - 1)) - chk * a.chunk_size;
    }
    const uint64_t stride = nw * WT;
    const uint64_t dbin = stride / a.bin_size, dbrem = stride - dbin * a.bin_size;
    const uint64_t dchk = a.chunk_size ? stride / a.chunk_size : 0, dcrem = a.chunk_size ? stride - dchk * a.chunk_size : 0;

    uint32_t pw = 0, pt = 0, vw = 0;
    auto fetch = [&](uint64_t t, uint32_t& fw, uint32_t& ft, uint32_t& fv) {
        fw = ft = fv = 0;
        if (t < n_wt) {                                           // (both arrays are zero-padded past the sequence)
            fw = __ldg(packed + t * 32 + lane);
            if (lane < 2) ft = __ldg(packed + t * 32 + 32 + lane);
            if (lane < 17) fv = __ldg(valid + t * 16 + lane);
        }
    };
    fetch(wt, pw, pt, vw);
    for (; wt < n_wt; wt += nw) {
        uint32_t w1 = __shfl_down_sync(FULL, pw, 1), w2 = __shfl_down_sync(FULL, pw, 2);
        const uint32_t t0 = __shfl_sync(FULL, pt, 0), t1 = __shfl_sync(FULL, pt, 1);
        if (lane == 31) { w1 = t0; w2 = t1; }
        else if (lane == 30) w2 = t0;
        const uint32_t w0 = pw;
        const uint32_t v0 = __shfl_sync(FULL, vw, lane >> 1), v1 = __shfl_sync(FULL, vw, (lane >> 1) + 1);
        const uint64_t vbits = (((uint64_t)v1 << 32) | v0) >> ((lane & 1) * 16);
        uint32_t npw, npt, nvw;
        fetch(wt + nw, npw, npt, nvw);

        const bool one_line = (brem + SPK_KMERS_PER_THREAD <= a.bin_size) &&
                              (!a.chunk_size || crem + SPK_KMERS_PER_THREAD <= a.chunk_size);
        const uint64_t line = bin + chk;
        const uint32_t lo_f = (uint32_t)line, hi_f = (uint32_t)(line >> 32);
        const uint32_t lo_0 = __shfl_sync(FULL, lo_f, 0), hi_0 = __shfl_sync(FULL, hi_f, 0);
        const bool fast = __all_sync(FULL, one_line && lo_f == lo_0 && hi_f == hi_0);
        if (fast) {
            // k-mers as in spk_kmers_from_words, but batch by batch inside a ROLLED loop (run-time shift amounts): the
            // fully unrolled body was 44 KB of code, and with every warp at a different place in it the warps stalled
            // on instruction fetch more than on anything else
            const uint32_t klo = (uint32_t)kp.kmask, khi = (uint32_t)(kp.kmask >> 32);
            const uint32_t n0 = ~w0, n1 = ~w1, n2 = ~w2;
            uint32_t a0 = spk_rev2_32(w2), a1 = spk_rev2_32(w1), a2 = spk_rev2_32(w0);
            const int s = 2 * (33 - k);
            if (s >= 64) { a0 = a2; a1 = 0; a2 = 0; }
            else if (s >= 32) { a0 = a1; a1 = a2; a2 = 0; }
            const uint32_t sl5 = (uint32_t)s & 31u;
            const uint32_t f0 = __funnelshift_r(a0, a1, sl5), f1 = __funnelshift_r(a1, a2, sl5), f2 = a2 >> sl5;
            uint32_t okmask;
            {
                uint64_t x = ~vbits & ((1ull << (15 + k)) - 1);
                if (x != 0) {
                    int covered = 1;
                    while (covered < k) {
                        const int sh = min(covered, k - covered);
                        x |= x >> sh;
                        covered += sh;
                    }
                }
                okmask = (uint32_t)(~x) & 0xffffu;
            }
            uint32_t c01 = 0, c23 = 0, rare = 0;
            constexpr int QB = 4;
#pragma unroll 1
            for (int j0 = 0; j0 < SPK_KMERS_PER_THREAD; j0 += QB) {
                uint4 bv[QB];
                uint32_t qq[QB];
#pragma unroll
                for (int jj = 0; jj < QB; jj++) {
                    const int j = j0 + jj;
                    const uint32_t sf = 2u * (15u - (uint32_t)j), sr = 2u * (uint32_t)j;
                    const uint32_t flo = __funnelshift_r(f0, f1, sf) & klo, fhi = __funnelshift_r(f1, f2, sf) & khi;
                    const uint32_t rlo = __funnelshift_r(n0, n1, sr) & klo, rhi = __funnelshift_r(n1, n2, sr) & khi;
                    const uint64_t fwd = ((uint64_t)fhi << 32) | flo, rc = ((uint64_t)rhi << 32) | rlo;
                    const uint64_t h = qa.mx.fwd_light(fwd < rc ? fwd : rc);
                    qq[jj] = (uint32_t)((h & rmask) << qa.sgbits) * 0x10001u;
                    bv[jj] = make_uint4(~0u, ~0u, ~0u, ~0u);
                    if ((okmask >> j) & 1u)
                        bv[jj] = __ldg(reinterpret_cast<const uint4*>(qa.buckets) + (uint32_t)(h >> qa.mx.rbits));
                }
#pragma unroll
                for (int jj = 0; jj < QB; jj++) {
                    const int j = j0 + jj;
                    const uint32_t w[4] = {bv[jj].x ^ qq[jj], bv[jj].y ^ qq[jj], bv[jj].z ^ qq[jj], bv[jj].w ^ qq[jj]};
                    uint32_t z = 0;
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        const uint32_t t = w[i] & HM;
                        const uint32_t zi = ~(((t & 0x7fff7fffu) + 0x7fff7fffu) | t) & 0x80008000u;   // exact zero-halfword flags
                        z |= (i == 3) ? (zi & 0x8000u) : zi;        // slot 7 (high half of the last word) is the marker
                    }
                    const bool okj = (okmask >> j) & 1u;
                    if (z && okj) {
                        int sg = -1;
#pragma unroll
                        for (int i = 0; i < 4; i++) {
                            const uint32_t lo = w[i] & 0xffffu, hi = w[i] >> 16;
                            if (lo < (uint32_t)S) sg = (int)lo;
                            if (i < 3 && hi < (uint32_t)S) sg = (int)hi;
                        }
                        if (sg >= 0) {
                            const uint32_t inc = 1u << ((sg & 1) * 16);
                            if (sg & 2) c23 += inc;
                            else c01 += inc;
                        } else if ((bv[jj].w >> 16) != 0xffffu) rare |= 1u << j;
                    } else if (okj && (bv[jj].w >> 16) != 0xffffu) rare |= 1u << j;
                }
            }
            while (rare) {                                          // bucket overflowed at build time: the key may be in the stash
                const int j = __ffs(rare) - 1;
                rare &= rare - 1;
                uint64_t sl = 0;
                const int sg = qt_stash_lookup(qa, spk_kmer_of_words(w0, w1, w2, j, kp), sl);
                if (sg >= 0) {
                    const uint32_t inc = 1u << ((sg & 1) * 16);
                    if (sg & 2) c23 += inc;
                    else c01 += inc;
                }
            }
            n_hit += (c01 & 0xffffu) + (c01 >> 16) + (c23 & 0xffffu) + (c23 >> 16);
            c01 = __reduce_add_sync(FULL, c01);
            c23 = __reduce_add_sync(FULL, c23);
            if (lane == 0 && line < a.n_lines) {
                const uint32_t c[4] = {c01 & 0xffffu, c01 >> 16, c23 & 0xffffu, c23 >> 16};
#pragma unroll
                for (int q4 = 0; q4 < 4; q4++)
                    if (c[q4] && q4 < S) atomicAdd(&a.line_counts[line * a.S + q4], c[q4]);
            }
        } else {
            // a bin / chunk border inside the warp's 512 positions (~5 % of the steps): position by position
            uint64_t x = ~vbits & ((1ull << (15 + k)) - 1);         // (validity as in spk_kmers_from_words)
            if (x != 0) {
                int covered = 1;
                while (covered < k) {
                    const int sh = min(covered, k - covered);
                    x |= x >> sh;
                    covered += sh;
                }
            }
            uint32_t ok = (uint32_t)(~x) & 0xffffu;
            while (ok) {
                const int j = __ffs(ok) - 1;
                ok &= ok - 1;
                const uint64_t kj = spk_kmer_of_words(w0, w1, w2, j, kp);
                const uint64_t h = qa.mx.fwd_light(kj);
                QtBucket<true> bq;
                bq.load(qa.buckets, (uint32_t)(h >> qa.mx.rbits));
                int at = 0;
                bool over;
                int sg = bq.match((uint32_t)((h & rmask) << qa.sgbits), (uint32_t)S, at, over);
                if (sg < 0 && over) {
                    uint64_t sl = 0;
                    sg = qt_stash_lookup(qa, kj, sl);
                }
                if (sg >= 0) {
                    n_hit++;
                    const uint64_t l = line_of(pos0 + j, k, a.bin_size, a.chunk_size);
                    if (l < a.n_lines) atomicAdd(&a.line_counts[l * a.S + sg], 1u);
                }
            }
        }
        pw = npw; pt = npt; vw = nvw;
        pos0 += stride;
        bin += dbin;
        brem += dbrem;
        if (brem >= a.bin_size) { brem -= a.bin_size; bin++; }
        if (a.chunk_size) {
            chk += dchk;
            crem += dcrem;
            if (crem >= a.chunk_size) { crem -= a.chunk_size; chk++; }
        }
    }
    n_hit = spk_warp_sum_u64(n_hit);
    if (lane == 0 && n_hit && a.nhits) atomicAdd((unsigned long long*)a.nhits, (unsigned long long)n_hit);
}

// Circos._bed_density(stack=True) (Circos.py:737-741): window row += bin row
__global__ void __launch_bounds__(256)
k_stack_windows(const int64_t* __restrict__ line_counts, const uint32_t* __restrict__ line_window,
                uint64_t n_lines, int S, int64_t* __restrict__ out) {
    for (uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n_lines * (uint64_t)S;
         e += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t l = e / S;
        const int c = (int)(e - l * S);
        const long long v = line_counts[e];
        if (v) atomicAdd((unsigned long long*)&out[(uint64_t)line_window[l] * S + c], (unsigned long long)v);
    }
}

// Device-side stack of the (bin, chunk) lines of one chromosome: line l covers positions starting at
// p_l = min{p : line_of(p) >= l}; its printed start is (p_l / bin_size) * bin_size and the reference
// assigns it to window start // window_size (Circos.py:732).
__global__ void __launch_bounds__(256)
k_stack_lines(const uint32_t* __restrict__ line_counts, uint64_t n_lines, int S, int k, uint64_t bin_size,
              uint64_t chunk_size, uint64_t window_size, uint64_t n_bases, int64_t* __restrict__ out,
              uint64_t n_windows) {
    const uint64_t l = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= n_lines) return;
    bool any = false;
    for (int c = 0; c < S; c++) any |= line_counts[l * S + c] != 0;
    if (!any) return;
    uint64_t lo = 0, hi = n_bases;  // first p with line_of(p) >= l
    while (lo < hi) {
        const uint64_t mid = lo + ((hi - lo) >> 1);
        if (line_of(mid, k, bin_size, chunk_size) >= l) hi = mid;
        else lo = mid + 1;
    }
    const uint64_t w = ((lo / bin_size) * bin_size) / window_size;
    if (w >= n_windows) return;
    for (int c = 0; c < S; c++) {
        const uint32_t v = line_counts[l * S + c];
        if (v) atomicAdd((unsigned long long*)&out[w * S + c], (unsigned long long)v);
    }
}

}  // namespace

extern "C" int spk_stack_lines(const uint32_t* d_line_counts, uint64_t n_lines, int S, int k,
                               uint64_t bin_size, uint64_t chunk_size, uint64_t window_size,
                               uint64_t n_bases, int64_t* d_out, uint64_t n_windows, void* stream) {
    SPK_CHECK_ARG(S >= 1 && bin_size >= 1 && window_size >= 1, "bad arguments");
    if (n_lines == 0) return SPK_OK;
    SPK_CHECK_ARG(d_line_counts && d_out, "null pointer");
    k_stack_lines<<<(unsigned)((n_lines + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        d_line_counts, n_lines, S, k, bin_size, chunk_size, window_size, n_bases, d_out, n_windows);
    SPK_LAUNCH_CHECK();
    return SPK_OK;
}

extern "C" int spk_stack_windows(const int64_t* d_line_counts, const uint32_t* d_line_window,
                                 uint64_t n_lines, int S, int64_t* d_out, void* stream) {
    SPK_CHECK_ARG(S >= 1, "S must be >= 1");
    if (n_lines == 0) return SPK_OK;
    SPK_CHECK_ARG(d_line_counts && d_line_window && d_out, "null pointer");
    const uint64_t blocks = (n_lines * (uint64_t)S + 255) / 256;
    const unsigned grid = (unsigned)min(blocks, (uint64_t)spk_num_sms() * 16);
    k_stack_windows<<<grid, 256, 0, (cudaStream_t)stream>>>(d_line_counts, d_line_window, n_lines, S, d_out);
    SPK_LAUNCH_CHECK();
    return SPK_OK;
}

extern "C" int spk_sig_table_build(const uint64_t* d_keys, const uint8_t* d_vals, uint64_t n,
                                   uint64_t* d_skeys, uint8_t* d_svals, uint64_t sslots,
                                   uint32_t* d_filter, uint64_t filter_bits, int pack_vals,
                                   uint64_t* d_fail, void* stream) {
    SPK_CHECK_ARG(d_skeys && d_svals && d_fail, "null pointer");
    SPK_CHECK_ARG(sslots >= 2, "sslots too small");
    SPK_CHECK_ARG(!d_filter || (filter_bits >= 32 && (filter_bits & (filter_bits - 1)) == 0),
                  "filter_bits must be a power of two >= 32");
    if (n == 0) return SPK_OK;
    SPK_CHECK_ARG(d_keys && d_vals, "null keys/vals");
    const uint64_t blocks = (n + 255) / 256;
    const unsigned grid = (unsigned)min(blocks, (uint64_t)spk_num_sms() * 16);
    k_sig_build<<<grid, 256, 0, (cudaStream_t)stream>>>(d_keys, d_vals, n, d_skeys, d_svals, sslots,
                                                        d_filter, d_filter ? filter_bits - 1 : 0, pack_vals, d_fail);
    SPK_LAUNCH_CHECK();
    return SPK_OK;
}

extern "C" uint64_t spk_map_num_lines(uint64_t n_bases, int k, uint64_t bin_size, uint64_t chunk_size) {
    if (bin_size == 0 || n_bases == 0) return 0;
    const uint64_t last = n_bases - 1;
    uint64_t l = last / bin_size;
    if (chunk_size) l += (last + (uint64_t)(k - 1)) / chunk_size;
    return l + 1;
}

extern "C" int spk_map_bins(const uint32_t* d_packed, const uint32_t* d_valid, uint64_t n_bases, int k,
                            const uint64_t* d_skeys, const uint8_t* d_svals, uint64_t sslots, int S,
                            const uint32_t* d_filter, uint64_t filter_bits, int pack_vals,
                            uint64_t bin_size, uint64_t chunk_size, uint32_t* d_line_counts, uint64_t n_lines,
                            uint8_t* d_hit_flags, uint64_t* d_nhits, void* stream) {
    SPK_CHECK_ARG(!d_filter || (filter_bits >= 32 && (filter_bits & (filter_bits - 1)) == 0),
                  "filter_bits must be a power of two >= 32");
    SPK_CHECK_ARG(d_packed && d_valid && d_skeys && d_svals && d_line_counts, "null pointer");
    SPK_CHECK_ARG(k >= 1 && k <= 32, "k must be in [1, 32]");
    SPK_CHECK_ARG(S >= 1 && S <= MP_MAX_S, "S must be in [1, 32]");
    SPK_CHECK_ARG(bin_size >= 1, "bin_size must be >= 1");
    SPK_CHECK_ARG(sslots >= 2, "sslots too small");
    SPK_CHECK_ARG(n_lines >= spk_map_num_lines(n_bases, k, bin_size, chunk_size), "n_lines too small");
    if (n_bases < (uint64_t)k) return SPK_OK;
    const uint64_t n_tiles = (n_bases + SPK_TILE_BASES - 1) / SPK_TILE_BASES;
    const unsigned grid = (unsigned)min((uint64_t)spk_num_sms() * 3, n_tiles);
    SPK_CHECK_ARG(!pack_vals || k <= 28, "packed values need k <= 28");
    MapArgs a{d_skeys, d_svals, sslots, d_filter, d_filter ? filter_bits - 1 : 0, pack_vals, S, bin_size, chunk_size,
              d_line_counts, n_lines, d_hit_flags, d_nhits, nullptr, nullptr, 0};
    k_map_bins<<<grid, SPK_TILE_THREADS, 0, (cudaStream_t)stream>>>(
        (const uint8_t*)d_packed, (const uint8_t*)d_valid, n_bases, k, a);
    SPK_LAUNCH_CHECK();
    return SPK_OK;
}

// ---- bucketed quotient table: plan / build / map ---------------------------------------------------------
static int qt_sgbits(int S) {
    int b = 1;
    while ((1 << b) < S + 1) b++;
    return b;
}

extern "C" int spk_qtable_plan(uint64_t n_keys, int k, int S, int* slot_bits, int* bucket_bits) {
    SPK_CHECK_ARG(slot_bits && bucket_bits, "null pointer");
    SPK_CHECK_ARG(k >= 1 && k <= 32 && S >= 1 && S <= MP_MAX_S, "bad k or S");
    const int sgbits = qt_sgbits(S);
    // mean occupancy <= 1.75 of 7 entry slots: a bucket overflows (-> stash probe, a divergent dependent load for
    // every later lookup of that bucket) with probability 3e-4; at 3.5 it was 1.6 % of the buckets, i.e. ~40 %
    // of the warp-wide lookups had a lane in the slow path
    int bn = 4;
    double mean = 1.75;
    if (const char* e = getenv("SPK_QT_MEAN")) {       // test hook: crowd the buckets to exercise the stash path
        const double v = atof(e);
        if (v > 0.0) mean = v;
    }
    while (bn < 40 && (double)n_keys / (double)(1ull << bn) > mean) bn++;
    if (bn > 2 * k) bn = 2 * k;
    int b16 = bn;
    if (2 * k + sgbits - 16 > b16) b16 = 2 * k + sgbits - 16;
    if (b16 <= 2 * k && b16 <= 22) {                   // <= 64 MB
        *slot_bits = 16;
        *bucket_bits = b16;
        return SPK_OK;
    }
    int b32 = bn;
    if (2 * k + sgbits - 32 > b32) b32 = 2 * k + sgbits - 32;
    if (b32 <= 2 * k && b32 <= 24) {                   // <= 512 MB
        *slot_bits = 32;
        *bucket_bits = b32;
        return SPK_OK;
    }
    spk_set_error("spk_qtable_plan: no bucketed layout for k=%d, S=%d, %llu keys", k, S, (unsigned long long)n_keys);
    return SPK_EINVAL;
}

extern "C" int spk_qtable_build(const uint64_t* d_keys, const uint8_t* d_vals, uint64_t n, int k, int S,
                                void* d_buckets, int slot_bits, int bucket_bits, uint64_t* d_skeys,
                                uint8_t* d_svals, uint64_t sslots, int pack_vals, uint64_t* d_fail, void* stream) {
    SPK_CHECK_ARG(d_buckets && d_skeys && d_svals && d_fail, "null pointer");
    SPK_CHECK_ARG(slot_bits == 16 || slot_bits == 32, "slot_bits must be 16 or 32");
    SPK_CHECK_ARG(k >= 1 && k <= 32 && bucket_bits >= 0 && bucket_bits <= 2 * k && bucket_bits <= 30, "bad geometry");
    SPK_CHECK_ARG(2 * k - bucket_bits + qt_sgbits(S) <= slot_bits, "remainder does not fit the slot");
    SPK_CHECK_ARG(sslots >= 2, "stash too small");
    SPK_CHECK_ARG(!pack_vals || k <= 28, "packed values need k <= 28");
    if (n == 0) return SPK_OK;
    SPK_CHECK_ARG(d_keys && d_vals, "null keys/vals");
    const Mixer mx = spk_make_mixer(k, bucket_bits);
    const uint64_t blocks = (n + 255) / 256;
    const unsigned grid = (unsigned)min(blocks, (uint64_t)spk_num_sms() * 16);
    if (slot_bits == 16)
        k_qt_build<unsigned short><<<grid, 256, 0, (cudaStream_t)stream>>>(
            d_keys, d_vals, n, (unsigned short*)d_buckets, qt_sgbits(S), mx, d_skeys, d_svals, sslots, pack_vals, d_fail);
    else
        k_qt_build<unsigned int><<<grid, 256, 0, (cudaStream_t)stream>>>(
            d_keys, d_vals, n, (unsigned int*)d_buckets, qt_sgbits(S), mx, d_skeys, d_svals, sslots, pack_vals, d_fail);
    SPK_LAUNCH_CHECK();
    return SPK_OK;
}

extern "C" int spk_map_bins_q(const uint32_t* d_packed, const uint32_t* d_valid, uint64_t n_bases, int k,
                              const void* d_buckets, int slot_bits, int bucket_bits, const uint64_t* d_skeys,
                              const uint8_t* d_svals, uint64_t sslots, int pack_vals, int S, uint64_t bin_size,
                              uint64_t chunk_size, uint32_t* d_line_counts, uint64_t n_lines,
                              uint8_t* d_hit_flags, uint64_t* d_nhits, const uint64_t* d_rec_start,
                              const uint64_t* d_rec_line0, uint32_t n_rec, void* stream) {
    SPK_CHECK_ARG(d_packed && d_valid && d_buckets && d_skeys && d_svals && d_line_counts, "null pointer");
    SPK_CHECK_ARG(slot_bits == 16 || slot_bits == 32, "slot_bits must be 16 or 32");
    SPK_CHECK_ARG(k >= 1 && k <= 32 && bucket_bits >= 0 && bucket_bits <= 2 * k && bucket_bits <= 30, "bad geometry");
    SPK_CHECK_ARG(S >= 1 && S <= MP_MAX_S, "S must be in [1, 32]");
    SPK_CHECK_ARG(2 * k - bucket_bits + qt_sgbits(S) <= slot_bits, "remainder does not fit the slot");
    SPK_CHECK_ARG(bin_size >= 1, "bin_size must be >= 1");
    SPK_CHECK_ARG(sslots >= 2, "stash too small");
    SPK_CHECK_ARG(!pack_vals || k <= 28, "packed values need k <= 28");
    SPK_CHECK_ARG(d_rec_start || n_lines >= spk_map_num_lines(n_bases, k, bin_size, chunk_size), "n_lines too small");
    SPK_CHECK_ARG(!d_rec_start || (d_rec_line0 && n_rec >= 1), "record index incomplete");
    if (n_bases < (uint64_t)k) return SPK_OK;
    const uint64_t n_tiles = (n_bases + SPK_TILE_BASES - 1) / SPK_TILE_BASES;
    const unsigned grid = (unsigned)min((uint64_t)spk_num_sms() * 3, n_tiles);
    MapArgs a{d_skeys, d_svals, sslots, nullptr, 0, pack_vals, S, bin_size, chunk_size,
              d_line_counts, n_lines, d_hit_flags, d_nhits, d_rec_start, d_rec_line0, n_rec};
    QtArgs qa{d_buckets, slot_bits, bucket_bits, qt_sgbits(S), spk_make_mixer(k, bucket_bits), d_skeys, d_svals,
              sslots, pack_vals};
    const char* mk = getenv("SPK_MAP_KERNEL");             // "tile": the CTA-tile kernel for every call (tests, A/B)
    if (slot_bits == 16 && S <= 4 && !d_rec_start && !d_hit_flags && bucket_bits < 2 * k && !(mk && mk[0] == 't')) {
        const uint64_t n_wt = (n_bases + 511) / 512;
        const unsigned gridw = (unsigned)min((uint64_t)spk_num_sms() * 8, (n_wt + MW_THREADS / 32 - 1) / (MW_THREADS / 32));
        k_map_bins_w<<<gridw, MW_THREADS, 0, (cudaStream_t)stream>>>(d_packed, d_valid, n_bases, k, a, qa);
    } else if (slot_bits == 16)
        k_map_bins_q<true><<<grid, SPK_TILE_THREADS, 0, (cudaStream_t)stream>>>(
            (const uint8_t*)d_packed, (const uint8_t*)d_valid, n_bases, k, a, qa);
    else
        k_map_bins_q<false><<<grid, SPK_TILE_THREADS, 0, (cudaStream_t)stream>>>(
            (const uint8_t*)d_packed, (const uint8_t*)d_valid, n_bases, k, a, qa);
    SPK_LAUNCH_CHECK();
    return SPK_OK;
}
