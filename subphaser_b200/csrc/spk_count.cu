// K2 + K3: canonical k-mer counting of one chromosome into an open-addressed table, and the table
// scan that turns it into `jellyfish dump -c -L lower_count` content.
//
// Replaces `jellyfish count -m K --canonical` + `jellyfish dump -c -L` (Jellyfish.py:697-700) and the
// per-line dump parse of Jellyfish.py:90-98 (lengths[i] = sum of dumped counts).
//
// Data flow of the count kernel (persistent grid, one CTA loops over 4096-base tiles):
//   HBM --(cp.async.bulk 1-D, mbarrier, double buffered)--> smem tile of 2-bit codes + validity bits
//   each thread owns 16 consecutive k-mer start positions: rolling forward / reverse-complement
//   words, canonical = min, murmur finaliser -> slot; 8 probes are loaded in one batch (ld.global.cg)
//   before any of them is resolved so every thread keeps 8 sectors in flight; equal k-mers of
//   neighbouring positions / lanes are merged first (run-merge + __match_any_sync), then one RED
//   (atomicAdd without return) or one CAS per distinct k-mer.
#include <stdlib.h>
#include "spk_common.cuh"
#include "spk_tile.cuh"

namespace {

constexpr int CT_THREADS = SPK_TILE_THREADS;
constexpr int CT_PER_THREAD = SPK_KMERS_PER_THREAD;
constexpr int CT_BATCH = 8;

struct TableView {
    uint64_t* keys;    // layout 0: packed slots; layout 1: keys
    uint32_t* counts;  // layout 1 only
    uint64_t slots;
    int cbits;         // layout 0: 64 - 2k
};

__host__ int layout_of(uint64_t n_bases, int k) {
    const int cbits = 64 - 2 * k;
    if (cbits <= 0) return 1;
    if (cbits >= 64) return 0;
    return (n_bases < (1ull << cbits)) ? 0 : 1;
}

__host__ uint64_t recommended_slots(uint64_t n_bases, int k) {
    // distinct canonical k-mers <= min(#windows, 4^k)
    uint64_t distinct = n_bases;
    if (k < 31) {
        const uint64_t space = 1ull << (2 * k);
        if (space < distinct) distinct = space;
    }
    uint64_t slots = (uint64_t)((double)distinct / 0.7) + 1024;
    return slots;
}

template <int LAYOUT>
__device__ __forceinline__ uint64_t probe_load(const TableView& t, uint64_t slot) {
    return __ldcg(t.keys + slot);
}

// Resolve one insert whose first probe value `cur` has already been loaded.  Returns false if the
// table is full.
template <int LAYOUT>
__device__ __forceinline__ bool insert_resolve(const TableView& t, uint64_t key, uint32_t add,
                                               uint64_t slot, uint64_t cur) {
    for (uint64_t probes = 0; probes < t.slots; probes++) {
        if (LAYOUT == 0) {
            if (cur == 0) {
                const uint64_t fresh = (key << t.cbits) | (uint64_t)add;
                const uint64_t old = atomicCAS((unsigned long long*)(t.keys + slot), 0ull,
                                               (unsigned long long)fresh);
                if (old == 0) return true;
                cur = old;
            }
            if ((cur >> t.cbits) == key) {
                atomicAdd((unsigned long long*)(t.keys + slot), (unsigned long long)add);
                return true;
            }
        } else {
            if (cur == SPK_EMPTY_KEY) {
                const uint64_t old = atomicCAS((unsigned long long*)(t.keys + slot),
                                               (unsigned long long)SPK_EMPTY_KEY,
                                               (unsigned long long)key);
                if (old == SPK_EMPTY_KEY) cur = key;
                else cur = old;
            }
            if (cur == key) {
                atomicAdd(t.counts + slot, add);
                return true;
            }
        }
        slot++;
        if (slot == t.slots) slot = 0;
        cur = __ldcg(t.keys + slot);
    }
    return false;
}

template <int LAYOUT, bool AGG>
__global__ void __launch_bounds__(CT_THREADS, 3)
k_count_canonical(const uint8_t* __restrict__ packed, const uint8_t* __restrict__ valid,
                  uint64_t n_bases, int k, TableView tab, uint64_t* __restrict__ stats) {
    __shared__ SpkTileSmem sm;
    const uint64_t n_tiles = (n_bases + SPK_TILE_BASES - 1) / SPK_TILE_BASES;
    const int tid = threadIdx.x;
    const SpkKmerParams kp = spk_kmer_params(k);
    spk_tile_init(sm);
    uint64_t n_valid = 0, n_fail = 0;

    uint64_t tile = blockIdx.x;
    if (tid == 0 && tile < n_tiles) spk_tile_issue(sm, packed, valid, tile, 0);

    for (uint32_t it = 0; tile < n_tiles; it++, tile += gridDim.x) {
        const int buf = it & 1;
        const uint32_t parity = (it >> 1) & 1;
        __syncthreads();  // everyone is done with buffer buf^1 (read in the previous iteration)
        if (tid == 0 && tile + gridDim.x < n_tiles)
            spk_tile_issue(sm, packed, valid, tile + gridDim.x, buf ^ 1);
        spk_mbar_wait(&sm.bar[buf], parity);

        uint64_t key[CT_PER_THREAD];
        uint32_t okmask;
        spk_tile_kmers(sm, buf, kp, key, okmask);
        n_valid += __popc(okmask);

        // ---- run-merge inside the thread: equal consecutive k-mers become one insert ----
        uint32_t mult[CT_PER_THREAD];
#pragma unroll
        for (int j = 0; j < CT_PER_THREAD; j++) mult[j] = (okmask >> j) & 1u;
#pragma unroll
        for (int j = CT_PER_THREAD - 1; j > 0; j--) {
            if (mult[j] && mult[j - 1] && key[j] == key[j - 1]) {
                mult[j - 1] += mult[j];
                mult[j] = 0;
            }
        }

        // ---- batched probe + insert ----
#pragma unroll
        for (int b0 = 0; b0 < CT_PER_THREAD; b0 += CT_BATCH) {
            uint64_t cur[CT_BATCH];
#pragma unroll
            for (int j = 0; j < CT_BATCH; j++) {
                if (AGG) {
                    // merge equal k-mers across the warp (microsatellites: period divides 16)
                    const uint64_t kk = mult[b0 + j] ? key[b0 + j] : (SPK_EMPTY_KEY - (tid & 31));
                    const uint32_t peers = __match_any_sync(0xffffffffu, kk);
                    if (mult[b0 + j] && peers != (1u << (tid & 31))) {
                        uint32_t sum = 0;
                        uint32_t m = peers;
                        const int leader = __ffs(peers) - 1;
                        // sum multiplicities of peers (few iterations; peers is warp-divergent)
                        while (m) {
                            const int l = __ffs(m) - 1;
                            m &= m - 1;
                            sum += __shfl_sync(peers, mult[b0 + j], l);
                        }
                        mult[b0 + j] = ((tid & 31) == leader) ? sum : 0;
                    }
                }
                cur[j] = mult[b0 + j]
                             ? probe_load<LAYOUT>(tab, spk_slot_of(spk_hash64(key[b0 + j]), tab.slots))
                             : 0;
            }
#pragma unroll
            for (int j = 0; j < CT_BATCH; j++) {
                if (mult[b0 + j]) {
                    const uint64_t slot = spk_slot_of(spk_hash64(key[b0 + j]), tab.slots);
                    if (!insert_resolve<LAYOUT>(tab, key[b0 + j], mult[b0 + j], slot, cur[j]))
                        n_fail += mult[b0 + j];
                }
            }
        }
    }

    n_valid = spk_warp_sum_u64(n_valid);
    n_fail = spk_warp_sum_u64(n_fail);
    if ((tid & 31) == 0) {
        if (n_valid) atomicAdd((unsigned long long*)&stats[0], (unsigned long long)n_valid);
        if (n_fail) atomicAdd((unsigned long long*)&stats[1], (unsigned long long)n_fail);
    }
}

// ---- table scan -------------------------------------------------------------------------------------
constexpr int SC_THREADS = 256;

__device__ __forceinline__ bool slot_entry(const TableView& t, int layout, uint64_t i, uint64_t& key,
                                           uint64_t& cnt) {
    if (layout == 0) {
        const uint64_t v = __ldcs(t.keys + i);
        if (v == 0) return false;
        key = v >> t.cbits;
        cnt = v & ((1ull << t.cbits) - 1);
        return true;
    }
    const uint64_t kk = __ldcs(t.keys + i);
    if (kk == SPK_EMPTY_KEY) return false;
    key = kk;
    cnt = __ldcs(t.counts + i);
    return true;
}

__global__ void __launch_bounds__(SC_THREADS)
k_table_stats(TableView tab, int layout, uint32_t lower, uint64_t* __restrict__ out,
              uint32_t* __restrict__ block_counts, uint64_t* __restrict__ histo, uint32_t histo_len) {
    __shared__ uint64_t s_red[4][SC_THREADS / 32];
    __shared__ uint32_t s_hist[256];
    const uint64_t per_block = (tab.slots + gridDim.x - 1) / gridDim.x;
    const uint64_t beg = (uint64_t)blockIdx.x * per_block;
    const uint64_t end = min(beg + per_block, tab.slots);
    uint64_t distinct = 0, nge = 0, sumge = 0, sumall = 0;
    uint32_t h1 = 0, h2 = 0;  // counts 1 and 2 dominate: keep them in registers
    if (histo) {
        s_hist[threadIdx.x] = 0;
        __syncthreads();
    }
    for (uint64_t i = beg + threadIdx.x; i < end; i += SC_THREADS) {
        uint64_t key, cnt;
        if (slot_entry(tab, layout, i, key, cnt)) {
            distinct++;
            sumall += cnt;
            if (cnt >= lower) {
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
        for (int w = 0; w < SC_THREADS / 32; w++) s += s_red[threadIdx.x][w];
        if (s) atomicAdd((unsigned long long*)&out[threadIdx.x], (unsigned long long)s);
        if (threadIdx.x == 1) block_counts[blockIdx.x] = (uint32_t)s;
    }
}

// in-place exclusive scan of the per-block survivor counts (<= a few thousand blocks)
__global__ void __launch_bounds__(1024) k_scan_blocks(uint32_t* counts, int n, uint64_t* offsets) {
    __shared__ uint64_t s_warp[32];
    __shared__ uint64_t s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < n; base += 1024) {
        const int i = base + threadIdx.x;
        const uint64_t v = (i < n) ? counts[i] : 0;
        uint64_t incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint64_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if ((threadIdx.x & 31) >= o) incl += t;
        }
        if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = incl;
        __syncthreads();
        uint64_t prefix = s_carry;
        for (int w = 0; w < (int)(threadIdx.x >> 5); w++) prefix += s_warp[w];
        if (i < n) offsets[i] = prefix + incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = prefix + incl;
        __syncthreads();
    }
}

// Survivors of each block are written in slot order: per 256-slot step a CTA-wide exclusive scan of
// the keep flags gives deterministic positions.
__global__ void __launch_bounds__(SC_THREADS)
k_table_extract(TableView tab, int layout, uint32_t lower, const uint64_t* __restrict__ block_off,
                uint64_t* __restrict__ out_keys, uint32_t* __restrict__ out_counts, uint64_t cap) {
    __shared__ uint32_t s_warp[SC_THREADS / 32];
    const uint64_t per_block = (tab.slots + gridDim.x - 1) / gridDim.x;
    const uint64_t beg = (uint64_t)blockIdx.x * per_block;
    const uint64_t end = min(beg + per_block, tab.slots);
    uint64_t wr = block_off[blockIdx.x];
    for (uint64_t base = beg; base < end; base += SC_THREADS) {
        const uint64_t i = base + threadIdx.x;
        uint64_t key = 0, cnt = 0;
        bool keep = false;
        if (i < end) keep = slot_entry(tab, layout, i, key, cnt) && cnt >= lower;
        const uint32_t ballot = __ballot_sync(0xffffffffu, keep);
        const uint32_t lane = threadIdx.x & 31;
        const uint32_t wprefix = __popc(ballot & ((1u << lane) - 1));
        if (lane == 0) s_warp[threadIdx.x >> 5] = __popc(ballot);
        __syncthreads();
        uint32_t prefix = 0, total = 0;
#pragma unroll
        for (int w = 0; w < SC_THREADS / 32; w++) {
            if (w < (int)(threadIdx.x >> 5)) prefix += s_warp[w];
            total += s_warp[w];
        }
        if (keep) {
            const uint64_t o = wr + prefix + wprefix;
            if (o < cap) {
                out_keys[o] = key;
                out_counts[o] = (uint32_t)cnt;
            }
        }
        wr += total;
        __syncthreads();
    }
}

__host__ int make_view(void* d_table, size_t table_bytes, int k, int layout, TableView* tv) {
    const size_t slot_bytes = (layout == 0) ? 8 : 12;
    const uint64_t slots = table_bytes / slot_bytes;
    if (slots < 2) return SPK_EINVAL;
    tv->keys = (uint64_t*)d_table;
    tv->counts = (layout == 0) ? nullptr : (uint32_t*)((char*)d_table + slots * 8);
    tv->slots = slots;
    tv->cbits = 64 - 2 * k;
    return SPK_OK;
}

#define SPK_TABLE_ARGS_CHECK()                                                       \
    SPK_CHECK_ARG(k >= 1 && k <= 32, "k must be in [1, 32]");                        \
    SPK_CHECK_ARG(layout == 0 || layout == 1, "layout must be 0 or 1");              \
    SPK_CHECK_ARG(layout == 1 || k < 32, "packed layout needs k < 32");              \
    TableView tv;                                                                    \
    if (make_view((void*)d_table, table_bytes, k, layout, &tv) != SPK_OK) {          \
        spk_set_error("%s: table too small", __func__);                              \
        return SPK_EINVAL;                                                           \
    }

}  // namespace

extern "C" int spk_count_layout(uint64_t n_bases, int k) { return layout_of(n_bases, k); }

extern "C" size_t spk_count_table_bytes(uint64_t n_bases, int k) {
    const uint64_t slots = recommended_slots(n_bases, k);
    const size_t slot_bytes = (layout_of(n_bases, k) == 0) ? 8 : 12;
    return (size_t)((slots * slot_bytes + 255) / 256 * 256);
}

extern "C" uint64_t spk_count_table_slots(size_t table_bytes, int layout) {
    return table_bytes / ((layout == 0) ? 8 : 12);
}

extern "C" int spk_count_table_init(void* d_table, size_t table_bytes, int k, int layout, void* stream) {
    SPK_CHECK_ARG(d_table, "null table");
    SPK_TABLE_ARGS_CHECK();
    cudaStream_t st = (cudaStream_t)stream;
    if (layout == 0) {
        SPK_CUDA(cudaMemsetAsync(d_table, 0, tv.slots * 8, st));
    } else {
        SPK_CUDA(cudaMemsetAsync(tv.keys, 0xFF, tv.slots * 8, st));
        SPK_CUDA(cudaMemsetAsync(tv.counts, 0, tv.slots * 4, st));
    }
    return SPK_OK;
}

static int count_agg_enabled() {
    static int cached = -1;
    if (cached < 0) {
        const char* e = getenv("SPK_COUNT_AGG");
        cached = (e && e[0] == '0') ? 0 : 1;
    }
    return cached;
}

extern "C" int spk_count_canonical(const uint32_t* d_packed, const uint32_t* d_valid, uint64_t n_bases,
                                   int k, void* d_table, size_t table_bytes, int layout,
                                   uint64_t* d_stats, void* stream) {
    SPK_CHECK_ARG(d_packed && d_valid && d_table && d_stats, "null pointer");
    SPK_CHECK_ARG(((uintptr_t)d_packed & 15) == 0 && ((uintptr_t)d_valid & 15) == 0,
                  "sequence buffers must be 16-byte aligned");
    SPK_TABLE_ARGS_CHECK();
    if (layout == 0 && layout_of(n_bases, k) != 0) {
        spk_set_error("spk_count_canonical: %llu bases can overflow the %d-bit packed count field",
                      (unsigned long long)n_bases, 64 - 2 * k);
        return SPK_EOVERFLOW;
    }
    if (n_bases < (uint64_t)k) return SPK_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const uint64_t n_tiles = (n_bases + SPK_TILE_BASES - 1) / SPK_TILE_BASES;
    const unsigned grid = (unsigned)min((uint64_t)spk_num_sms() * 3, n_tiles);
    const uint8_t* pk = (const uint8_t*)d_packed;
    const uint8_t* vl = (const uint8_t*)d_valid;
    const bool agg = count_agg_enabled();
    if (layout == 0) {
        if (agg) k_count_canonical<0, true><<<grid, CT_THREADS, 0, st>>>(pk, vl, n_bases, k, tv, d_stats);
        else k_count_canonical<0, false><<<grid, CT_THREADS, 0, st>>>(pk, vl, n_bases, k, tv, d_stats);
    } else {
        if (agg) k_count_canonical<1, true><<<grid, CT_THREADS, 0, st>>>(pk, vl, n_bases, k, tv, d_stats);
        else k_count_canonical<1, false><<<grid, CT_THREADS, 0, st>>>(pk, vl, n_bases, k, tv, d_stats);
    }
    SPK_LAUNCH_CHECK();
    return SPK_OK;
}

extern "C" int spk_table_scan_blocks(void) { return spk_num_sms() * 8; }

extern "C" int spk_table_stats(const void* d_table, size_t table_bytes, int k, int layout,
                               uint32_t lower_count, uint64_t* d_out, uint32_t* d_block_counts,
                               uint64_t* d_histo, uint32_t histo_len, void* stream) {
    SPK_CHECK_ARG(d_table && d_out && d_block_counts, "null pointer");
    SPK_CHECK_ARG(!d_histo || histo_len >= 2, "histo_len must be >= 2");
    SPK_TABLE_ARGS_CHECK();
    cudaStream_t st = (cudaStream_t)stream;
    const int blocks = spk_table_scan_blocks();
    SPK_CUDA(cudaMemsetAsync(d_out, 0, 4 * sizeof(uint64_t), st));
    k_table_stats<<<blocks, SC_THREADS, 0, st>>>(tv, layout, lower_count, d_out, d_block_counts,
                                                 d_histo, histo_len);
    SPK_LAUNCH_CHECK();
    return SPK_OK;
}

extern "C" int spk_table_extract(const void* d_table, size_t table_bytes, int k, int layout,
                                 uint32_t lower_count, uint32_t* d_block_counts, uint64_t* d_keys,
                                 uint32_t* d_counts, uint64_t cap, void* stream) {
    SPK_CHECK_ARG(d_table && d_block_counts && d_keys && d_counts, "null pointer");
    SPK_TABLE_ARGS_CHECK();
    cudaStream_t st = (cudaStream_t)stream;
    const int blocks = spk_table_scan_blocks();
    // the uint64 offsets live behind the uint32 counts (d_block_counts holds 3*blocks+2 uint32)
    uint64_t* offsets = (uint64_t*)(d_block_counts + ((blocks + 1) & ~1));
    k_scan_blocks<<<1, 1024, 0, st>>>(d_block_counts, blocks, offsets);
    SPK_LAUNCH_CHECK();
    k_table_extract<<<blocks, SC_THREADS, 0, st>>>(tv, layout, lower_count, offsets, d_keys, d_counts,
                                                   cap);
    SPK_LAUNCH_CHECK();
    return SPK_OK;
}
