// K3b+K4 (partitioned): union of the per-chromosome dumps -> count rows -> differential filter, one hash
// partition at a time, entirely in shared memory.
//
// Replaces JellyfishDumps.to_matrix + JellyfishDumps.filter / _filter_kmer (Jellyfish.py:439-512,611-648)
// for dumps produced by spk_pcount_canonical_ex with a COMMON number of partition bits: partition p of
// every chromosome holds exactly the k-mers u with (f(u) >> rbits) == p, and its dump entries are
// contiguous (pindex).  So the union of partition p over the n chromosomes is a few hundred rows: they are
// merged in a shared-memory table (key -> n counters) and every row goes through the exact integer
// pre-screen of the differential test (spk_filter_prescreen: a row with too many all-zero homoeologous sets
// cannot pass the fold test).  Only the candidates (typically ~2 % of the union) are written to HBM as a
// compact count matrix, on which the ordinary filter kernels (spk_filter_differential ... spk_filter_emit)
// then run at full occupancy — the fp64 work of a handful of candidates per partition would otherwise
// serialise whole CTAs behind two or three active lanes.  The multi-GB union table and the [U x n] matrix
// of the plain path (spk_matrix.cu: ~6 random HBM accesses per dump entry) never exist.
#include <stdlib.h>
#include "spk_common.cuh"
#include "spk_filter.cuh"

namespace {

constexpr int PM_THREADS = 512;
constexpr int PM_MAX_COLS = 512;
constexpr int PM_B = 4;              // dump entries in flight per thread

struct PmArgs {
    const uint64_t* const* keys;      // [n] device pointers to the dump keys of every chromosome
    const uint32_t* const* counts;    // [n]
    const uint32_t* const* pindex;    // [n] uint32[2P]: first entry / number of entries of partition p
    int n;
    uint64_t P;
    uint32_t nparts, part;            // this call handles partitions p % nparts == part
    const uint64_t* lengths;
    FilterCfg cfg;
    int n_groups;
    int union_only;                   // 1: count the union rows only (no filter, nothing written)
    uint64_t* out_keys;               // candidate rows (integer pre-screen passed), arbitrary order
    uint32_t* out_counts;             // [cap x n]
    uint64_t cap;
    uint64_t* counters;               // [0] union rows [2] candidate rows [3] table overflows [4] rows written
    uint32_t tslots;                  // table slots (power of two)
};

__device__ __forceinline__ uint32_t pm_hash(uint64_t key) { return (uint32_t)(spk_hash64(key) >> 32); }

__global__ void __launch_bounds__(PM_THREADS, 1) k_pmatrix_filter(PmArgs a, const uint64_t* run_if) {
    extern __shared__ __align__(16) uint8_t s_raw[];
    if (run_if && *run_if == 0) return;                  // (k_pmatrix_filter2 took the whole call)
    const int n = a.n;
    const uint32_t TS = a.tslots, TM = TS - 1;
    uint64_t* s_key = (uint64_t*)s_raw;                        // [TS]
    uint32_t* s_cnt = (uint32_t*)(s_key + TS);                 // [TS x n]
    uint32_t* s_start = s_cnt + (size_t)TS * n;                // [n]   first entry of this partition
    uint32_t* s_off = s_start + n;                             // [n+1] prefix of entry counts
    uint16_t* s_rows = (uint16_t*)(s_off + n + 2);             // [TS]  occupied slots of the current round
    uint8_t* s_cfg = (uint8_t*)(s_rows + TS);                  // staged filter configuration (8-byte aligned)
    __shared__ uint32_t s_fail, s_nrows;
    const uint64_t* lengths = a.lengths;
    FilterCfg cfg = a.cfg;
    if (!a.union_only) cfg = spk_filter_stage(a.cfg, a.n_groups, n, a.lengths, s_cfg, &lengths);
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    for (uint32_t i = tid; i < TS; i += PM_THREADS) s_key[i] = SPK_EMPTY_KEY;
    for (uint32_t i = tid; i < TS * (uint32_t)n; i += PM_THREADS) s_cnt[i] = 0;
    if (tid == 0) {
        s_fail = 0;
        s_nrows = 0;
    }
    uint64_t n_union = 0, n_keep = 0;
    const uint64_t stride = (uint64_t)gridDim.x * a.nparts;
    uint64_t p = (uint64_t)blockIdx.x * a.nparts + a.part;
    // index of the CTA's next partition, prefetched while the current one is processed (threads c < n)
    uint32_t nx_start = 0, nx_len = 0;
    if (p < a.P)
        for (int c = tid; c < n; c += PM_THREADS) {          // n <= PM_MAX_COLS = blockDim: one c per thread
            nx_start = a.pindex[c][2 * p];
            nx_len = a.pindex[c][2 * p + 1];
        }
    __syncthreads();
    for (; p < a.P; p += stride) {
        if (tid < n) {
            s_start[tid] = nx_start;
            s_off[tid + 1] = nx_len;
        }
        const uint64_t pn = p + stride;
        if (pn < a.P && tid < n) {
            nx_start = a.pindex[tid][2 * pn];
            nx_len = a.pindex[tid][2 * pn + 1];
        }
        __syncthreads();
        if (tid == 0) {
            uint32_t run = 0;
            s_off[0] = 0;
            for (int c = 0; c < n; c++) {
                const uint32_t l = s_off[c + 1];
                s_off[c + 1] = run + l;
                run += l;
            }
        }
        __syncthreads();
        const uint32_t E = s_off[n];
        // rows <= E: when the partition could crowd the table it is processed in `rounds` key-hash classes
        uint32_t rounds = 1;
        while ((uint64_t)E > (uint64_t)rounds * (TS - TS / 8)) rounds <<= 1;
        for (uint32_t rd = 0; rd < rounds; rd++) {
            // ---- insert: (key, count of chromosome c) -> table row.  Loads are issued in batches of PM_B per
            //      thread before any of them is consumed (the entries of a partition are 2n short runs scattered
            //      over the dumps: their DRAM latency must overlap) ----
            for (uint32_t e0 = 0; e0 < E; e0 += PM_B * PM_THREADS) {
                uint64_t key[PM_B];
                uint32_t cnt[PM_B];
                int col[PM_B];
                int c = 0;
#pragma unroll
                for (int u = 0; u < PM_B; u++) {
                    const uint32_t e = e0 + u * PM_THREADS + tid;
                    col[u] = -1;
                    if (e < E) {
                        while (e >= s_off[c + 1]) c++;
                        const uint32_t i = s_start[c] + (e - s_off[c]);
                        key[u] = __ldg(a.keys[c] + i);
                        cnt[u] = __ldg(a.counts[c] + i);
                        col[u] = c;
                    }
                }
#pragma unroll
                for (int u = 0; u < PM_B; u++) {
                    if (col[u] < 0) continue;
                    const uint32_t h = pm_hash(key[u]);
                    if (rounds > 1 && ((h >> 20) & (rounds - 1)) != rd) continue;
                    uint32_t s = h & TM;
                    bool done = false;
                    for (uint32_t pr = 0; pr < TS; pr++) {
                        uint64_t cur = s_key[s];
                        if (cur == SPK_EMPTY_KEY) {
                            cur = atomicCAS((unsigned long long*)&s_key[s], (unsigned long long)SPK_EMPTY_KEY,
                                            (unsigned long long)key[u]);
                            if (cur == SPK_EMPTY_KEY) {
                                cur = key[u];
                                s_rows[atomicAdd(&s_nrows, 1u)] = (uint16_t)s;     // new row
                            }
                        }
                        if (cur == key[u]) {
                            s_cnt[(size_t)s * n + col[u]] = cnt[u];   // one entry per (k-mer, chromosome)
                            done = true;
                            break;
                        }
                        s = (s + 1) & TM;
                    }
                    if (!done) s_fail = 1;
                }
            }
            __syncthreads();
            // ---- filter every row (dense list of occupied slots), emit the survivors, clear the table ----
            const uint32_t nrows = s_nrows;
            if (tid == 0) n_union += nrows;
            for (uint32_t r0 = 0; r0 < nrows; r0 += PM_THREADS) {
                const uint32_t r = r0 + tid;
                const bool occ = r < nrows;
                const uint32_t s = occ ? (uint32_t)s_rows[r] : 0u;
                bool cand = false;
                uint64_t key = 0;
                if (occ) {
                    key = s_key[s];
                    if (!a.union_only) cand = spk_filter_prescreen(s_cnt + (size_t)s * n, cfg);
                    n_keep += cand ? 1 : 0;
                }
                const uint32_t bk = __ballot_sync(0xffffffffu, cand);
                if (bk) {
                    uint64_t wb = 0;
                    if (lane == 0) wb = atomicAdd((unsigned long long*)&a.counters[4], (unsigned long long)__popc(bk));
                    wb = __shfl_sync(0xffffffffu, wb, 0);
                    const uint64_t at = wb + __popc(bk & ((1u << lane) - 1));
                    if (cand && at < a.cap) {
                        a.out_keys[at] = key;
                        for (int cc = 0; cc < n; cc++) a.out_counts[at * n + cc] = s_cnt[(size_t)s * n + cc];
                    }
                }
                if (occ) {
                    s_key[s] = SPK_EMPTY_KEY;
                    for (int cc = 0; cc < n; cc++) s_cnt[(size_t)s * n + cc] = 0;
                }
            }
            __syncthreads();
            if (tid == 0) s_nrows = 0;
            __syncthreads();
        }
    }
    n_union = spk_warp_sum_u64(n_union);
    n_keep = spk_warp_sum_u64(n_keep);
    if (lane == 0) {
        if (n_union) atomicAdd((unsigned long long*)&a.counters[0], (unsigned long long)n_union);
        if (n_keep) atomicAdd((unsigned long long*)&a.counters[2], (unsigned long long)n_keep);
    }
    __syncthreads();
    if (tid == 0 && s_fail) atomicAdd((unsigned long long*)&a.counters[3], 1ull);
    (void)lengths;
}


// ---- k_pmatrix_filter2: presence masks first, count rows only for candidates ------------------------------------
// The integer pre-screen only asks WHICH homoeologous sets have a nonzero count, and a dumped count is never zero:
// the decision depends on the set of chromosomes a k-mer was dumped by.  So the table keeps, per k-mer, a 64-bit
// mask of the (multi-group) sets that contain one of its chromosomes — 16 bytes per slot instead of 8 + 4n — and
// only the ~2 % of rows whose mask reaches `min_include` sets get a count row, filled from the entries the threads
// still hold in registers.  No n-wide zero / test / clear loop runs over the union rows any more, the table (4096
// slots) takes a whole partition in one round, and two CTAs share an SM (k_pmatrix_filter: one, 190 KB).
// What it cannot take — a partition with more entries than the threads hold, more candidates than its row buffer,
// more than 64 sets — it reports in counters[5]; k_pmatrix_filter then redoes the call (decided on the device).
constexpr int PM2_THREADS = 512;
constexpr int PM2_B = 6;                 // dump entries per thread held in registers (3072 per CTA)
constexpr int PM2_TS = 4096;             // table slots
constexpr int PM2_MAXC = 192;            // candidate rows per partition

__global__ void __launch_bounds__(PM2_THREADS, 2) k_pmatrix_filter2(PmArgs a) {
    extern __shared__ __align__(16) uint8_t s_raw[];
    const int n = a.n;
    constexpr uint32_t TS = PM2_TS, TM = TS - 1;
    constexpr uint32_t NONE = 0xffffffffu, PENDING = 0xfffffffeu;
    uint64_t* s_key = (uint64_t*)s_raw;                               // [TS]
    unsigned long long* s_mask = (unsigned long long*)(s_key + TS);   // [TS] sets present
    uint64_t* s_colmask = (uint64_t*)(s_mask + TS);                   // [n]  sets that contain chromosome c
    uint64_t* s_ckey = s_colmask + n;                                 // [MAXC]
    const uint64_t** s_pk = (const uint64_t**)(s_ckey + PM2_MAXC);    // [n] dump pointers (no double indirection per entry)
    const uint32_t** s_pc = (const uint32_t**)(s_pk + n);             // [n]
    uint32_t* s_rowid = (uint32_t*)(s_pc + n);                        // [TS] candidate row of a slot, or NONE
    uint32_t* s_crow = s_rowid + TS;                                  // [MAXC x n]
    uint32_t* s_start = s_crow + (size_t)PM2_MAXC * n;                // [n]
    uint32_t* s_off = s_start + n;                                    // [n + 1]
    uint8_t* s_cfg = (uint8_t*)(((uintptr_t)(s_off + n + 2) + 7) & ~(uintptr_t)7);
    __shared__ uint32_t s_fail, s_ncand, s_need;
    __shared__ unsigned long long s_wbase;
    const uint64_t* lengths = a.lengths;
    FilterCfg cfg = a.cfg;
    const int tid = threadIdx.x, lane = tid & 31;
    if (!a.union_only) cfg = spk_filter_stage(a.cfg, a.n_groups, n, a.lengths, s_cfg, &lengths);
    for (int c = tid; c < n; c += PM2_THREADS) {
        s_colmask[c] = 0;
        s_pk[c] = a.keys[c];
        s_pc[c] = a.counts[c];
    }
    for (uint32_t i = tid; i < TS; i += PM2_THREADS) {
        s_key[i] = SPK_EMPTY_KEY;
        s_mask[i] = 0ull;
        s_rowid[i] = NONE;
    }
    for (uint32_t i = tid; i < (uint32_t)PM2_MAXC * n; i += PM2_THREADS) s_crow[i] = 0;
    if (tid == 0) {
        s_fail = 0;
        s_ncand = 0;
    }
    __syncthreads();
    if (tid == 0) {
        uint32_t need = 0;
        if (!a.union_only) {
            int bit = 0;
            for (int st = 0; st < cfg.n_sets; st++) {
                const int g0 = cfg.set_off[st], g1 = cfg.set_off[st + 1];
                if (g1 - g0 < 2) continue;
                for (int m = cfg.grp_off[g0]; m < cfg.grp_off[g1]; m++) s_colmask[cfg.members[m]] |= 1ull << bit;
                bit++;
            }
            need = (uint32_t)max(cfg.min_include, 0);
        }
        s_need = need;
    }
    uint64_t n_union = 0, n_keep = 0;
    const uint64_t stride = (uint64_t)gridDim.x * a.nparts;
    uint64_t p = (uint64_t)blockIdx.x * a.nparts + a.part;
    uint32_t nx_start = 0, nx_len = 0;
    if (p < a.P && tid < n) {
        nx_start = a.pindex[tid][2 * p];
        nx_len = a.pindex[tid][2 * p + 1];
    }
    __syncthreads();
    const uint32_t need = s_need;
    for (; p < a.P; p += stride) {
        if (tid < n) {
            s_start[tid] = nx_start;
            s_off[tid + 1] = nx_len;
        }
        const uint64_t pn = p + stride;
        if (pn < a.P && tid < n) {
            nx_start = a.pindex[tid][2 * pn];
            nx_len = a.pindex[tid][2 * pn + 1];
        }
        __syncthreads();
        if (tid < 32) {                                    // exclusive prefix of the n run lengths
            uint32_t carry = 0;
            for (int base = 0; base < n; base += 32) {
                const int c = base + lane;
                const uint32_t l = c < n ? s_off[c + 1] : 0u;
                uint32_t incl = l;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += t;
                }
                if (c < n) s_off[c + 1] = carry + incl;
                carry += __shfl_sync(0xffffffffu, incl, 31);
            }
            if (lane == 0) s_off[0] = 0;
        }
        __syncthreads();
        const uint32_t E = s_off[n];
        if (E > (uint32_t)(PM2_B * PM2_THREADS)) {          // (block-uniform) not a partition for this kernel
            if (tid == 0) s_fail = 1;
            continue;
        }
        uint64_t key[PM2_B];
        uint32_t cnt[PM2_B], slot[PM2_B];
        int col[PM2_B];
        // ---- pass 1: key -> slot, OR the chromosome's set mask into the slot ----
        {
            int c = 0;
#pragma unroll
            for (int u = 0; u < PM2_B; u++) {
                const uint32_t e = u * PM2_THREADS + tid;
                col[u] = -1;
                if (e < E) {
                    while (e >= s_off[c + 1]) c++;
                    const uint32_t i = s_start[c] + (e - s_off[c]);
                    key[u] = __ldg(s_pk[c] + i);
                    cnt[u] = __ldg(s_pc[c] + i);
                    col[u] = c;
                }
            }
        }
#pragma unroll
        for (int u = 0; u < PM2_B; u++) {
            if (col[u] < 0) continue;
            uint32_t sl = pm_hash(key[u]) & TM;
            bool done = false;
            for (uint32_t pr = 0; pr < TS; pr++) {
                uint64_t cur = s_key[sl];
                if (cur == SPK_EMPTY_KEY) {
                    cur = atomicCAS((unsigned long long*)&s_key[sl], (unsigned long long)SPK_EMPTY_KEY,
                                    (unsigned long long)key[u]);
                    if (cur == SPK_EMPTY_KEY) {
                        cur = key[u];
                        n_union++;
                    }
                }
                if (cur == key[u]) {
                    const uint64_t m = s_colmask[col[u]];
                    if (m && cnt[u]) atomicOr(&s_mask[sl], (unsigned long long)m);   // (a zero count is no presence)
                    slot[u] = sl;
                    done = true;
                    break;
                }
                sl = (sl + 1) & TM;
            }
            if (!done) { s_fail = 1; col[u] = -1; }         // (cannot happen: E <= 3072 < TS)
        }
        __syncthreads();
        if (!a.union_only) {
            // ---- pass 2a: slots whose sets reach min_include become candidate rows ----
#pragma unroll
            for (int u = 0; u < PM2_B; u++) {
                if (col[u] < 0) continue;
                if ((uint32_t)__popcll(s_mask[slot[u]]) < need) { col[u] = -1; continue; }
                if (atomicCAS(&s_rowid[slot[u]], NONE, PENDING) == NONE) {
                    const uint32_t rid = atomicAdd(&s_ncand, 1u);
                    if (rid < (uint32_t)PM2_MAXC) s_ckey[rid] = key[u];
                    s_rowid[slot[u]] = rid;                  // read after the barrier below
                }
            }
            __syncthreads();
            // ---- pass 2b: the entries of the candidates fill their rows ----
            const uint32_t nc_all = s_ncand;
#pragma unroll
            for (int u = 0; u < PM2_B; u++) {
                if (col[u] < 0) continue;
                const uint32_t rid = s_rowid[slot[u]];
                if (rid < (uint32_t)PM2_MAXC) s_crow[(size_t)rid * n + col[u]] = cnt[u];
            }
            if (tid == 0) {
                if (nc_all > (uint32_t)PM2_MAXC) s_fail = 1;
                const uint32_t ncw = min(nc_all, (uint32_t)PM2_MAXC);
                s_wbase = ncw ? atomicAdd((unsigned long long*)&a.counters[4], (unsigned long long)ncw) : 0ull;
                n_keep += nc_all;
            }
            __syncthreads();
            // ---- emit ----
            const uint32_t nc = min(nc_all, (uint32_t)PM2_MAXC);
            const uint64_t wb = s_wbase;
            for (uint32_t i = tid; i < nc * (uint32_t)n; i += PM2_THREADS) {
                const uint32_t r = i / (uint32_t)n;
                if (wb + r < a.cap) a.out_counts[(wb + r) * n + (i - r * n)] = s_crow[i];
                s_crow[i] = 0;
            }
            for (uint32_t r = tid; r < nc; r += PM2_THREADS)
                if (wb + r < a.cap) a.out_keys[wb + r] = s_ckey[r];
        }
        // ---- clear the table ----
        for (uint32_t i = tid; i < TS; i += PM2_THREADS) {
            s_key[i] = SPK_EMPTY_KEY;
            s_mask[i] = 0ull;
            s_rowid[i] = NONE;
        }
        if (tid == 0) s_ncand = 0;
        __syncthreads();
    }
    n_union = spk_warp_sum_u64(n_union);
    n_keep = spk_warp_sum_u64(n_keep);
    if (lane == 0) {
        if (n_union) atomicAdd((unsigned long long*)&a.counters[0], (unsigned long long)n_union);
        if (n_keep) atomicAdd((unsigned long long*)&a.counters[2], (unsigned long long)n_keep);
    }
    __syncthreads();
    if (tid == 0 && s_fail) atomicAdd((unsigned long long*)&a.counters[5], 1ull);
    (void)lengths;
}

// counters of a failed k_pmatrix_filter2 pass are reset before the general kernel redoes the call
__global__ void k_pm_reset_if(uint64_t* counters) {
    if (counters[5]) counters[0] = counters[1] = counters[2] = counters[3] = counters[4] = 0;
}

// Regroup a partition-indexed dump: partition p's run [pindex[2p], +pindex[2p+1]) moves to new_start[p].
// One warp per partition (runs are a few dozen entries); used by the multi-GPU exchange to make every
// destination rank's partition class contiguous before the all-to-all.
__global__ void __launch_bounds__(256)
k_dump_regroup(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ counts,
               const uint32_t* __restrict__ pindex, const uint32_t* __restrict__ new_start, uint64_t P,
               uint64_t* __restrict__ okeys, uint32_t* __restrict__ ocounts) {
    const uint64_t nw = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    const int lane = threadIdx.x & 31;
    for (uint64_t p = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; p < P; p += nw) {
        const uint32_t s = pindex[2 * p], c = pindex[2 * p + 1], d = new_start[p];
        for (uint32_t i = lane; i < c; i += 32) {
            okeys[d + i] = keys[s + i];
            ocounts[d + i] = counts[s + i];
        }
    }
}


// ---- multi-GPU exchange over peer memory (NVLink / NVSwitch P2P stores) ----------------------------------------
// Rank r merges the union rows of the hash partitions p = r (mod world), so it needs that class of every
// chromosome's dump.  Instead of regrouping the dump and handing it to a collective (which needs the sizes on the
// host: a synchronisation per chromosome), the owner writes every partition run straight into the destination
// rank's receive buffer (symmetric allocation, peer-mapped): k_class_offsets gives partition p its offset inside
// its class (exclusive scan over the partitions of the class, ascending), k_dump_scatter_peers copies the run with
// one warp and stores the run length into the destination's per-partition size table.  No host round trip, no
// collective; the transfers ride the same stream as the counting kernels and overlap the next chromosome's work.
__global__ void __launch_bounds__(1024)
k_class_offsets(const uint32_t* __restrict__ pindex, uint64_t P, uint32_t world, uint32_t* __restrict__ dst_off,
                uint64_t* __restrict__ class_tot) {
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_carry;
    const uint32_t cls = blockIdx.x;
    const uint64_t npc = (P - cls + world - 1) / world;             // partitions of this class
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (uint64_t base = 0; base < npc; base += 1024) {
        const uint64_t q = base + threadIdx.x;
        const uint64_t p = cls + q * world;
        const uint32_t c = q < npc ? pindex[2 * p + 1] : 0u;
        uint32_t incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if ((threadIdx.x & 31) >= o) incl += t;
        }
        if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = incl;
        __syncthreads();
        uint32_t wt = s_warp[threadIdx.x & 31], wi = wt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, wi, o);
            if ((threadIdx.x & 31) >= o) wi += t;
        }
        const uint32_t tot = __shfl_sync(0xffffffffu, wi, 31);
        const uint32_t prefix = __shfl_sync(0xffffffffu, wi - wt, threadIdx.x >> 5);
        if (q < npc) dst_off[p] = s_carry + prefix + incl - c;
        __syncthreads();
        if (threadIdx.x == 0) s_carry += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) class_tot[cls] = s_carry;
}

__global__ void __launch_bounds__(256)
k_dump_scatter_peers(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ counts,
                     const uint32_t* __restrict__ pindex, const uint32_t* __restrict__ dst_off, uint64_t P,
                     uint32_t world, uint64_t region_off, uint64_t region_cap, uint64_t psize_off,
                     uint64_t* const* __restrict__ peer_keys, uint32_t* const* __restrict__ peer_counts,
                     uint32_t* const* __restrict__ peer_psize, uint64_t* __restrict__ overflow) {
    const uint64_t nw = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    const int lane = threadIdx.x & 31;
    for (uint64_t p = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; p < P; p += nw) {
        const uint32_t s = pindex[2 * p], c = pindex[2 * p + 1], d = dst_off[p];
        const uint32_t dst = (uint32_t)(p % world);
        const bool fits = (uint64_t)d + c <= region_cap;
        if (lane == 0) {
            peer_psize[dst][psize_off + p / world] = fits ? c : 0u;
            if (!fits) atomicAdd((unsigned long long*)overflow, 1ull);
        }
        if (!fits) continue;
        uint64_t* ok = peer_keys[dst] + region_off + d;
        uint32_t* oc = peer_counts[dst] + region_off + d;
        for (uint32_t i = lane; i < c; i += 32) {
            ok[i] = keys[s + i];
            oc[i] = counts[s + i];
        }
    }
}

uint32_t pm_table_slots(int n, double mean_entries) {
    // smallest power of two that takes an average partition in ONE round (entries <= 7/8 of the slots), capped
    // by shared memory: a slot is an 8-byte key + n 4-byte counters and the table must stay under ~190 KB
    // (one 512-thread CTA per SM)
    uint32_t ts = 256;
    while (ts < 8192 && (double)ts * 0.875 < mean_entries * 1.05) ts <<= 1;
    while (ts > 256 && (size_t)ts * (8 + 4 * (size_t)n) > 190 * 1024) ts >>= 1;
    return ts;
}

}  // namespace

extern "C" int spk_pmatrix_filter(const uint64_t* const* d_keys, const uint32_t* const* d_counts,
                                  const uint32_t* const* d_pindex, int n, int pbits, uint32_t nparts, uint32_t part,
                                  const uint64_t* d_lengths, const int32_t* d_set_off, int n_sets,
                                  const int32_t* d_grp_off, int n_groups, const int32_t* d_members, int n_members,
                                  double min_fold, int baseline, int by_count, double ratio, double min_freq,
                                  double max_freq,
                                  uint64_t* d_out_keys, uint32_t* d_out_counts, uint64_t cap,
                                  uint64_t total_entries, uint64_t* d_counters, void* stream) {
    SPK_CHECK_ARG(d_keys && d_counts && d_pindex && d_counters, "null pointer");
    SPK_CHECK_ARG(n_sets == 0 || (d_lengths && d_set_off && d_grp_off && d_members), "null configuration");
    SPK_CHECK_ARG(n >= 1 && n <= PM_MAX_COLS, "1 <= n <= 512 chromosomes");
    SPK_CHECK_ARG(pbits >= 0 && pbits <= 30, "bad pbits");
    SPK_CHECK_ARG(nparts >= 1 && part < nparts, "bad partition");
    SPK_CHECK_ARG((n_sets >= 1 && n_groups >= 1 && n_members >= n_groups) || (n_sets == 0 && cap == 0),
                  "bad shape (n_sets == 0 with cap == 0 counts the union only)");
    SPK_CHECK_ARG(baseline < MX_MAX_GROUPS_PER_SET && baseline >= -MX_MAX_GROUPS_PER_SET, "baseline out of range");
    SPK_CHECK_ARG(cap == 0 || (d_out_keys && d_out_counts), "null output");
    cudaStream_t st = (cudaStream_t)stream;
    SPK_CUDA(cudaMemsetAsync(d_counters, 0, 8 * sizeof(uint64_t), st));
    PmArgs a;
    a.keys = d_keys;
    a.counts = d_counts;
    a.pindex = d_pindex;
    a.n = n;
    a.P = 1ull << pbits;
    a.nparts = nparts;
    a.part = part;
    a.lengths = d_lengths;
    a.cfg = FilterCfg{d_set_off, d_grp_off, d_members, n_sets, min_fold, baseline, by_count, ratio, min_freq, max_freq,
                      nullptr, 0, 0};
    a.n_groups = n_groups;
    a.union_only = (cap == 0 && n_sets == 0) ? 1 : 0;
    a.out_keys = d_out_keys;
    a.out_counts = d_out_counts;
    a.cap = cap;
    a.counters = d_counters;
    a.tslots = pm_table_slots(n, (double)total_entries / (double)(1ull << pbits));
    size_t smem = (size_t)a.tslots * (8 + 4 * (size_t)n) + (size_t)(2 * n + 2) * 4 + (size_t)a.tslots * 2;
    smem = (smem + 7) / 8 * 8;
    if (!a.union_only) smem += spk_filter_stage_bytes(n_sets, n_groups, n_members, n);
    SPK_CHECK_ARG(smem <= 220 * 1024, "homoeolog configuration too large");
    SPK_CUDA(cudaFuncSetAttribute(k_pmatrix_filter, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const uint64_t mine = (a.P + nparts - 1 - part) / nparts;
    if (mine == 0) return SPK_OK;
    const unsigned grid = (unsigned)min((uint64_t)spk_num_sms(), mine);
    // presence-mask kernel first (<= 64 sets, a fold threshold that all-zero sets fail); the general kernel runs only
    // if that pass reports something it could not take (device-side decision, no host synchronisation)
    const uint64_t* run_if = nullptr;
    const char* pm = getenv("SPK_PMATRIX_KERNEL");                 // "general": skip the presence-mask kernel (tests)
    if (n_sets <= 64 && (a.union_only || !(0.0 >= min_fold)) && !(pm && pm[0] == 'g')) {
        size_t smem2 = (size_t)PM2_TS * 16 + (size_t)n * 8 + (size_t)PM2_MAXC * 8 + (size_t)n * 16 + (size_t)PM2_TS * 4 +
                       (size_t)PM2_MAXC * n * 4 + (size_t)(2 * n + 2) * 4 + 16;
        if (!a.union_only) smem2 += spk_filter_stage_bytes(n_sets, n_groups, n_members, n);
        if (smem2 <= 110 * 1024) {
            SPK_CUDA(cudaFuncSetAttribute(k_pmatrix_filter2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
            k_pmatrix_filter2<<<(unsigned)min((uint64_t)spk_num_sms() * 2, mine), PM2_THREADS, smem2, st>>>(a);
            SPK_LAUNCH_CHECK();
            k_pm_reset_if<<<1, 1, 0, st>>>(d_counters);
            SPK_LAUNCH_CHECK();
            run_if = d_counters + 5;
        }
    }
    k_pmatrix_filter<<<grid, PM_THREADS, smem, st>>>(a, run_if);
    SPK_LAUNCH_CHECK();
    return SPK_OK;
}

extern "C" int spk_dump_regroup(const uint64_t* d_keys, const uint32_t* d_counts, const uint32_t* d_pindex,
                                const uint32_t* d_new_start, int pbits, uint64_t* d_out_keys, uint32_t* d_out_counts,
                                void* stream) {
    SPK_CHECK_ARG(d_pindex && d_new_start, "null pointer");
    SPK_CHECK_ARG(pbits >= 0 && pbits <= 30, "bad pbits");
    SPK_CHECK_ARG((d_keys && d_counts && d_out_keys && d_out_counts) || true, "null pointer");
    const uint64_t P = 1ull << pbits;
    const unsigned grid = (unsigned)min((P * 32 + 255) / 256, (uint64_t)spk_num_sms() * 16);
    k_dump_regroup<<<grid, 256, 0, (cudaStream_t)stream>>>(d_keys, d_counts, d_pindex, d_new_start, P, d_out_keys,
                                                           d_out_counts);
    SPK_LAUNCH_CHECK();
    return SPK_OK;
}

extern "C" int spk_dump_scatter_peers(const uint64_t* d_keys, const uint32_t* d_counts, const uint32_t* d_pindex,
                                      int pbits, uint32_t world, uint64_t region_off, uint64_t region_cap,
                                      uint64_t psize_off, uint64_t* const* d_peer_keys,
                                      uint32_t* const* d_peer_counts, uint32_t* const* d_peer_psize,
                                      uint32_t* d_dst_off, uint64_t* d_class_tot, uint64_t* d_overflow, void* stream) {
    SPK_CHECK_ARG(d_pindex && d_peer_keys && d_peer_counts && d_peer_psize && d_dst_off && d_class_tot && d_overflow,
                  "null pointer");
    SPK_CHECK_ARG(pbits >= 0 && pbits <= 30, "bad pbits");
    SPK_CHECK_ARG(world >= 1 && world <= 1024, "bad world size");
    const uint64_t P = 1ull << pbits;
    cudaStream_t st = (cudaStream_t)stream;
    k_class_offsets<<<world, 1024, 0, st>>>(d_pindex, P, world, d_dst_off, d_class_tot);
    SPK_LAUNCH_CHECK();
    const unsigned grid = (unsigned)min((P * 32 + 255) / 256, (uint64_t)spk_num_sms() * 16);
    k_dump_scatter_peers<<<grid, 256, 0, st>>>(d_keys, d_counts, d_pindex, d_dst_off, P, world, region_off, region_cap,
                                               psize_off, d_peer_keys, d_peer_counts, d_peer_psize, d_overflow);
    SPK_LAUNCH_CHECK();
    return SPK_OK;
}
