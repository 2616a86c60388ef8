// K3b/K4: union of the per-chromosome (k-mer, count) lists -> count matrix -> differential filter.
//
// Replaces JellyfishDumps.to_matrix (Jellyfish.py:439-460: dict kmer -> [count per chromosome], 0
// where the k-mer was not dumped) and JellyfishDumps.filter / _filter_kmer (Jellyfish.py:462-512,
// 611-648).  The fp64 arithmetic of _filter_kmer is reproduced operation by operation (this file is
// compiled with -fmad=false): int sums, one IEEE division per group, `max/(min+1e-20) >= min_fold`.
#include "spk_common.cuh"
#include "spk_filter.cuh"

namespace {

constexpr int MX_THREADS = 256;

__global__ void __launch_bounds__(MX_THREADS)
k_union_insert(const uint64_t* __restrict__ keys, uint64_t n, uint64_t* __restrict__ ukeys,
               uint32_t* __restrict__ urows, uint64_t uslots, uint32_t* __restrict__ nrows,
               uint64_t* __restrict__ fail, uint32_t nparts, uint32_t part) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t key = keys[i];
        const uint64_t h = spk_hash64(key);
        if (nparts > 1 && (uint32_t)(h & 0xffffffffu) % nparts != part) continue;   // another rank's row
        uint64_t slot = spk_slot_of(h, uslots);
        bool done = false;
        for (uint64_t p = 0; p < uslots; p++) {
            uint64_t cur = __ldcg(ukeys + slot);
            if (cur == SPK_EMPTY_KEY) {
                const uint64_t old = atomicCAS((unsigned long long*)(ukeys + slot),
                                               (unsigned long long)SPK_EMPTY_KEY,
                                               (unsigned long long)key);
                if (old == SPK_EMPTY_KEY) {
                    // winner assigns the row id; readers of urows run in later kernels
                    urows[slot] = atomicAdd(nrows, 1u);
                    done = true;
                    break;
                }
                cur = old;
            }
            if (cur == key) {
                done = true;
                break;
            }
            slot++;
            if (slot == uslots) slot = 0;
        }
        if (!done) atomicAdd((unsigned long long*)fail, 1ull);
    }
}

__global__ void __launch_bounds__(MX_THREADS)
k_matrix_fill(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ counts, uint64_t n,
              const uint64_t* __restrict__ ukeys, const uint32_t* __restrict__ urows, uint64_t uslots,
              uint32_t* __restrict__ matrix, uint64_t* __restrict__ row_keys, int ncol, int col,
              uint32_t nparts, uint32_t part) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t key = keys[i];
        const uint64_t h = spk_hash64(key);
        if (nparts > 1 && (uint32_t)(h & 0xffffffffu) % nparts != part) continue;
        uint64_t slot = spk_slot_of(h, uslots);
        for (uint64_t p = 0; p < uslots; p++) {
            const uint64_t cur = __ldg(ukeys + slot);
            if (cur == key) {
                const uint32_t row = urows[slot];
                matrix[(uint64_t)row * ncol + col] = counts[i];
                row_keys[row] = key;
                break;
            }
            if (cur == SPK_EMPTY_KEY) break;  // not in the union (cannot happen after union_insert)
            slot++;
            if (slot == uslots) slot = 0;
        }
    }
}

// One thread per matrix row.  Mirrors _filter_kmer (Jellyfish.py:611-648) with outfig set.
__global__ void __launch_bounds__(MX_THREADS)
k_filter(const uint32_t* __restrict__ matrix, uint64_t nrows, int ncol,
         const uint64_t* __restrict__ lengths_g, FilterCfg cfg_g, int n_groups, uint8_t* __restrict__ flags,
         uint64_t* __restrict__ tot_out, uint64_t* __restrict__ counters) {
    extern __shared__ __align__(16) uint8_t s_cfg[];
    const uint64_t* lengths;
    const FilterCfg cfg = spk_filter_stage(cfg_g, n_groups, ncol, lengths_g, s_cfg, &lengths);
    uint64_t n_fold = 0, n_keep = 0;
    for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < nrows;
         r += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t tot;
        const uint8_t fl = spk_filter_row(matrix + r * ncol, ncol, lengths, cfg, tot);
        flags[r] = fl;
        tot_out[r] = tot;
        n_fold += fl & 1;
        n_keep += (fl >> 1) & 1;
    }
    n_fold = spk_warp_sum_u64(n_fold);
    n_keep = spk_warp_sum_u64(n_keep);
    if ((threadIdx.x & 31) == 0) {
        if (n_fold) atomicAdd((unsigned long long*)&counters[0], (unsigned long long)n_fold);
        if (n_keep) atomicAdd((unsigned long long*)&counters[1], (unsigned long long)n_keep);
    }
}

// Exclusive scan of the keep flags -> output positions, three passes: per-CTA totals over 8192-row
// segments, a single-CTA scan of those totals, then the in-segment scan.
constexpr int SF_PER = 8;
constexpr int SF_SEG = 1024 * SF_PER;

__device__ __forceinline__ uint32_t sf_block_scan(uint32_t sum, uint32_t* s_warp, uint32_t* total) {
    uint32_t incl = sum;
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
    return prefix + incl - sum;
}

__global__ void __launch_bounds__(1024) k_flags_seg_totals(const uint8_t* __restrict__ flags, uint64_t n,
                                                            uint32_t* __restrict__ seg_tot) {
    __shared__ uint32_t s_warp[32];
    const uint64_t i0 = (uint64_t)blockIdx.x * SF_SEG + (uint64_t)threadIdx.x * SF_PER;
    uint32_t sum = 0;
#pragma unroll
    for (int q = 0; q < SF_PER; q++) sum += (i0 + q < n) ? ((flags[i0 + q] >> 1) & 1u) : 0u;
    uint32_t tot;
    sf_block_scan(sum, s_warp, &tot);
    if (threadIdx.x == 0) seg_tot[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(1024) k_flags_scan_totals(uint32_t* seg_tot, uint64_t nseg,
                                                             uint32_t* __restrict__ grand) {
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (uint64_t base = 0; base < nseg; base += 1024) {
        const uint64_t i = base + threadIdx.x;
        const uint32_t v = (i < nseg) ? seg_tot[i] : 0;
        uint32_t tot;
        const uint32_t excl = sf_block_scan(v, s_warp, &tot);
        if (i < nseg) seg_tot[i] = s_carry + excl;
        __syncthreads();
        if (threadIdx.x == 0) s_carry += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) *grand = s_carry;
}

__global__ void __launch_bounds__(1024) k_flags_seg_scan(const uint8_t* __restrict__ flags, uint64_t n,
                                                          const uint32_t* __restrict__ seg_off,
                                                          uint32_t* __restrict__ pos) {
    __shared__ uint32_t s_warp[32];
    const uint64_t i0 = (uint64_t)blockIdx.x * SF_SEG + (uint64_t)threadIdx.x * SF_PER;
    uint32_t loc[SF_PER], sum = 0;
#pragma unroll
    for (int q = 0; q < SF_PER; q++) {
        loc[q] = (i0 + q < n) ? ((flags[i0 + q] >> 1) & 1u) : 0u;
        sum += loc[q];
    }
    uint32_t tot;
    uint32_t run = seg_off[blockIdx.x] + sf_block_scan(sum, s_warp, &tot);
#pragma unroll
    for (int q = 0; q < SF_PER; q++) {
        if (i0 + q < n) pos[i0 + q] = run;
        run += loc[q];
    }
}

// compaction of the kept rows in row order: (key, row id) pairs
__global__ void __launch_bounds__(MX_THREADS)
k_filter_select(const uint64_t* __restrict__ row_keys, const uint8_t* __restrict__ flags, uint64_t nrows,
                const uint32_t* __restrict__ pos, uint64_t* __restrict__ out_keys,
                uint32_t* __restrict__ out_rows, uint64_t cap) {
    for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < nrows;
         r += (uint64_t)gridDim.x * blockDim.x) {
        if (!(flags[r] & 2)) continue;
        const uint64_t o = pos[r];
        if (o >= cap) continue;
        out_keys[o] = row_keys[r];
        out_rows[o] = (uint32_t)r;
    }
}

// normalised rows in the order given by `rows`: count / length (Jellyfish.py:648)
__global__ void __launch_bounds__(MX_THREADS)
k_filter_emit(const uint32_t* __restrict__ matrix, const uint64_t* __restrict__ tot,
              const uint32_t* __restrict__ rows, uint64_t m, int ncol,
              const uint64_t* __restrict__ lengths, double* __restrict__ out_norm,
              uint64_t* __restrict__ out_tot) {
    for (uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < m * (uint64_t)ncol;
         e += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t j = e / ncol;
        const int c = (int)(e - j * ncol);
        const uint64_t r = rows[j];
        out_norm[e] = (double)matrix[r * ncol + c] / (double)lengths[c];
        if (c == 0) out_tot[j] = tot[r];
    }
}

unsigned grid_for(uint64_t n) {
    const uint64_t blocks = (n + MX_THREADS - 1) / MX_THREADS;
    const uint64_t cap = (uint64_t)spk_num_sms() * 16;
    return (unsigned)(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
}

}  // namespace

extern "C" int spk_union_insert(const uint64_t* d_keys, uint64_t n, uint64_t* d_ukeys,
                                uint32_t* d_urows, uint64_t uslots, uint32_t* d_nrows,
                                uint64_t* d_fail, uint32_t nparts, uint32_t part, void* stream) {
    SPK_CHECK_ARG(nparts >= 1 && part < nparts, "bad partition");
    SPK_CHECK_ARG(d_ukeys && d_urows && d_nrows && d_fail, "null pointer");
    SPK_CHECK_ARG(uslots >= 2, "uslots too small");
    if (n == 0) return SPK_OK;
    SPK_CHECK_ARG(d_keys, "null keys");
    k_union_insert<<<grid_for(n), MX_THREADS, 0, (cudaStream_t)stream>>>(d_keys, n, d_ukeys, d_urows,
                                                                         uslots, d_nrows, d_fail, nparts, part);
    SPK_LAUNCH_CHECK();
    return SPK_OK;
}

extern "C" int spk_matrix_fill(const uint64_t* d_keys, const uint32_t* d_counts, uint64_t n,
                               const uint64_t* d_ukeys, const uint32_t* d_urows, uint64_t uslots,
                               uint32_t* d_matrix, uint64_t* d_row_keys, int ncol, int col,
                               uint32_t nparts, uint32_t part, void* stream) {
    SPK_CHECK_ARG(nparts >= 1 && part < nparts, "bad partition");
    SPK_CHECK_ARG(d_ukeys && d_urows && d_matrix && d_row_keys, "null pointer");
    SPK_CHECK_ARG(ncol >= 1 && col >= 0 && col < ncol, "bad column");
    if (n == 0) return SPK_OK;
    SPK_CHECK_ARG(d_keys && d_counts, "null keys/counts");
    k_matrix_fill<<<grid_for(n), MX_THREADS, 0, (cudaStream_t)stream>>>(
        d_keys, d_counts, n, d_ukeys, d_urows, uslots, d_matrix, d_row_keys, ncol, col, nparts, part);
    SPK_LAUNCH_CHECK();
    return SPK_OK;
}

extern "C" int spk_filter_differential(const uint32_t* d_matrix, uint64_t nrows, int ncol,
                                       const uint64_t* d_lengths, const int32_t* d_set_off, int n_sets,
                                       const int32_t* d_grp_off, int n_groups,
                                       const int32_t* d_members, int n_members, double min_fold, int baseline,
                                       int by_count, double ratio, double min_freq, double max_freq,
                                       uint8_t* d_flags, uint64_t* d_tot, uint64_t* d_counters,
                                       void* stream) {
    SPK_CHECK_ARG(d_lengths && d_set_off && d_grp_off && d_members && d_flags && d_tot && d_counters,
                  "null pointer");
    SPK_CHECK_ARG(ncol >= 1 && n_sets >= 1 && n_groups >= 1, "bad shape");
    SPK_CHECK_ARG(baseline < MX_MAX_GROUPS_PER_SET && baseline >= -MX_MAX_GROUPS_PER_SET,
                  "baseline out of range");
    if (nrows == 0) return SPK_OK;
    SPK_CHECK_ARG(d_matrix, "null matrix");
    FilterCfg cfg{d_set_off, d_grp_off, d_members, n_sets, min_fold, baseline, by_count, ratio,
                  min_freq, max_freq, nullptr, 0, 0};
    SPK_CHECK_ARG(n_members >= n_groups, "n_members must be the length of d_members");
    const size_t smem = spk_filter_stage_bytes(n_sets, n_groups, n_members, ncol);   // configuration staged in smem
    SPK_CHECK_ARG(smem <= 48 * 1024, "homoeolog configuration too large");
    k_filter<<<grid_for(nrows), MX_THREADS, smem, (cudaStream_t)stream>>>(d_matrix, nrows, ncol, d_lengths,
                                                                          cfg, n_groups, d_flags, d_tot, d_counters);
    SPK_LAUNCH_CHECK();
    return SPK_OK;
}

extern "C" int spk_filter_select(const uint64_t* d_row_keys, const uint8_t* d_flags, uint64_t nrows,
                                 uint32_t* d_scan_ws, uint64_t* d_out_keys, uint32_t* d_out_rows,
                                 uint64_t cap, void* stream) {
    SPK_CHECK_ARG(d_flags && d_scan_ws, "null pointer");
    SPK_CHECK_ARG(nrows < 0xffffffffull, "too many rows");
    cudaStream_t st = (cudaStream_t)stream;
    // d_scan_ws: [0, nrows] positions (+ grand total at [nrows]), then the per-segment totals
    const uint64_t nseg = (nrows + SF_SEG - 1) / SF_SEG;
    uint32_t* seg = d_scan_ws + nrows + 1;
    if (nseg) {
        k_flags_seg_totals<<<(unsigned)nseg, 1024, 0, st>>>(d_flags, nrows, seg);
        SPK_LAUNCH_CHECK();
    }
    k_flags_scan_totals<<<1, 1024, 0, st>>>(seg, nseg, d_scan_ws + nrows);
    SPK_LAUNCH_CHECK();
    if (nseg) {
        k_flags_seg_scan<<<(unsigned)nseg, 1024, 0, st>>>(d_flags, nrows, seg, d_scan_ws);
        SPK_LAUNCH_CHECK();
    }
    if (nrows == 0 || cap == 0) return SPK_OK;
    SPK_CHECK_ARG(d_row_keys && d_out_keys && d_out_rows, "null pointer");
    k_filter_select<<<grid_for(nrows), MX_THREADS, 0, st>>>(d_row_keys, d_flags, nrows, d_scan_ws,
                                                            d_out_keys, d_out_rows, cap);
    SPK_LAUNCH_CHECK();
    return SPK_OK;
}

extern "C" int spk_filter_emit(const uint32_t* d_matrix, const uint64_t* d_tot, const uint32_t* d_rows,
                               uint64_t m, int ncol, const uint64_t* d_lengths, double* d_out_norm,
                               uint64_t* d_out_tot, void* stream) {
    SPK_CHECK_ARG(ncol >= 1, "bad shape");
    if (m == 0) return SPK_OK;
    SPK_CHECK_ARG(d_matrix && d_tot && d_rows && d_lengths && d_out_norm && d_out_tot, "null pointer");
    k_filter_emit<<<grid_for(m * (uint64_t)ncol), MX_THREADS, 0, (cudaStream_t)stream>>>(
        d_matrix, d_tot, d_rows, m, ncol, d_lengths, d_out_norm, d_out_tot);
    SPK_LAUNCH_CHECK();
    return SPK_OK;
}

// ---- histogram of the totals of all fold-passing k-mers (the `.kmer_freq.pdf` data, Jellyfish.py:499-511,650-666) --------
// The reference collects `tot` of every k-mer that passes the fold test into a Python list and hands it to
// matplotlib (bins of 25 occurrences, x axis cut at the 99th percentile).  Here the totals never leave the device:
// one pass for count / min / max, one pass that bins them with numpy's `histogram` edge rules, and a radix select
// (16 bits per pass) for the order statistics np.percentile interpolates between.
namespace {

__global__ void __launch_bounds__(256)
k_tot_minmax(const uint64_t* __restrict__ tot, const uint8_t* __restrict__ flags, uint64_t n,
             unsigned long long* __restrict__ out /* [0] count, [1] min, [2] max */) {
    unsigned long long cnt = 0, mn = ~0ull, mx = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        if (flags[i] & 1) {
            const unsigned long long v = tot[i];
            cnt++;
            mn = min(mn, v);
            mx = max(mx, v);
        }
    cnt = spk_warp_sum_u64(cnt);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mn = min(mn, (unsigned long long)__shfl_xor_sync(0xffffffffu, mn, o));
        mx = max(mx, (unsigned long long)__shfl_xor_sync(0xffffffffu, mx, o));
    }
    if ((threadIdx.x & 31) == 0 && cnt) {
        atomicAdd(&out[0], cnt);
        atomicMin(&out[1], mn);
        atomicMax(&out[2], mx);
    }
}

// numpy.histogram(data, bins=nbins) with range (mn, mx): uniform edges e_i = mn + i * (mx - mn) / nbins (last = mx),
// index = int((x - mn) / (mx - mn) * nbins) corrected against the edges, the last bin closed on the right
__global__ void __launch_bounds__(256)
k_tot_hist(const uint64_t* __restrict__ tot, const uint8_t* __restrict__ flags, uint64_t n, double mn, double mx,
           uint32_t nbins, unsigned long long* __restrict__ hist) {
    const double step = (mx - mn) / (double)nbins;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        if (!(flags[i] & 1)) continue;
        const double x = (double)tot[i];
        uint32_t b = 0;
        if (mx > mn) {
            long long f = (long long)(((x - mn) / (mx - mn)) * (double)nbins);     // numpy's order of operations
            if (f >= (long long)nbins) f = nbins - 1;
            b = (uint32_t)f;
            const double lo = (b == nbins) ? mx : mn + (double)b * step;
            const double hi = (b + 1 == nbins) ? mx : mn + (double)(b + 1) * step;
            if (x < lo && b > 0) b--;
            else if (x >= hi && b + 1 != nbins) b++;
        }
        atomicAdd(&hist[b], 1ull);
    }
}

// radix select pass: 65536-bin histogram of bits [shift, shift + 16) of the values whose higher bits equal `prefix`
__global__ void __launch_bounds__(256)
k_tot_select(const uint64_t* __restrict__ tot, const uint8_t* __restrict__ flags, uint64_t n, int shift, uint64_t prefix,
             unsigned long long* __restrict__ hist) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        if (!(flags[i] & 1)) continue;
        const uint64_t v = tot[i];
        if (shift + 16 < 64 && (v >> (shift + 16)) != prefix) continue;
        atomicAdd(&hist[(v >> shift) & 0xffffu], 1ull);
    }
}

}  // namespace

extern "C" int spk_tot_minmax(const uint64_t* d_tot, const uint8_t* d_flags, uint64_t n, uint64_t* d_out, void* stream) {
    SPK_CHECK_ARG(d_out, "null output");
    cudaStream_t st = (cudaStream_t)stream;
    const uint64_t init[3] = {0ull, ~0ull, 0ull};
    SPK_CUDA(cudaMemcpyAsync(d_out, init, sizeof(init), cudaMemcpyHostToDevice, st));
    if (n == 0) return SPK_OK;
    SPK_CHECK_ARG(d_tot && d_flags, "null pointer");
    k_tot_minmax<<<(unsigned)min((n + 255) / 256, (uint64_t)spk_num_sms() * 16), 256, 0, st>>>(d_tot, d_flags, n,
                                                                                                (unsigned long long*)d_out);
    SPK_LAUNCH_CHECK();
    return SPK_OK;
}

extern "C" int spk_tot_histogram(const uint64_t* d_tot, const uint8_t* d_flags, uint64_t n, double mn, double mx,
                                 uint32_t nbins, uint64_t* d_hist, void* stream) {
    SPK_CHECK_ARG(d_hist && nbins >= 1, "bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    SPK_CUDA(cudaMemsetAsync(d_hist, 0, (size_t)nbins * 8, st));
    if (n == 0) return SPK_OK;
    SPK_CHECK_ARG(d_tot && d_flags, "null pointer");
    k_tot_hist<<<(unsigned)min((n + 255) / 256, (uint64_t)spk_num_sms() * 16), 256, 0, st>>>(d_tot, d_flags, n, mn, mx, nbins,
                                                                                              (unsigned long long*)d_hist);
    SPK_LAUNCH_CHECK();
    return SPK_OK;
}

extern "C" int spk_tot_select_pass(const uint64_t* d_tot, const uint8_t* d_flags, uint64_t n, int shift, uint64_t prefix,
                                   uint64_t* d_hist65536, void* stream) {
    SPK_CHECK_ARG(d_hist65536 && shift >= 0 && shift <= 48 && shift % 16 == 0, "bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    SPK_CUDA(cudaMemsetAsync(d_hist65536, 0, 65536 * 8, st));
    if (n == 0) return SPK_OK;
    SPK_CHECK_ARG(d_tot && d_flags, "null pointer");
    k_tot_select<<<(unsigned)min((n + 255) / 256, (uint64_t)spk_num_sms() * 16), 256, 0, st>>>(d_tot, d_flags, n, shift, prefix,
                                                                                                (unsigned long long*)d_hist65536);
    SPK_LAUNCH_CHECK();
    return SPK_OK;
}
