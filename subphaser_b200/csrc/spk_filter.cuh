// The differential-k-mer test of one matrix row, shared by the plain filter (spk_matrix.cu) and the
// partitioned union+filter (spk_pmatrix.cu).  Mirrors _filter_kmer (Jellyfish.py:611-648) with outfig set,
// operation by operation in fp64 (compile the including file with -fmad=false).
#pragma once
#include <stdint.h>

constexpr int MX_MAX_GROUPS_PER_SET = 32;

struct FilterCfg {
    const int32_t* set_off;
    const int32_t* grp_off;
    const int32_t* members;
    int n_sets;
    double min_fold;
    int baseline;
    int by_count;
    double ratio;
    double min_freq;
    double max_freq;
    const double* grp_rlen;   // optional [n_groups]: 1 / (sum of the group's chromosome lengths) — enables the
                              // division-free decision below (nullptr: always the operation-by-operation test)
    int n_multi;              // with grp_rlen: number of sets with >= 2 groups (`all` of Jellyfish.py:621-645)
    int min_include;          // with grp_rlen: smallest m with !(1.0*m/all < ratio)  (all+1: no m passes)
};

// shared-memory copy of the configuration (+ the per-group reciprocal lengths): a CTA calls spk_filter_stage
// once; rows are then tested without touching global memory.  Layout in `smem` (int32 words, then doubles):
// set_off[n_sets+1], grp_off[n_groups+1], members[n_members], pad, lengths u64[ncol], rlen f64[n_groups]
__host__ __device__ inline size_t spk_filter_stage_bytes(int n_sets, int n_groups, int n_members, int ncol) {
    size_t w = (size_t)(n_sets + 1) + (size_t)(n_groups + 1) + (size_t)n_members;
    w = (w + 1) / 2 * 2;
    return w * 4 + (size_t)ncol * 8 + (size_t)n_groups * 8;
}

__device__ __forceinline__ FilterCfg spk_filter_stage(const FilterCfg& g, int n_groups, int ncol,
                                                      const uint64_t* __restrict__ lengths, void* smem,
                                                      const uint64_t** lengths_out) {
    const int n_members = g.grp_off[n_groups];
    int32_t* s_set = (int32_t*)smem;
    int32_t* s_grp = s_set + g.n_sets + 1;
    int32_t* s_mem = s_grp + n_groups + 1;
    size_t w = (size_t)(g.n_sets + 1) + (size_t)(n_groups + 1) + (size_t)n_members;
    w = (w + 1) / 2 * 2;
    uint64_t* s_len = (uint64_t*)((int32_t*)smem + w);
    double* s_rl = (double*)(s_len + ncol);
    for (int i = threadIdx.x; i <= g.n_sets; i += blockDim.x) s_set[i] = g.set_off[i];
    for (int i = threadIdx.x; i <= n_groups; i += blockDim.x) s_grp[i] = g.grp_off[i];
    for (int i = threadIdx.x; i < n_members; i += blockDim.x) s_mem[i] = g.members[i];
    for (int i = threadIdx.x; i < ncol; i += blockDim.x) s_len[i] = lengths[i];
    for (int i = threadIdx.x; i < n_groups; i += blockDim.x) {
        uint64_t ls = 0;
        for (int m = g.grp_off[i]; m < g.grp_off[i + 1]; m++) ls += lengths[g.members[m]];
        s_rl[i] = 1.0 / (double)ls;
    }
    __syncthreads();
    FilterCfg c = g;
    int all = 0;
    for (int i = 0; i < g.n_sets; i++) all += (g.set_off[i + 1] - g.set_off[i] >= 2) ? 1 : 0;
    int need = all + 1;
    for (int m = all; m >= 0; m--)                  // m/all is monotonic in m: the passing m form a suffix
        if (!(1.0 * (double)m / (double)all < g.ratio)) need = m;
        else break;
    c.n_multi = all;
    c.min_include = need;
    c.set_off = s_set;
    c.grp_off = s_grp;
    c.members = s_mem;
    c.grp_rlen = s_rl;
    *lengths_out = s_len;
    return c;
}

// Integer pre-screen (needs a staged configuration): false = the row certainly fails the fold test.  A
// non-singleton set whose counts are all zero has every frequency exactly 0.0 and passes only if
// 1.0*0.0/(0.0+1e-20) >= min_fold (false for any positive min_fold); if even passing every other set cannot
// reach `min_include` sets the row is rejected without any floating-point work.  For a real union (k-mers
// dumped by a few chromosomes only) this removes ~98 % of the rows.
template <typename RowT>
__device__ __forceinline__ bool spk_filter_prescreen(const RowT* row, const FilterCfg& cfg) {
    if (1.0 * 0.0 / (0.0 + 1e-20) >= cfg.min_fold) return true;   // degenerate fold: nothing can be rejected
    int zsets = 0;
    for (int s = 0; s < cfg.n_sets; s++) {
        const int g0 = cfg.set_off[s], g1 = cfg.set_off[s + 1];
        if (g1 - g0 < 2) continue;
        uint64_t any = 0;
        for (int m = cfg.grp_off[g0]; m < cfg.grp_off[g1]; m++) any |= row[cfg.members[m]];
        zsets += any == 0;
    }
    return cfg.n_multi - zsets >= cfg.min_include;
}

// -> flags: bit 0 = passed the fold test (include/all >= ratio), bit 1 = also min_freq <= tot <= max_freq.
// Two shortcuts that cannot change the result: a group whose counts are all zero has frequency exactly 0.0
// (0/len is exact; the division is kept when len == 0 so 0/0 still gives NaN), and the set loop stops as
// soon as even passing every remaining set could not reach `ratio` (IEEE division is monotonic in the
// numerator, so (include + remaining)/all < ratio implies include_final/all < ratio).
template <typename RowT>
__device__ __forceinline__ uint8_t spk_filter_row(const RowT* row, int ncol, const uint64_t* __restrict__ lengths,
                                                  const FilterCfg& cfg, uint64_t& tot_out) {
    uint64_t tot = 0;
    for (int c = 0; c < ncol; c++) tot += row[c];
    tot_out = tot;
    // integer pre-screen: a non-singleton set whose counts are all zero has every frequency exactly 0.0, so it
    // passes iff 1.0*0.0/(0.0+1e-20) >= min_fold (false for any positive min_fold).  If even passing all the
    // other sets cannot reach `ratio`, the row fails the fold test — decided without any fp64 division
    // (this rejects the bulk of a real union: k-mers seen in a few chromosomes only).  Chromosome lengths are
    // positive here: the caller refuses `lengths[i] == 0` (Jellyfish.py:487-489) before any row is tested.
    int all = 0, zsets = 0;
    for (int s = 0; s < cfg.n_sets; s++) {
        const int g0 = cfg.set_off[s], g1 = cfg.set_off[s + 1];
        if (g1 - g0 < 2) continue;                             // singleton sets are ignored (Jellyfish.py:622-623)
        all++;
        uint64_t any = 0;
        for (int m = cfg.grp_off[g0]; m < cfg.grp_off[g1]; m++) any |= row[cfg.members[m]];
        zsets += any == 0;
    }
    if (zsets && !(1.0 * 0.0 / (0.0 + 1e-20) >= cfg.min_fold) &&
        1.0 * (double)(all - zsets) / (double)all < cfg.ratio)
        return 0;
    // Division-free decision (when the reciprocal group lengths are staged): f~ = count * (1/len) is within
    // 2 ulp of count/len, order statistics move by no more than their inputs, so
    //   fmax/(fbase+1e-20) >= min_fold   <=>   f~max >= min_fold*(f~base+1e-20)
    // whenever the two sides differ by more than 1e-9 relative — seven orders of magnitude above the rounding
    // of either formulation.  Rows inside that band (or with inf/NaN) take the exact path below.
    if (cfg.grp_rlen) {
        int include = 0, seen = 0;
        bool certain = true, fail = false;
        for (int s = 0; s < cfg.n_sets && certain && !fail; s++) {
            const int g0 = cfg.set_off[s], g1 = cfg.set_off[s + 1];
            const int ng = g1 - g0;
            if (ng < 2) continue;
            seen++;
            const int bidx = cfg.baseline >= 0 ? cfg.baseline : ng + cfg.baseline;   // rank (descending) of the baseline
            double top = -1.0, base = 0.0;
            if (ng <= 4) {
                double f0 = -1.0, f1 = -1.0, f2 = -1.0, f3 = -1.0;   // frequencies are >= 0: -1 pads absent groups
                for (int g = g0; g < g1; g++) {
                    uint64_t cs = 0;
                    for (int m = cfg.grp_off[g]; m < cfg.grp_off[g + 1]; m++) cs += row[cfg.members[m]];
                    const double v = cfg.by_count ? (double)cs : (double)cs * cfg.grp_rlen[g];
                    const int i = g - g0;
                    if (i == 0) f0 = v;
                    else if (i == 1) f1 = v;
                    else if (i == 2) f2 = v;
                    else f3 = v;
                }
                // descending sorting network on four registers
#define SPK_CSWAP(a, b) { const double hi_ = fmax(a, b), lo_ = fmin(a, b); a = hi_; b = lo_; }
                SPK_CSWAP(f0, f1) SPK_CSWAP(f2, f3) SPK_CSWAP(f0, f2) SPK_CSWAP(f1, f3) SPK_CSWAP(f1, f2)
#undef SPK_CSWAP
                top = f0;
                base = bidx == 0 ? f0 : bidx == 1 ? f1 : bidx == 2 ? f2 : f3;
            } else {
                certain = false;                             // wide sets: exact path
            }
            if (!certain) break;
            const double rhs = cfg.min_fold * (base + 1e-20);
            if (top > rhs * (1.0 + 1e-9) && rhs >= 0.0) include++;
            else if (top < rhs * (1.0 - 1e-9) && rhs >= 0.0) { /* fails */ }
            else certain = false;                            // too close (or NaN/inf/negative fold): exact path
            if (certain && include + (all - seen) < cfg.min_include) fail = true;   // cannot reach `ratio` any more
        }
        if (certain) {
            if (fail) return 0;
            uint8_t fl = 0;
            if (include >= cfg.min_include) {                 // <=> !(1.0*include/all < ratio)
                fl |= 1;
                const double t = (double)tot;
                if (!(t < cfg.min_freq || t > cfg.max_freq)) fl |= 2;
            }
            return fl;
        }
    }
    int include = 0, seen = 0;
    bool decided_fail = false;
    for (int s = 0; s < cfg.n_sets; s++) {
        const int g0 = cfg.set_off[s], g1 = cfg.set_off[s + 1];
        const int ng = g1 - g0;
        if (ng < 2) continue;
        seen++;
        double f[MX_MAX_GROUPS_PER_SET];
        for (int g = g0; g < g1; g++) {
            uint64_t cs = 0, ls = 0;
            for (int m = cfg.grp_off[g]; m < cfg.grp_off[g + 1]; m++) {
                const int c = cfg.members[m];
                cs += row[c];
                ls += lengths[c];
            }
            // count/lens or sum(count)/sum(lens); by_count: the raw (summed) count
            if (cfg.by_count) f[g - g0] = (double)cs;
            else f[g - g0] = (cs == 0 && ls != 0) ? 0.0 : (double)cs / (double)ls;
        }
        // sorted(freqs, reverse=1): insertion sort, descending
        for (int i = 1; i < ng; i++) {
            const double v = f[i];
            int j = i - 1;
            while (j >= 0 && f[j] < v) {
                f[j + 1] = f[j];
                j--;
            }
            f[j + 1] = v;
        }
        const double fmax = f[0];
        const double fmin = f[cfg.baseline >= 0 ? cfg.baseline : ng + cfg.baseline];
        if (1.0 * fmax / (fmin + 1e-20) >= cfg.min_fold) include++;
        if (1.0 * (double)(include + (all - seen)) / (double)all < cfg.ratio) {
            decided_fail = true;
            break;
        }
    }
    uint8_t fl = 0;
    const double rr = 1.0 * (double)include / (double)all;
    if (!decided_fail && !(rr < cfg.ratio)) {
        fl |= 1;
        const double t = (double)tot;
        if (!(t < cfg.min_freq || t > cfg.max_freq)) fl |= 2;
    }
    return fl;
}
