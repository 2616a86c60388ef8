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
};

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
