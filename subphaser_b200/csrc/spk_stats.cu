// K10: batched fp64 Fisher exact test (right tail), the enrichment decision, and Benjamini-Hochberg.
//
// Replaces Stats.fisher_test (Stats.py:14-31; third-party `fisher` 0.1.9 pvalue(...).right_tail),
// Pvalues.get_enriched + _enrich (Stats.py:150-192) and correct_pvals (Stats.py:11-12; statsmodels
// multipletests(method='fdr_bh')).  Compiled with -fmad=false: the ratios / decisions reproduce the
// numpy/Python operation order exactly.
//
// Hypergeometric point masses come from Loader's saddle-point expansion ("Fast and accurate
// computation of binomial probabilities", 2000; the algorithm behind R's dhyper): accurate to ~1e-15
// relative for table margins up to 2^31, where lgamma differences would lose 6 digits.  The tail is
// then summed with the term-ratio recurrence away from the mode until terms no longer change the sum.
#include <math.h>
#include "spk_common.cuh"

namespace {

constexpr int64_t FISHER_MAX_INT = 2147483647 / 10;  // Stats.py:9

__device__ __forceinline__ double stirlerr(double n) {
    // log(n!) - log(sqrt(2*pi*n) * (n/e)^n), integer n >= 0
    const double S0 = 0.083333333333333333333, S1 = 0.00277777777777777777778,
                 S2 = 0.00079365079365079365079365, S3 = 0.000595238095238095238095238,
                 S4 = 0.0008417508417508417508417508;
    const double sfe[16] = {0.0,
                            0.081061466795327258219670264,
                            0.041340695955409294093822081,
                            0.0276779256849983391487892927,
                            0.020790672103765093111522771,
                            0.016644691189821192163194865,
                            0.013876128823070747998745727,
                            0.011896709945891770095055724,
                            0.010411265261972096497478567,
                            0.0092554621827127329177286366,
                            0.0083305634333628712564693187,
                            0.0075736754879518407949720242,
                            0.0069428401072095298656641527,
                            0.0064089941880042070684396311,
                            0.0059513701127588477356244160,
                            0.0055547335519628013710386900};
    if (n <= 15.0) return sfe[(int)n];
    const double nn = n * n;
    if (n > 500) return (S0 - S1 / nn) / n;
    if (n > 80) return (S0 - (S1 - S2 / nn) / nn) / n;
    if (n > 35) return (S0 - (S1 - (S2 - S3 / nn) / nn) / nn) / n;
    return (S0 - (S1 - (S2 - (S3 - S4 / nn) / nn) / nn) / nn) / n;
}

__device__ __forceinline__ double bd0(double x, double np) {
    // x*log(x/np) + np - x, stable for x ~ np
    if (fabs(x - np) < 0.1 * (x + np)) {
        double v = (x - np) / (x + np);
        double s = (x - np) * v;
        if (fabs(s) < 2.2250738585072014e-308) return s;
        double ej = 2 * x * v;
        v = v * v;
        for (int j = 1; j < 1000; j++) {
            ej *= v;
            const double s1 = s + ej / ((j << 1) + 1);
            if (s1 == s) return s1;
            s = s1;
        }
    }
    return x * log(x / np) + np - x;
}

// log of the binomial point mass b(x; n, p) with q = 1-p given separately
__device__ double dbinom_raw_log(double x, double n, double p, double q) {
    const double NEG_INF = -INFINITY;
    if (p == 0) return (x == 0) ? 0.0 : NEG_INF;
    if (q == 0) return (x == n) ? 0.0 : NEG_INF;
    if (x == 0) {
        if (n == 0) return 0.0;
        return (p < 0.1) ? -bd0(n, n * q) - n * p : n * log(q);
    }
    if (x == n) return (q < 0.1) ? -bd0(n, n * p) - n * q : n * log(p);
    if (x < 0 || x > n) return NEG_INF;
    const double lc = stirlerr(n) - stirlerr(x) - stirlerr(n - x) - bd0(x, n * p) - bd0(n - x, n * q);
    const double lf = 1.8378770664093454835606594728112 + log(x) + log1p(-x / n);
    return lc - 0.5 * lf;
}

// hypergeometric pmf: x white in n draws from r white + b black
__device__ double dhyper(double x, double r, double b, double n) {
    if (n < x || r < x || n - x > b) return 0.0;
    if (n == 0) return (x == 0) ? 1.0 : 0.0;
    const double p = n / (r + b);
    const double q = (r + b - n) / (r + b);
    const double l1 = dbinom_raw_log(x, r, p, q);
    const double l2 = dbinom_raw_log(n - x, b, p, q);
    const double l3 = dbinom_raw_log(n, r + b, p, q);
    return exp(l1 + l2 - l3);
}

// P(X >= x11) for the 2x2 table [[x11, x12], [x21, x22]]
__device__ double fisher_right_tail(int64_t x11, int64_t x12, int64_t x21, int64_t x22) {
    const double K = (double)(x11 + x21);  // white balls (column 1)
    const double n = (double)(x11 + x12);  // draws (row 1)
    const double N = (double)(x11 + x12 + x21 + x22);
    const double B = N - K;
    const double lo = fmax(0.0, n - B);
    const double hi = fmin(n, K);
    const double x0 = (double)x11;
    if (x0 <= lo) return 1.0;
    if (x0 > hi) return 0.0;
    const double mode = floor((n + 1.0) * (K + 1.0) / (N + 2.0));
    if (x0 > mode) {
        double t = dhyper(x0, K, B, n);
        double s = t;
        for (double x = x0; x < hi; x += 1.0) {
            t *= ((K - x) / (x + 1.0)) * ((n - x) / (B - n + x + 1.0));
            const double s1 = s + t;
            if (s1 == s) break;
            s = s1;
        }
        return fmin(s, 1.0);
    }
    // lower side: p = 1 - P(X <= x11-1), summed downwards from x11-1
    double x = x0 - 1.0;
    double t = dhyper(x, K, B, n);
    double s = t;
    for (; x > lo; x -= 1.0) {
        t *= (x / (K - x + 1.0)) * ((B - n + x) / (n - x + 1.0));
        const double s1 = s + t;
        if (s1 == s) break;
        s = s1;
    }
    return fmax(0.0, 1.0 - s);
}

__global__ void __launch_bounds__(128)
k_fisher(const int64_t* __restrict__ counts, const int64_t* __restrict__ totals, uint64_t W, int S,
         double* __restrict__ pvals) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= W * (uint64_t)S) return;
    const uint64_t r = t / S;
    const int i = (int)(t % S);
    int64_t sum_each = 0, sum_total = 0;
    for (int c = 0; c < S; c++) {
        sum_each += counts[r * S + c];
        sum_total += totals[c];
    }
    const int64_t x11 = counts[r * S + i];
    const int64_t x12 = sum_each - x11;
    int64_t x21 = totals[i] - x11;
    int64_t x22 = sum_total - x21 - x12;  // as coded in Stats.py:23
    x21 = min(x21, FISHER_MAX_INT);
    x22 = min(x22, FISHER_MAX_INT);
    pvals[t] = fisher_right_tail(x11, x12, x21, x22);
}

// Self-check hook: total probability mass of Hypergeom(N, K, n) summed with the same point-mass routine
// and term recurrences the Fisher kernel uses (from the mode outwards).  Must be 1 to ~1e-13; used by
// the tests to establish the kernel's absolute accuracy where no exact reference is computable.
__global__ void k_hypergeom_mass(const int64_t* __restrict__ NKn, uint64_t count, double* __restrict__ out) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count) return;
    const double N = (double)NKn[3 * t], K = (double)NKn[3 * t + 1], n = (double)NKn[3 * t + 2];
    const double B = N - K;
    const double lo = fmax(0.0, n - B), hi = fmin(n, K);
    double mode = floor((n + 1.0) * (K + 1.0) / (N + 2.0));
    mode = fmin(fmax(mode, lo), hi);
    const double t0 = dhyper(mode, K, B, n);
    double s = t0, tt = t0;
    for (double x = mode; x < hi; x += 1.0) {
        tt *= ((K - x) / (x + 1.0)) * ((n - x) / (B - n + x + 1.0));
        const double s1 = s + tt;
        if (s1 == s) break;
        s = s1;
    }
    tt = t0;
    for (double x = mode; x > lo; x -= 1.0) {
        tt *= (x / (K - x + 1.0)) * ((B - n + x) / (n - x + 1.0));
        const double s1 = s + tt;
        if (s1 == s) break;
        s = s1;
    }
    out[t] = s;
}

// numpy's pairwise summation for a contiguous run of n < 128 doubles (np.add.reduce)
__device__ __forceinline__ double np_sum_small(const double* a, int n) {
    if (n < 8) {
        double res = 0.0;
        for (int i = 0; i < n; i++) res += a[i];
        return res;
    }
    double r[8];
    for (int j = 0; j < 8; j++) r[j] = a[j];
    int i;
    for (i = 8; i < n - (n % 8); i += 8)
        for (int j = 0; j < 8; j++) r[j] += a[i + j];
    double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; i++) res += a[i];
    return res;
}

constexpr int EN_MAX_S = 64;

__global__ void __launch_bounds__(128)
k_enrich_rows(const int64_t* __restrict__ counts, const int64_t* __restrict__ totals,
              const double* __restrict__ pvals, uint64_t W, int S, double max_pval, double cutoff,
              double min_ratio, int32_t* __restrict__ idx_out, uint8_t* __restrict__ sig_out,
              double* __restrict__ ratios_out, double* __restrict__ pmin_out) {
    const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= W) return;
    const double* p = pvals + r * S;
    // stable sort by p: first two entries
    int i0 = 0;
    for (int i = 1; i < S; i++)
        if (p[i] < p[i0]) i0 = i;
    int i1 = -1;
    for (int i = 0; i < S; i++) {
        if (i == i0) continue;
        if (i1 < 0 || p[i] < p[i1]) i1 = i;
    }
    const double pmin = p[i0], psub = p[i1];
    bool sig = true;
    if (pmin > max_pval) sig = false;
    if (pmin == 0) {
    } else if (psub / pmin < max_pval / psub * cutoff) sig = false;
    double ra[EN_MAX_S];
    for (int c = 0; c < S; c++) ra[c] = (double)counts[r * S + c] / (double)totals[c];
    const double rs = np_sum_small(ra, S);
    for (int c = 0; c < S; c++) {
        ra[c] = ra[c] / rs;
        ratios_out[r * S + c] = ra[c];
    }
    if (ra[i0] < min_ratio) sig = false;
    idx_out[r] = i0;
    sig_out[r] = sig ? 1 : 0;
    pmin_out[r] = pmin;
}

// column sums of the window x subgenome count matrix (Stats.py:145 `arr.sum(axis=0)`), exact integers
__global__ void __launch_bounds__(256)
k_colsum_i64(const int64_t* __restrict__ counts, uint64_t W, int S, int64_t* __restrict__ totals) {
    __shared__ long long s_part[256];
    const int c = blockIdx.x;
    long long acc = 0;
    for (uint64_t r = threadIdx.x; r < W; r += blockDim.x) acc += counts[r * S + c];
    s_part[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) s_part[threadIdx.x] += s_part[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) totals[c] = s_part[0];
}

// ---- Benjamini-Hochberg (statsmodels fdr_bh): q_(i) = p_(i) / (i/n), reverse running min, clip 1 ----
__global__ void k_bh_keys(const double* __restrict__ p, uint64_t n, uint64_t* __restrict__ keys,
                          uint32_t* __restrict__ idx) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (uint64_t)gridDim.x * blockDim.x) {
        keys[i] = (uint64_t)__double_as_longlong(p[i]);  // p >= 0: bit pattern is order preserving
        idx[i] = (uint32_t)i;
    }
}

// single CTA: reverse inclusive min-scan of p_sorted[i] / ((i+1)/n), then scatter to original order
__global__ void __launch_bounds__(1024)
k_bh_finish(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ idx, uint64_t n,
            double* __restrict__ q) {
    __shared__ double s_warp[32];
    __shared__ double s_carry;
    if (threadIdx.x == 0) s_carry = INFINITY;
    __syncthreads();
    const double dn = (double)n;
    for (uint64_t base = 0; base < n; base += 1024) {
        const uint64_t jrev = base + threadIdx.x;  // position from the end
        const bool ok = jrev < n;
        const uint64_t i = ok ? (n - 1 - jrev) : 0;
        double v = INFINITY;
        if (ok) {
            const double ps = __longlong_as_double((long long)keys[i]);
            const double ecdf = (double)(i + 1) / dn;
            v = ps / ecdf;
        }
        double incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double t = __shfl_up_sync(0xffffffffu, incl, o);
            if ((threadIdx.x & 31) >= o) incl = fmin(incl, t);
        }
        if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = incl;
        __syncthreads();
        double prefix = s_carry;
        for (int w = 0; w < (int)(threadIdx.x >> 5); w++) prefix = fmin(prefix, s_warp[w]);
        const double res = fmin(prefix, incl);
        if (ok) q[idx[i]] = res > 1.0 ? 1.0 : res;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = res;
        __syncthreads();
    }
}

inline size_t al256(size_t x) { return (x + 255) / 256 * 256; }

}  // namespace

extern "C" int spk_fisher_right_tail(const int64_t* d_counts, const int64_t* d_totals, uint64_t W,
                                     int S, double* d_pvals, void* stream) {
    SPK_CHECK_ARG(S >= 1, "S must be >= 1");
    if (W == 0) return SPK_OK;
    SPK_CHECK_ARG(d_counts && d_totals && d_pvals, "null pointer");
    const uint64_t n = W * (uint64_t)S;
    k_fisher<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(d_counts, d_totals, W, S,
                                                                          d_pvals);
    SPK_LAUNCH_CHECK();
    return SPK_OK;
}

extern "C" int spk_enrich_rows(const int64_t* d_counts, const int64_t* d_totals, const double* d_pvals,
                               uint64_t W, int S, double max_pval, double cutoff, double min_ratio,
                               int32_t* d_idx, uint8_t* d_sig, double* d_ratios, double* d_pmin,
                               void* stream) {
    SPK_CHECK_ARG(S >= 2 && S <= EN_MAX_S, "S must be in [2, 64] (Stats.py:172 asserts > 1)");
    if (W == 0) return SPK_OK;
    SPK_CHECK_ARG(d_counts && d_totals && d_pvals && d_idx && d_sig && d_ratios && d_pmin, "null pointer");
    k_enrich_rows<<<(unsigned)((W + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
        d_counts, d_totals, d_pvals, W, S, max_pval, cutoff, min_ratio, d_idx, d_sig, d_ratios, d_pmin);
    SPK_LAUNCH_CHECK();
    return SPK_OK;
}

extern "C" int spk_debug_hypergeom_mass(const int64_t* d_NKn, uint64_t count, double* d_out, void* stream) {
    if (count == 0) return SPK_OK;
    SPK_CHECK_ARG(d_NKn && d_out, "null pointer");
    k_hypergeom_mass<<<(unsigned)((count + 63) / 64), 64, 0, (cudaStream_t)stream>>>(d_NKn, count, d_out);
    SPK_LAUNCH_CHECK();
    return SPK_OK;
}

extern "C" int spk_colsum_i64(const int64_t* d_counts, uint64_t W, int S, int64_t* d_totals, void* stream) {
    SPK_CHECK_ARG(S >= 1 && d_totals, "bad arguments");
    SPK_CHECK_ARG(W == 0 || d_counts, "null counts");
    k_colsum_i64<<<S, 256, 0, (cudaStream_t)stream>>>(d_counts, W, S, d_totals);
    SPK_LAUNCH_CHECK();
    return SPK_OK;
}

extern "C" size_t spk_bh_workspace_bytes(uint64_t n) {
    return al256(n * 8) * 2 + al256(n * 4) * 2 + al256(spk_sort_workspace_bytes(n)) + 256;
}

extern "C" int spk_bh_adjust(const double* d_p, double* d_q, uint64_t n, void* d_ws, size_t ws_bytes,
                             void* stream) {
    if (n == 0) return SPK_OK;
    SPK_CHECK_ARG(d_p && d_q && d_ws, "null pointer");
    SPK_CHECK_ARG(n < 0xffffffffull, "n too large");
    if (ws_bytes < spk_bh_workspace_bytes(n)) {
        spk_set_error("spk_bh_adjust: workspace too small");
        return SPK_ECAP;
    }
    char* w = (char*)d_ws;
    uint64_t* keys = (uint64_t*)w;
    w += al256(n * 8);
    uint64_t* keys_tmp = (uint64_t*)w;
    w += al256(n * 8);
    uint32_t* idx = (uint32_t*)w;
    w += al256(n * 4);
    uint32_t* idx_tmp = (uint32_t*)w;
    w += al256(n * 4);
    void* sort_ws = w;
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned grid = (unsigned)min((n + 255) / 256, (uint64_t)spk_num_sms() * 8);
    k_bh_keys<<<grid, 256, 0, st>>>(d_p, n, keys, idx);
    SPK_LAUNCH_CHECK();
    int rc = spk_sort_pairs_u64(keys, idx, keys_tmp, idx_tmp, n, 64, sort_ws,
                                spk_sort_workspace_bytes(n), stream);
    if (rc != SPK_OK) return rc;
    k_bh_finish<<<1, 1024, 0, st>>>(keys, idx, n, d_q);
    SPK_LAUNCH_CHECK();
    return SPK_OK;
}
