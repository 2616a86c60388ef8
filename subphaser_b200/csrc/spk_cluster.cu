// K5-K8: statistics over the chromosome x k-mer matrix, all fp64 (compiled with -fmad=false).
//
// Replaces Cluster.normalize_data / fit / bootstrap / _output_kmers / pca (Cluster.py:48-118,178-194)
// and the sklearn / scipy calls behind them.
//
// Design: the reference clusters n chromosomes (n ~ 20) living in M dimensions (M = #k-mers, up to
// 1e7).  Every quantity K-Means and PCA need about n points and their centroids (centroid = mean of
// member points) follows from the n x n Gram matrix G = Z Z^T:
//     |z_i - c_J|^2 = G_ii - 2/|J| sum_{l in J} G_il + 1/|J|^2 sum_{l,m in J} G_lm
// so one streaming pass over Z (8 n M bytes, the HBM roofline of this stage) replaces every Lloyd
// iteration's pass, and the 1000 bootstrap replicates become 1000 tiny Gram problems.
#include <math.h>
#include <stdlib.h>
#include "spk_common.cuh"

namespace {

constexpr int CL_MAX_N = 128;   // chromosomes
constexpr int CL_MAX_S = 16;    // clusters / subgenomes
constexpr int GR_THREADS = 512;
constexpr int GR_TILE_ROWS = 32;
constexpr int GR_MAX_ACC = 17;  // ceil(128*129/2 / 512)

// numpy's pairwise summation of a contiguous run (np.add.reduce along a contiguous axis) for
// n <= 128 = CL_MAX_N (above 128 numpy recurses on halves; not needed here).  No device recursion:
// these kernels keep per-thread arrays in local memory and must have a static stack size.
__device__ __forceinline__ double np_pairwise(const double* a, int n) {
    if (n < 8) {
        double res = 0.0;
        for (int i = 0; i < n; i++) res += a[i];
        return res;
    }
    double r[8];
#pragma unroll
    for (int j = 0; j < 8; j++) r[j] = a[j];
    int i;
    for (i = 8; i < n - (n % 8); i += 8) {
#pragma unroll
        for (int j = 0; j < 8; j++) r[j] += a[i + j];
    }
    double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; i++) res += a[i];
    return res;
}

// ---- K5 z-score ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_zscore_rows(const double* __restrict__ X, uint64_t M, int n, double* __restrict__ Z) {
    double buf[CL_MAX_N];
    for (uint64_t m = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; m < M;
         m += (uint64_t)gridDim.x * blockDim.x) {
        const double* x = X + m * n;
        for (int c = 0; c < n; c++) buf[c] = x[c];
        const double mean = np_pairwise(buf, n) / (double)n;
        for (int c = 0; c < n; c++) {
            const double d = x[c] - mean;
            buf[c] = d * d;
        }
        const double sd = sqrt(np_pairwise(buf, n) / (double)n);
        for (int c = 0; c < n; c++) Z[m * n + c] = (x[c] - mean) / sd;
    }
}

// ---- Gram ------------------------------------------------------------------------------------------
// pair p -> (a, b), a <= b, row-major over the upper triangle
__device__ __forceinline__ void pair_of(int p, int n, int& a, int& b) {
    int row = 0, rem = p;
    while (rem >= n - row) {
        rem -= n - row;
        row++;
    }
    a = row;
    b = row + rem;
}

// Accumulate the upper triangle of sum_r z_r^T z_r over the rows [r0, r1) (optionally gathered through
// idx) into per-thread registers, tile by tile through shared memory; then store to out[n*n].
__device__ void gram_accumulate(const double* __restrict__ Z, int n, const uint32_t* __restrict__ idx,
                                uint64_t r0, uint64_t r1, double* s_tile, double* __restrict__ out) {
    const int npairs = n * (n + 1) / 2;
    double acc[GR_MAX_ACC];
    int pa[GR_MAX_ACC], pb[GR_MAX_ACC];
#pragma unroll
    for (int q = 0; q < GR_MAX_ACC; q++) {
        acc[q] = 0.0;
        const int p = threadIdx.x + q * GR_THREADS;
        pa[q] = pb[q] = 0;
        if (p < npairs) pair_of(p, n, pa[q], pb[q]);
    }
    for (uint64_t base = r0; base < r1; base += GR_TILE_ROWS) {
        const int rows = (int)min((uint64_t)GR_TILE_ROWS, r1 - base);
        __syncthreads();
        for (int e = threadIdx.x; e < rows * n; e += GR_THREADS) {
            const int rr = e / n, cc = e - rr * n;
            const uint64_t src = idx ? (uint64_t)idx[base + rr] : (base + rr);
            s_tile[e] = Z[src * n + cc];
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < GR_MAX_ACC; q++) {
            if (threadIdx.x + q * GR_THREADS < npairs) {
                double a = acc[q];
                for (int rr = 0; rr < rows; rr++) a += s_tile[rr * n + pa[q]] * s_tile[rr * n + pb[q]];
                acc[q] = a;
            }
        }
    }
#pragma unroll
    for (int q = 0; q < GR_MAX_ACC; q++) {
        if (threadIdx.x + q * GR_THREADS < npairs) {
            out[pa[q] * n + pb[q]] = acc[q];
            out[pb[q] * n + pa[q]] = acc[q];
        }
    }
}

__global__ void __launch_bounds__(GR_THREADS)
k_gram_partial(const double* __restrict__ Z, uint64_t nrows, int n, const uint32_t* __restrict__ idx,
               double* __restrict__ partial) {
    __shared__ double s_tile[GR_TILE_ROWS * CL_MAX_N];
    const uint64_t per = (nrows + gridDim.x - 1) / gridDim.x;
    const uint64_t r0 = min((uint64_t)blockIdx.x * per, nrows);
    const uint64_t r1 = min(r0 + per, nrows);
    gram_accumulate(Z, n, idx, r0, r1, s_tile, partial + (uint64_t)blockIdx.x * n * n);
}

__global__ void k_gram_reduce(const double* __restrict__ partial, int nblocks, int nn,
                              double* __restrict__ G) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nn) return;
    double s = 0.0;
    for (int b = 0; b < nblocks; b++) s += partial[(uint64_t)b * nn + e];  // fixed order: deterministic
    G[e] = s;
}

// one CTA per replicate: Gram over the B gathered rows idx[r*B .. r*B+B)
__global__ void __launch_bounds__(GR_THREADS)
k_gram_batched(const double* __restrict__ Z, int n, const uint32_t* __restrict__ idx, int B,
               double* __restrict__ G) {
    __shared__ double s_tile[GR_TILE_ROWS * CL_MAX_N];
    const uint64_t r = blockIdx.x;
    gram_accumulate(Z, n, idx + r * B, 0, (uint64_t)B, s_tile, G + r * n * n);
}

// ---- K6 K-Means on a Gram matrix (one thread per problem) -------------------------------------------
struct Rng {
    uint64_t s;
    __device__ uint64_t next() {
        s += 0x9E3779B97F4A7C15ull;
        uint64_t z = s;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    __device__ double uniform() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
};

struct KmScratch {
    double T[CL_MAX_S][CL_MAX_N];  // T[j][i] = sum_{l in C_j} G[i][l]
    double Q[CL_MAX_S];            // sum_{l,m in C_j} G[l][m]
    int cnt[CL_MAX_S];
    int labels[CL_MAX_N];
    int prev[CL_MAX_N];
    double closest[CL_MAX_N];
};

__device__ __forceinline__ double km_dist(const double* G, int n, const KmScratch& w, int i, int j) {
    const double c = (double)w.cnt[j];
    return G[i * n + i] - 2.0 * w.T[j][i] / c + w.Q[j] / (c * c);
}

__device__ void km_update(const double* G, int n, int S, KmScratch& w) {
    for (int j = 0; j < S; j++) {
        w.cnt[j] = 0;
        w.Q[j] = 0.0;
        for (int i = 0; i < n; i++) w.T[j][i] = 0.0;
    }
    for (int l = 0; l < n; l++) {
        const int j = w.labels[l];
        w.cnt[j]++;
        for (int i = 0; i < n; i++) w.T[j][i] += G[i * n + l];
    }
    for (int l = 0; l < n; l++) w.Q[w.labels[l]] += w.T[w.labels[l]][l];
}

// one full K-Means run (k-means++ seeding + Lloyd to a fixed point); returns inertia
__device__ double km_run(const double* G, int n, int S, int max_iter, Rng& rng, KmScratch& w) {
    // ---- k-means++ (Arthur & Vassilvitskii; greedy variant with 2+log(S) local trials as sklearn) ----
    int centers[CL_MAX_S];
    const int trials = 2 + (int)log((double)S);
    centers[0] = min((int)(rng.uniform() * n), n - 1);
    double pot = 0.0;
    for (int i = 0; i < n; i++) {
        const int c = centers[0];
        w.closest[i] = fmax(G[i * n + i] - 2.0 * G[i * n + c] + G[c * n + c], 0.0);
        pot += w.closest[i];
    }
    for (int cix = 1; cix < S; cix++) {
        int best_c = -1;
        double best_pot = 0.0;
        for (int t = 0; t < trials; t++) {
            const double rv = rng.uniform() * pot;
            double cum = 0.0;
            int cand = n - 1;
            for (int i = 0; i < n; i++) {
                cum += w.closest[i];
                if (cum > rv) {
                    cand = i;
                    break;
                }
            }
            double np = 0.0;
            for (int i = 0; i < n; i++) {
                const double d = fmax(G[i * n + i] - 2.0 * G[i * n + cand] + G[cand * n + cand], 0.0);
                np += fmin(w.closest[i], d);
            }
            if (best_c < 0 || np < best_pot) {
                best_c = cand;
                best_pot = np;
            }
        }
        centers[cix] = best_c;
        pot = best_pot;
        for (int i = 0; i < n; i++) {
            const double d = fmax(G[i * n + i] - 2.0 * G[i * n + best_c] + G[best_c * n + best_c], 0.0);
            w.closest[i] = fmin(w.closest[i], d);
        }
    }
    // centres are single points: express them as one-member clusters
    for (int j = 0; j < S; j++) {
        w.cnt[j] = 1;
        const int c = centers[j];
        w.Q[j] = G[c * n + c];
        for (int i = 0; i < n; i++) w.T[j][i] = G[i * n + c];
    }
    for (int i = 0; i < n; i++) w.prev[i] = -1;
    // ---- Lloyd ----
    for (int it = 0; it < max_iter; it++) {
        bool same = true;
        for (int i = 0; i < n; i++) {
            int bj = 0;
            double bd = km_dist(G, n, w, i, 0);
            for (int j = 1; j < S; j++) {
                const double d = km_dist(G, n, w, i, j);
                if (d < bd) {
                    bd = d;
                    bj = j;
                }
            }
            w.labels[i] = bj;
            w.closest[i] = bd;
            if (bj != w.prev[i]) same = false;
        }
        // empty clusters take the points farthest from their centres (sklearn _relocate_empty_clusters)
        int cnt[CL_MAX_S];
        for (int j = 0; j < S; j++) cnt[j] = 0;
        for (int i = 0; i < n; i++) cnt[w.labels[i]]++;
        for (int j = 0; j < S; j++) {
            if (cnt[j] == 0) {
                int far = -1;
                for (int i = 0; i < n; i++)
                    if (cnt[w.labels[i]] > 1 && (far < 0 || w.closest[i] > w.closest[far])) far = i;
                if (far >= 0) {
                    cnt[w.labels[far]]--;
                    w.labels[far] = j;
                    cnt[j] = 1;
                    w.closest[far] = -1.0;
                    same = false;
                }
            }
        }
        km_update(G, n, S, w);
        if (same) break;
        for (int i = 0; i < n; i++) w.prev[i] = w.labels[i];
    }
    double inertia = 0.0;
    for (int i = 0; i < n; i++) inertia += fmax(km_dist(G, n, w, i, w.labels[i]), 0.0);
    return inertia;
}

// Cluster.sort_subgenomes (Cluster.py:119-126): relabel by first appearance over name-sorted chromosomes
__device__ void canonical_relabel(const int* labels, int n, int S, const int32_t* order, int32_t* out) {
    int map[CL_MAX_S];
    for (int j = 0; j < S; j++) map[j] = -1;
    int next = 0;
    for (int t = 0; t < n; t++) {
        const int i = order ? order[t] : t;
        if (map[labels[i]] < 0) map[labels[i]] = next++;
    }
    for (int i = 0; i < n; i++) out[i] = map[labels[i]];
}

__global__ void __launch_bounds__(32)
k_kmeans_gram(const double* __restrict__ G, int R, int n, int S, int n_init, int max_iter,
              uint64_t seed, int r0, const int32_t* __restrict__ order, int32_t* __restrict__ labels_out,
              double* __restrict__ inertia_out, KmScratch* __restrict__ scratch) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    const double* g = G + (uint64_t)r * n * n;
    KmScratch& w = scratch[r];
    // (the random stream belongs to the GLOBAL replicate number r0 + r: ranks that share the replicates of a
    //  bootstrap get the results one rank would)
    Rng rng{seed ^ (0xD1B54A32D192ED03ull * (uint64_t)(r0 + r + 1))};
    int best[CL_MAX_N];
    double best_inertia = INFINITY;
    for (int t = 0; t < n_init; t++) {
        const double inertia = km_run(g, n, S, max_iter, rng, w);
        if (t == 0 || inertia < best_inertia) {
            best_inertia = inertia;
            for (int i = 0; i < n; i++) best[i] = w.labels[i];
        }
    }
    canonical_relabel(best, n, S, order, labels_out + (uint64_t)r * n);
    if (inertia_out) inertia_out[r] = best_inertia;
}

// ---- ARI / V-measure (sklearn.metrics.adjusted_rand_score / v_measure_score) -------------------------
__global__ void k_cluster_scores(const int32_t* __restrict__ ref, const int32_t* __restrict__ labels,
                                 int R, int n, double* __restrict__ ari, double* __restrict__ vm) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    const int32_t* lab = labels + (uint64_t)r * n;
    long long cont[CL_MAX_S][CL_MAX_S];
    long long nc[CL_MAX_S], nk[CL_MAX_S];
    for (int a = 0; a < CL_MAX_S; a++) {
        nc[a] = nk[a] = 0;
        for (int b = 0; b < CL_MAX_S; b++) cont[a][b] = 0;
    }
    for (int i = 0; i < n; i++) {
        cont[ref[i]][lab[i]]++;
        nc[ref[i]]++;
        nk[lab[i]]++;
    }
    long long ss = 0, snk = 0, snc = 0;
    for (int a = 0; a < CL_MAX_S; a++) {
        snc += nc[a] * nc[a];
        snk += nk[a] * nk[a];
        for (int b = 0; b < CL_MAX_S; b++) ss += cont[a][b] * cont[a][b];
    }
    const long long N = n;
    const long long tp = ss - N, fp = snk - ss, fn = snc - ss, tn = N * N - fp - fn - ss;
    if (fn == 0 && fp == 0) ari[r] = 1.0;
    else ari[r] = 2.0 * (double)(tp * tn - fn * fp) / (double)((tp + fn) * (fn + tn) + (tp + fp) * (fp + tn));
    // entropies and mutual information (natural log), as sklearn computes them
    double hC = 0.0, hK = 0.0;
    const double dn = (double)n;
    for (int a = 0; a < CL_MAX_S; a++) {
        if (nc[a] > 0) hC -= ((double)nc[a] / dn) * (log((double)nc[a]) - log(dn));
        if (nk[a] > 0) hK -= ((double)nk[a] / dn) * (log((double)nk[a]) - log(dn));
    }
    double mi = 0.0;
    for (int a = 0; a < CL_MAX_S; a++)
        for (int b = 0; b < CL_MAX_S; b++)
            if (cont[a][b] > 0) {
                const double nij = (double)cont[a][b];
                const double lo = -log((double)nc[a] * (double)nk[b]) + log(dn) + log(dn);
                double term = (nij / dn) * (log(nij) - log(dn)) + (nij / dn) * lo;
                if (fabs(term) < 2.220446049250313e-16) term = 0.0;
                mi += term;
            }
    if (mi < 0) mi = 0;
    const double h = (hC != 0.0) ? mi / hC : 1.0;
    const double c = (hK != 0.0) ? mi / hK : 1.0;
    vm[r] = (h + c == 0.0) ? 0.0 : (2.0 * h * c / (h + c));
}

// ---- centroids -----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_centroids(const double* __restrict__ Z, uint64_t M, int n, const int32_t* __restrict__ labels, int S,
            double* __restrict__ C) {
    for (uint64_t m = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; m < M;
         m += (uint64_t)gridDim.x * blockDim.x) {
        double sum[CL_MAX_S];
        int cnt[CL_MAX_S];
        for (int j = 0; j < S; j++) {
            sum[j] = 0.0;
            cnt[j] = 0;
        }
        for (int c = 0; c < n; c++) {
            sum[labels[c]] += Z[m * n + c];
            cnt[labels[c]]++;
        }
        for (int j = 0; j < S; j++) C[(uint64_t)j * M + m] = sum[j] / (double)cnt[j];
    }
}

// ---- K7 Student t-test of the two highest-mean groups ---------------------------------------------------
__device__ double betacf(double a, double b, double x) {
    const double FPMIN = 1e-300, EPS = 1e-16;
    const double qab = a + b, qap = a + 1.0, qam = a - 1.0;
    double c = 1.0, d = 1.0 - qab * x / qap;
    if (fabs(d) < FPMIN) d = FPMIN;
    d = 1.0 / d;
    double h = d;
    for (int m = 1; m <= 500; m++) {
        const int m2 = 2 * m;
        double aa = m * (b - m) * x / ((qam + m2) * (a + m2));
        d = 1.0 + aa * d;
        if (fabs(d) < FPMIN) d = FPMIN;
        c = 1.0 + aa / c;
        if (fabs(c) < FPMIN) c = FPMIN;
        d = 1.0 / d;
        h *= d * c;
        aa = -(a + m) * (qab + m) * x / ((a + m2) * (qap + m2));
        d = 1.0 + aa * d;
        if (fabs(d) < FPMIN) d = FPMIN;
        c = 1.0 + aa / c;
        if (fabs(c) < FPMIN) c = FPMIN;
        d = 1.0 / d;
        const double del = d * c;
        h *= del;
        if (fabs(del - 1.0) < EPS) break;
    }
    return h;
}

// regularised incomplete beta I_x(a, b) with y = 1 - x supplied exactly
__device__ double betainc_xy(double a, double b, double x, double y) {
    if (x <= 0.0) return 0.0;
    if (y <= 0.0) return 1.0;
    const double lfront = lgamma(a + b) - lgamma(a) - lgamma(b) + a * log(x) + b * log(y);
    if (x < (a + 1.0) / (a + b + 2.0)) return exp(lfront) * betacf(a, b, x) / a;
    return 1.0 - exp(lfront) * betacf(b, a, y) / b;
}

__global__ void __launch_bounds__(128)
k_ttest_groups(const double* __restrict__ X, uint64_t M, int n, const int32_t* __restrict__ col_group,
               int S, int32_t* __restrict__ best, double* __restrict__ pval,
               double* __restrict__ means, int singleton_nan) {
    double vals[CL_MAX_N];   // values regrouped: group g occupies [off[g], off[g+1])
    int off[CL_MAX_S + 1];
    for (uint64_t m = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; m < M;
         m += (uint64_t)gridDim.x * blockDim.x) {
        const double* x = X + m * n;
        int pos = 0;
        for (int g = 0; g < S; g++) {
            off[g] = pos;
            for (int c = 0; c < n; c++)
                if (col_group[c] == g) vals[pos++] = x[c];
        }
        off[S] = pos;
        // np.mean per group (reported) and python sum()/len (sort key)
        double npmean[CL_MAX_S], key[CL_MAX_S];
        for (int g = 0; g < S; g++) {
            const int ng = off[g + 1] - off[g];
            npmean[g] = np_pairwise(vals + off[g], ng) / (double)ng;
            double s = 0.0;
            for (int i = off[g]; i < off[g + 1]; i++) s += vals[i];
            key[g] = -s / (double)ng;
            means[m * S + g] = npmean[g];
        }
        // stable ascending sort by key: pick the first two
        int g0 = 0;
        for (int g = 1; g < S; g++)
            if (key[g] < key[g0]) g0 = g;
        int g1 = -1;
        for (int g = 0; g < S; g++) {
            if (g == g0) continue;
            if (g1 < 0 || key[g] < key[g1]) g1 = g;
        }
        const int n1 = off[g0 + 1] - off[g0], n2 = off[g1 + 1] - off[g1];
        // scipy.stats.ttest_ind (equal_var=True).  Variance as scipy's _var: mean((x-m)^2) * n/(n-1);
        // a single-observation sample contributes variance 0 (_equal_var_ttest_denom, scipy >= 1.9;
        // scipy 1.7.1, pinned by the reference, would propagate NaN there — see DESIGN.md quirks).
        double d[CL_MAX_N];
        const double m1 = npmean[g0], m2 = npmean[g1];
        for (int i = 0; i < n1; i++) {
            const double t = vals[off[g0] + i] - m1;
            d[i] = t * t;
        }
        double v1 = (np_pairwise(d, n1) / (double)n1) * ((double)n1 / (double)(n1 - 1));
        for (int i = 0; i < n2; i++) {
            const double t = vals[off[g1] + i] - m2;
            d[i] = t * t;
        }
        double v2 = (np_pairwise(d, n2) / (double)n2) * ((double)n2 / (double)(n2 - 1));
        // one observation in a group: scipy 1.7.1 (the reference's pin) propagates the NaN variance to the p-value — and
        // the reference KEEPS NaN p-values (Cluster.py:167); scipy >= 1.9 uses variance 0 (SPK_TTEST_SINGLETON=zero)
        if (n1 == 1) v1 = singleton_nan ? NAN : 0.0;
        if (n2 == 1) v2 = singleton_nan ? NAN : 0.0;
        const double df = (double)n1 + (double)n2 - 2.0;
        const double svar = ((double)(n1 - 1) * v1 + (double)(n2 - 1) * v2) / df;
        const double denom = sqrt(svar * (1.0 / (double)n1 + 1.0 / (double)n2));
        const double t = (m1 - m2) / denom;
        double p;
        if (isnan(t) || !(df > 0.0)) p = NAN;
        else if (isinf(t)) p = 0.0;
        else {
            const double t2 = t * t;
            p = betainc_xy(0.5 * df, 0.5, df / (df + t2), t2 / (df + t2));
        }
        best[m] = g0;
        pval[m] = p;
    }
}


// ---- K7b: rank tests for `-test_method kruskal | mannwhitneyu | wilcoxon` (Cluster.py:160,191) -----------------------
// Same row handling as k_ttest_groups (regroup by subgenome, np.mean per group, groups ordered by descending
// Python-sum mean, top two tested); the test is scipy.stats.<method>(top, second) with default arguments:
//   kruskal       H with tie correction, p = chi2.sf(H, 1)                         (scipy/stats/_stats_py.py kruskal)
//   mannwhitneyu  two-sided, continuity correction, method 'auto': exact null distribution when min(n1, n2) <= 8 and
//                 the pooled values have no ties, normal approximation with tie correction otherwise
//   wilcoxon      paired, zero_method 'wilcox', two-sided, no continuity correction, mode 'auto' AS IN THE PINNED
//                 scipy 1.7.1: exact distribution when n <= 25 and no difference is zero (ranks of tied |d| are
//                 averaged and r_plus truncated, as that version does), normal approximation otherwise
// Exact null distributions are built on the host with integer arithmetic (Cluster.py side) and indexed per ordered
// pair of groups: pair_off[g0 * S + g1] = offset into `tables` or -1.
//   mannwhitneyu table: cdf[u] for u = 0 .. n1*n2/2 ;  wilcoxon table: cdf[k], k = 0..K, then sf[k], k = 0..K (K = n(n+1)/2)
// flags: bit 0 = wilcoxon on groups of different size (scipy raises ValueError), bit 1 = kruskal on identical values
// (scipy raises ValueError): the row's p-value is NaN and the host raises like the reference would.
enum { RT_KRUSKAL = 1, RT_MANNWHITNEYU = 2, RT_WILCOXON = 3 };

__device__ __forceinline__ double norm_sf(double z) { return 0.5 * erfc(z * 0.70710678118654752440); }

__global__ void __launch_bounds__(128)
k_ranktest_groups(const double* __restrict__ X, uint64_t M, int n, const int32_t* __restrict__ col_group, int S,
                  int method, const int32_t* __restrict__ pair_off, const double* __restrict__ tables,
                  int32_t* __restrict__ best, double* __restrict__ pval, double* __restrict__ means,
                  uint32_t* __restrict__ flags) {
    double vals[CL_MAX_N];
    double pool[CL_MAX_N];
    int off[CL_MAX_S + 1];
    for (uint64_t m = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; m < M;
         m += (uint64_t)gridDim.x * blockDim.x) {
        const double* x = X + m * n;
        int pos = 0;
        for (int g = 0; g < S; g++) {
            off[g] = pos;
            for (int c = 0; c < n; c++)
                if (col_group[c] == g) vals[pos++] = x[c];
        }
        off[S] = pos;
        double key[CL_MAX_S];
        for (int g = 0; g < S; g++) {
            const int ng = off[g + 1] - off[g];
            means[m * S + g] = np_pairwise(vals + off[g], ng) / (double)ng;
            double s = 0.0;
            for (int i = off[g]; i < off[g + 1]; i++) s += vals[i];
            key[g] = -s / (double)ng;
        }
        int g0 = 0;
        for (int g = 1; g < S; g++)
            if (key[g] < key[g0]) g0 = g;
        int g1 = -1;
        for (int g = 0; g < S; g++) {
            if (g == g0) continue;
            if (g1 < 0 || key[g] < key[g1]) g1 = g;
        }
        const int n1 = off[g0 + 1] - off[g0], n2 = off[g1 + 1] - off[g1];
        const double* a = vals + off[g0];
        const double* b = vals + off[g1];
        double p = NAN;
        if (method == RT_WILCOXON) {
            if (n1 != n2) {
                atomicOr(flags, 1u);
            } else {
                // d = x - y; zeros dropped ('wilcox'); ranks of |d| (average ranks)
                int cnt = 0, nzero = 0;
                for (int i = 0; i < n1; i++) {
                    const double d = a[i] - b[i];
                    if (d == 0.0) nzero++;
                    else pool[cnt++] = d;
                }
                double r_plus = 0.0, r_minus = 0.0, tie = 0.0;
                for (int i = 0; i < cnt; i++) {
                    const double ai = fabs(pool[i]);
                    int less = 0, eq = 0;
                    for (int j = 0; j < cnt; j++) {
                        const double aj = fabs(pool[j]);
                        less += aj < ai;
                        eq += aj == ai;
                    }
                    const double r = (double)less + 0.5 * (double)(eq + 1);
                    if (pool[i] > 0) r_plus += r;
                    else r_minus += r;
                    tie += (double)eq * (double)eq - 1.0;            // sum over ties of t^3 - t, element-wise
                }
                const int tab = pair_off[g0 * S + g1];
                if (n1 <= 25 && nzero == 0 && tab >= 0) {
                    const int K = n1 * (n1 + 1) / 2;
                    const int rp = (int)r_plus;                        // truncation, as scipy 1.7.1
                    if (rp == K / 2) p = 1.0;                          // `r_plus == (len(cnt) - 1) // 2`: centre of the distribution
                    else p = 2.0 * fmin(tables[tab + rp], tables[tab + K + 1 + rp]);
                } else {
                    const double c = (double)cnt;
                    const double mn = c * (c + 1.0) * 0.25;
                    double se = c * (c + 1.0) * (2.0 * c + 1.0);
                    se = sqrt((se - 0.5 * tie) / 24.0);
                    const double T = fmin(r_plus, r_minus);
                    const double z = (T - mn) / se;
                    p = 2.0 * norm_sf(fabs(z));
                }
            }
        } else {
            // pooled ranks
            const int N = n1 + n2;
            for (int i = 0; i < n1; i++) pool[i] = a[i];
            for (int i = 0; i < n2; i++) pool[n1 + i] = b[i];
            double R1 = 0.0, R2 = 0.0, tie = 0.0;
            bool has_ties = false;
            for (int i = 0; i < N; i++) {
                int less = 0, eq = 0;
                for (int j = 0; j < N; j++) {
                    less += pool[j] < pool[i];
                    eq += pool[j] == pool[i];
                }
                const double r = (double)less + 0.5 * (double)(eq + 1);
                if (i < n1) R1 += r;
                else R2 += r;
                tie += (double)eq * (double)eq - 1.0;
                has_ties |= eq > 1;
            }
            const double dN = (double)N;
            if (method == RT_KRUSKAL) {
                const double ties = 1.0 - tie / (dN * dN * dN - dN);
                if (ties == 0.0) {
                    atomicOr(flags, 2u);
                } else {
                    const double ssbn = R1 * R1 / (double)n1 + R2 * R2 / (double)n2;
                    double h = 12.0 / (dN * (dN + 1.0)) * ssbn - 3.0 * (dN + 1.0);
                    h /= ties;
                    p = h <= 0.0 ? 1.0 : erfc(sqrt(0.5 * h));      // chi2.sf(h, df = 1)
                }
            } else {
                const double U1 = R1 - (double)n1 * ((double)n1 + 1.0) * 0.5;
                const double U2 = (double)n1 * (double)n2 - U1;
                const double U = fmax(U1, U2);
                const int tab = pair_off[g0 * S + g1];
                if (!((n1 > 8 && n2 > 8) || has_ties) && tab >= 0) {
                    const int kc = n1 * n2 - (int)U;                   // sf(U) = cdf(n1 n2 - U) by symmetry (U >= n1 n2 / 2)
                    p = fmin(1.0, fmax(0.0, 2.0 * tables[tab + kc]));
                } else {
                    const double mu = (double)n1 * (double)n2 * 0.5;
                    const double sd = sqrt((double)n1 * (double)n2 / 12.0 * ((dN + 1.0) - tie / (dN * (dN - 1.0))));
                    double num = U - mu;
                    if (num > 0.0) num -= 0.5;                        // continuity correction, sign(num) * 0.5
                    const double z = num / sd;
                    p = fmin(1.0, fmax(0.0, 2.0 * norm_sf(z)));
                }
            }
        }
        best[m] = g0;
        pval[m] = p;
    }
}

// ---- K8 PCA from the Gram matrix: cyclic Jacobi eigen-decomposition (one CTA) -----------------------------
__global__ void __launch_bounds__(128)
k_pca_gram(const double* __restrict__ G, int n, int ncomp, double* __restrict__ eigvals,
           double* __restrict__ scores, double* __restrict__ ratio, double* __restrict__ ws) {
    double* A = ws;            // n x n working copy
    double* V = ws + n * n;    // eigenvectors (columns)
    __shared__ double s_c, s_s;
    __shared__ int s_rot;
    __shared__ double s_off;
    const int tid = threadIdx.x;
    for (int e = tid; e < n * n; e += blockDim.x) {
        A[e] = G[e];
        V[e] = (e / n == e % n) ? 1.0 : 0.0;
    }
    __syncthreads();
    for (int sweep = 0; sweep < 60; sweep++) {
        if (tid == 0) {
            double off = 0.0, diag = 0.0;
            for (int i = 0; i < n; i++)
                for (int j = 0; j < n; j++) {
                    if (i == j) diag += A[i * n + i] * A[i * n + i];
                    else off += A[i * n + j] * A[i * n + j];
                }
            s_off = (off <= 1e-30 * diag || off == 0.0) ? 0.0 : off;
        }
        __syncthreads();
        if (s_off == 0.0) break;
        for (int p = 0; p < n - 1; p++) {
            for (int q = p + 1; q < n; q++) {
                if (tid == 0) {
                    const double apq = A[p * n + q];
                    if (fabs(apq) < 1e-300) {
                        s_rot = 0;
                    } else {
                        const double app = A[p * n + p], aqq = A[q * n + q];
                        const double theta = (aqq - app) / (2.0 * apq);
                        const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                        s_c = 1.0 / sqrt(t * t + 1.0);
                        s_s = t * s_c;
                        s_rot = 1;
                    }
                }
                __syncthreads();
                if (s_rot) {
                    const double c = s_c, s = s_s;
                    // columns p, q
                    for (int k = tid; k < n; k += blockDim.x) {
                        const double akp = A[k * n + p], akq = A[k * n + q];
                        A[k * n + p] = c * akp - s * akq;
                        A[k * n + q] = s * akp + c * akq;
                        const double vkp = V[k * n + p], vkq = V[k * n + q];
                        V[k * n + p] = c * vkp - s * vkq;
                        V[k * n + q] = s * vkp + c * vkq;
                    }
                    __syncthreads();
                    // rows p, q
                    for (int k = tid; k < n; k += blockDim.x) {
                        const double apk = A[p * n + k], aqk = A[q * n + k];
                        A[p * n + k] = c * apk - s * aqk;
                        A[q * n + k] = s * apk + c * aqk;
                    }
                }
                __syncthreads();
            }
        }
    }
    if (tid == 0) {
        // selection sort of eigenvalues, descending
        int ord[CL_MAX_N];
        for (int i = 0; i < n; i++) ord[i] = i;
        for (int i = 0; i < n; i++) {
            int b = i;
            for (int j = i + 1; j < n; j++)
                if (A[ord[j] * n + ord[j]] > A[ord[b] * n + ord[b]]) b = j;
            const int t = ord[i];
            ord[i] = ord[b];
            ord[b] = t;
        }
        double total = 0.0;
        for (int i = 0; i < n; i++) total += fmax(A[i * n + i], 0.0);
        for (int cix = 0; cix < n; cix++) eigvals[cix] = A[ord[cix] * n + ord[cix]];
        for (int cix = 0; cix < ncomp; cix++) {
            const double lam = fmax(A[ord[cix] * n + ord[cix]], 0.0);
            ratio[cix] = lam / total;
            const double sl = sqrt(lam);
            // sign: make the entry of largest magnitude positive (sklearn svd_flip, u-based)
            int big = 0;
            for (int i = 1; i < n; i++)
                if (fabs(V[i * n + ord[cix]]) > fabs(V[big * n + ord[cix]])) big = i;
            const double sg = V[big * n + ord[cix]] < 0 ? -1.0 : 1.0;
            for (int i = 0; i < n; i++) scores[i * ncomp + cix] = sg * V[i * n + ord[cix]] * sl;
        }
    }
}

inline unsigned row_grid(uint64_t M, int threads) {
    const uint64_t blocks = (M + threads - 1) / threads;
    const uint64_t cap = (uint64_t)spk_num_sms() * 16;
    return (unsigned)(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
}

}  // namespace

extern "C" int spk_zscore_rows(const double* d_X, uint64_t M, int n, double* d_Z, void* stream) {
    SPK_CHECK_ARG(n >= 1 && n <= CL_MAX_N, "n must be in [1, 128]");
    if (M == 0) return SPK_OK;
    SPK_CHECK_ARG(d_X && d_Z, "null pointer");
    k_zscore_rows<<<row_grid(M, 128), 128, 0, (cudaStream_t)stream>>>(d_X, M, n, d_Z);
    SPK_LAUNCH_CHECK();
    return SPK_OK;
}

extern "C" size_t spk_gram_workspace_bytes(int n) {
    return (size_t)spk_num_sms() * 2 * (size_t)n * n * sizeof(double) + 256;
}

extern "C" int spk_gram(const double* d_Z, uint64_t M, int n, const uint32_t* d_idx, uint64_t n_idx,
                        double* d_G, void* d_ws, size_t ws_bytes, void* stream) {
    SPK_CHECK_ARG(n >= 1 && n <= CL_MAX_N, "n must be in [1, 128]");
    SPK_CHECK_ARG(d_Z && d_G && d_ws, "null pointer");
    if (ws_bytes < spk_gram_workspace_bytes(n)) {
        spk_set_error("spk_gram: workspace too small");
        return SPK_ECAP;
    }
    const uint64_t nrows = d_idx ? n_idx : M;
    cudaStream_t st = (cudaStream_t)stream;
    uint64_t want = (nrows + GR_TILE_ROWS * 8 - 1) / (GR_TILE_ROWS * 8);
    const uint64_t cap = (uint64_t)spk_num_sms() * 2;
    const int nblocks = (int)(want < 1 ? 1 : (want > cap ? cap : want));
    k_gram_partial<<<nblocks, GR_THREADS, 0, st>>>(d_Z, nrows, n, d_idx, (double*)d_ws);
    SPK_LAUNCH_CHECK();
    k_gram_reduce<<<(n * n + 127) / 128, 128, 0, st>>>((const double*)d_ws, nblocks, n * n, d_G);
    SPK_LAUNCH_CHECK();
    return SPK_OK;
}

extern "C" int spk_gram_batched(const double* d_Z, uint64_t M, int n, const uint32_t* d_idx, int R,
                                int B, double* d_G, void* stream) {
    SPK_CHECK_ARG(n >= 1 && n <= CL_MAX_N, "n must be in [1, 128]");
    SPK_CHECK_ARG(R >= 0 && B >= 1, "bad replicate shape");
    if (R == 0) return SPK_OK;
    SPK_CHECK_ARG(d_Z && d_idx && d_G, "null pointer");
    (void)M;
    k_gram_batched<<<R, GR_THREADS, 0, (cudaStream_t)stream>>>(d_Z, n, d_idx, B, d_G);
    SPK_LAUNCH_CHECK();
    return SPK_OK;
}

extern "C" size_t spk_kmeans_workspace_bytes(int R) { return (size_t)R * sizeof(KmScratch) + 256; }

extern "C" int spk_kmeans_gram_at(const double* d_G, int R, int r0, int n, int S, int n_init, int max_iter,
                                  uint64_t seed, const int32_t* d_order, int32_t* d_labels,
                                  double* d_inertia, void* d_ws, size_t ws_bytes, void* stream);

extern "C" int spk_kmeans_gram(const double* d_G, int R, int n, int S, int n_init, int max_iter,
                               uint64_t seed, const int32_t* d_order, int32_t* d_labels,
                               double* d_inertia, void* d_ws, size_t ws_bytes, void* stream) {
    return spk_kmeans_gram_at(d_G, R, 0, n, S, n_init, max_iter, seed, d_order, d_labels, d_inertia, d_ws, ws_bytes,
                              stream);
}

extern "C" int spk_kmeans_gram_at(const double* d_G, int R, int r0, int n, int S, int n_init, int max_iter,
                                  uint64_t seed, const int32_t* d_order, int32_t* d_labels,
                                  double* d_inertia, void* d_ws, size_t ws_bytes, void* stream) {
    SPK_CHECK_ARG(r0 >= 0, "r0 must be >= 0");
    SPK_CHECK_ARG(n >= 1 && n <= CL_MAX_N, "n must be in [1, 128]");
    SPK_CHECK_ARG(S >= 1 && S <= CL_MAX_S && S <= n, "n_clusters must be in [1, min(16, n)]");
    SPK_CHECK_ARG(n_init >= 1 && max_iter >= 1, "n_init/max_iter must be >= 1");
    if (R == 0) return SPK_OK;
    SPK_CHECK_ARG(d_G && d_labels && d_ws, "null pointer");
    if (ws_bytes < spk_kmeans_workspace_bytes(R)) {
        spk_set_error("spk_kmeans_gram: workspace too small");
        return SPK_ECAP;
    }
    k_kmeans_gram<<<(R + 31) / 32, 32, 0, (cudaStream_t)stream>>>(d_G, R, n, S, n_init, max_iter, seed, r0,
                                                                  d_order, d_labels, d_inertia,
                                                                  (KmScratch*)d_ws);
    SPK_LAUNCH_CHECK();
    return SPK_OK;
}

extern "C" int spk_cluster_scores(const int32_t* d_ref_labels, const int32_t* d_labels, int R, int n,
                                  double* d_ari, double* d_vmeasure, void* stream) {
    if (R == 0) return SPK_OK;
    SPK_CHECK_ARG(d_ref_labels && d_labels && d_ari && d_vmeasure, "null pointer");
    SPK_CHECK_ARG(n >= 1 && n <= CL_MAX_N, "n must be in [1, 128]");
    k_cluster_scores<<<(R + 63) / 64, 64, 0, (cudaStream_t)stream>>>(d_ref_labels, d_labels, R, n, d_ari,
                                                                     d_vmeasure);
    SPK_LAUNCH_CHECK();
    return SPK_OK;
}

extern "C" int spk_centroids(const double* d_Z, uint64_t M, int n, const int32_t* d_labels, int S,
                             double* d_C, void* stream) {
    SPK_CHECK_ARG(n >= 1 && n <= CL_MAX_N && S >= 1 && S <= CL_MAX_S, "bad shape");
    if (M == 0) return SPK_OK;
    SPK_CHECK_ARG(d_Z && d_labels && d_C, "null pointer");
    k_centroids<<<row_grid(M, 128), 128, 0, (cudaStream_t)stream>>>(d_Z, M, n, d_labels, S, d_C);
    SPK_LAUNCH_CHECK();
    return SPK_OK;
}

extern "C" int spk_ttest_groups(const double* d_X, uint64_t M, int n, const int32_t* d_col_group, int S,
                                int32_t* d_best, double* d_pval, double* d_means, void* stream) {
    SPK_CHECK_ARG(n >= 2 && n <= CL_MAX_N && S >= 2 && S <= CL_MAX_S, "bad shape");
    if (M == 0) return SPK_OK;
    SPK_CHECK_ARG(d_X && d_col_group && d_best && d_pval && d_means, "null pointer");
    const char* sm = getenv("SPK_TTEST_SINGLETON");
    const int singleton_nan = (sm && sm[0] == 'z') ? 0 : 1;
    k_ttest_groups<<<row_grid(M, 128), 128, 0, (cudaStream_t)stream>>>(d_X, M, n, d_col_group, S, d_best,
                                                                      d_pval, d_means, singleton_nan);
    SPK_LAUNCH_CHECK();
    return SPK_OK;
}

extern "C" size_t spk_pca_workspace_bytes(int n) { return (size_t)2 * n * n * sizeof(double) + 256; }

extern "C" int spk_pca_gram(const double* d_G, int n, int ncomp, double* d_eigvals, double* d_scores,
                            double* d_ratio, void* d_ws, size_t ws_bytes, void* stream) {
    SPK_CHECK_ARG(n >= 1 && n <= CL_MAX_N, "n must be in [1, 128]");
    SPK_CHECK_ARG(ncomp >= 1 && ncomp <= n, "ncomp must be in [1, n]");
    SPK_CHECK_ARG(d_G && d_eigvals && d_scores && d_ratio && d_ws, "null pointer");
    if (ws_bytes < spk_pca_workspace_bytes(n)) {
        spk_set_error("spk_pca_gram: workspace too small");
        return SPK_ECAP;
    }
    k_pca_gram<<<1, 128, 0, (cudaStream_t)stream>>>(d_G, n, ncomp, d_eigvals, d_scores, d_ratio,
                                                    (double*)d_ws);
    SPK_LAUNCH_CHECK();
    return SPK_OK;
}

extern "C" int spk_ranktest_groups(const double* d_X, uint64_t M, int n, const int32_t* d_col_group, int S, int method,
                                   const int32_t* d_pair_off, const double* d_tables, int32_t* d_best, double* d_pval,
                                   double* d_means, uint32_t* d_flags, void* stream) {
    SPK_CHECK_ARG(n >= 2 && n <= CL_MAX_N && S >= 2 && S <= CL_MAX_S, "bad shape");
    SPK_CHECK_ARG(method >= RT_KRUSKAL && method <= RT_WILCOXON, "method: 1 kruskal, 2 mannwhitneyu, 3 wilcoxon");
    if (M == 0) return SPK_OK;
    SPK_CHECK_ARG(d_X && d_col_group && d_pair_off && d_best && d_pval && d_means && d_flags, "null pointer");
    k_ranktest_groups<<<row_grid(M, 128), 128, 0, (cudaStream_t)stream>>>(d_X, M, n, d_col_group, S, method, d_pair_off,
                                                                         d_tables, d_best, d_pval, d_means, d_flags);
    SPK_LAUNCH_CHECK();
    return SPK_OK;
}
