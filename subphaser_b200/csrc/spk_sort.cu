// Stable LSD radix sort (4-bit digits) of uint64 keys with a uint32 payload.
// Used for the deterministic (sorted) row order of the differential matrix and for the BH step.
// Sizes here are M <= ~1e7 rows / W <= ~1e6 windows, so simplicity wins over peak sort throughput.
#include "spk_common.cuh"

namespace {

constexpr int SO_THREADS = 256;
constexpr int SO_ITEMS = 8;
constexpr int SO_CHUNK = SO_THREADS * SO_ITEMS;  // 2048 keys per CTA
constexpr int SO_RADIX = 16;

__global__ void __launch_bounds__(SO_THREADS)
k_sort_hist(const uint64_t* __restrict__ keys, uint64_t n, int shift, uint32_t* __restrict__ hist,
            uint32_t nblocks) {
    __shared__ uint32_t s_h[SO_RADIX];
    if (threadIdx.x < SO_RADIX) s_h[threadIdx.x] = 0;
    __syncthreads();
    const uint64_t base = (uint64_t)blockIdx.x * SO_CHUNK;
    for (int q = 0; q < SO_ITEMS; q++) {
        const uint64_t i = base + (uint64_t)q * SO_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&s_h[(keys[i] >> shift) & (SO_RADIX - 1)], 1u);
    }
    __syncthreads();
    if (threadIdx.x < SO_RADIX) hist[(uint64_t)threadIdx.x * nblocks + blockIdx.x] = s_h[threadIdx.x];
}

// exclusive scan over hist laid out digit-major [16][nblocks] (single CTA)
__global__ void __launch_bounds__(1024) k_sort_scan(uint32_t* hist, uint64_t n) {
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (uint64_t base = 0; base < n; base += 1024) {
        const uint64_t i = base + threadIdx.x;
        const uint32_t v = (i < n) ? hist[i] : 0;
        uint32_t incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if ((threadIdx.x & 31) >= o) incl += t;
        }
        if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = incl;
        __syncthreads();
        uint32_t prefix = s_carry;
        for (int w = 0; w < (int)(threadIdx.x >> 5); w++) prefix += s_warp[w];
        if (i < n) hist[i] = prefix + incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = prefix + incl;
        __syncthreads();
    }
}

// Each thread owns SO_ITEMS *consecutive* keys so that (thread, item) order == input order.
__global__ void __launch_bounds__(SO_THREADS)
k_sort_scatter(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ vals, uint64_t n,
               int shift, const uint32_t* __restrict__ hist, uint32_t nblocks,
               uint64_t* __restrict__ out_keys, uint32_t* __restrict__ out_vals) {
    __shared__ uint16_t s_cnt[SO_RADIX][SO_THREADS];
    const uint64_t base = (uint64_t)blockIdx.x * SO_CHUNK + (uint64_t)threadIdx.x * SO_ITEMS;
    uint64_t k[SO_ITEMS];
    uint32_t v[SO_ITEMS];
    uint8_t dg[SO_ITEMS];
    uint32_t cnt_lo = 0, cnt_hi = 0;  // 16 x 4-bit counters (SO_ITEMS <= 15)
#pragma unroll
    for (int q = 0; q < SO_ITEMS; q++) {
        const uint64_t i = base + q;
        if (i < n) {
            k[q] = keys[i];
            v[q] = vals[i];
            dg[q] = (uint8_t)((k[q] >> shift) & (SO_RADIX - 1));
            if (dg[q] < 8) cnt_lo += 1u << (4 * dg[q]);
            else cnt_hi += 1u << (4 * (dg[q] - 8));
        } else {
            k[q] = 0;
            v[q] = 0;
            dg[q] = 0xFF;
        }
    }
#pragma unroll
    for (int d = 0; d < SO_RADIX; d++)
        s_cnt[d][threadIdx.x] = (uint16_t)(((d < 8 ? cnt_lo >> (4 * d) : cnt_hi >> (4 * (d - 8)))) & 15u);
    __syncthreads();
    // exclusive scan of each digit row across the 256 threads: warp w handles digits 2w, 2w+1
    {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        for (int d = warp * 2; d < warp * 2 + 2; d++) {
            uint32_t carry = 0;
            for (int seg = 0; seg < SO_THREADS / 32; seg++) {
                const uint32_t x = s_cnt[d][seg * 32 + lane];
                uint32_t incl = x;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += t;
                }
                s_cnt[d][seg * 32 + lane] = (uint16_t)(carry + incl - x);
                carry += __shfl_sync(0xffffffffu, incl, 31);
            }
        }
    }
    __syncthreads();
    uint32_t seen_lo = 0, seen_hi = 0;
#pragma unroll
    for (int q = 0; q < SO_ITEMS; q++) {
        if (dg[q] != 0xFF) {
            const int d = dg[q];
            const uint32_t within =
                (d < 8 ? (seen_lo >> (4 * d)) : (seen_hi >> (4 * (d - 8)))) & 15u;
            const uint64_t o = (uint64_t)hist[(uint64_t)d * nblocks + blockIdx.x] + s_cnt[d][threadIdx.x] + within;
            out_keys[o] = k[q];
            out_vals[o] = v[q];
            if (d < 8) seen_lo += 1u << (4 * d);
            else seen_hi += 1u << (4 * (d - 8));
        }
    }
}

}  // namespace

extern "C" size_t spk_sort_workspace_bytes(uint64_t n) {
    const uint64_t nblocks = (n + SO_CHUNK - 1) / SO_CHUNK;
    return (size_t)((nblocks > 0 ? nblocks : 1) * SO_RADIX * sizeof(uint32_t) + 256);
}

extern "C" int spk_sort_pairs_u64(uint64_t* d_keys, uint32_t* d_vals, uint64_t* d_keys_tmp,
                                  uint32_t* d_vals_tmp, uint64_t n, int key_bits, void* d_ws,
                                  size_t ws_bytes, void* stream) {
    SPK_CHECK_ARG(key_bits >= 1 && key_bits <= 64, "key_bits must be in [1, 64]");
    SPK_CHECK_ARG(n < 0xffffffffull, "n too large for 32-bit offsets");
    if (n <= 1) return SPK_OK;
    SPK_CHECK_ARG(d_keys && d_vals && d_keys_tmp && d_vals_tmp && d_ws, "null pointer");
    if (ws_bytes < spk_sort_workspace_bytes(n)) {
        spk_set_error("spk_sort_pairs_u64: workspace too small");
        return SPK_ECAP;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const uint32_t nblocks = (uint32_t)((n + SO_CHUNK - 1) / SO_CHUNK);
    uint32_t* hist = (uint32_t*)d_ws;
    int passes = (key_bits + 3) / 4;
    if (passes & 1) passes++;  // even number of passes: the result ends in d_keys/d_vals
    if (passes > 16) passes = 16;
    uint64_t *src_k = d_keys, *dst_k = d_keys_tmp;
    uint32_t *src_v = d_vals, *dst_v = d_vals_tmp;
    for (int p = 0; p < passes; p++) {
        const int shift = 4 * p;
        k_sort_hist<<<nblocks, SO_THREADS, 0, st>>>(src_k, n, shift, hist, nblocks);
        SPK_LAUNCH_CHECK();
        k_sort_scan<<<1, 1024, 0, st>>>(hist, (uint64_t)nblocks * SO_RADIX);
        SPK_LAUNCH_CHECK();
        k_sort_scatter<<<nblocks, SO_THREADS, 0, st>>>(src_k, src_v, n, shift, hist, nblocks, dst_k,
                                                       dst_v);
        SPK_LAUNCH_CHECK();
        uint64_t* tk = src_k; src_k = dst_k; dst_k = tk;
        uint32_t* tv = src_v; src_v = dst_v; dst_v = tv;
    }
    return SPK_OK;
}
