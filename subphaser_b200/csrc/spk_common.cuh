// Shared host/device helpers for libspk (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/spk.h"

void spk_set_error(const char* fmt, ...);
int spk_num_sms();

#define SPK_CHECK_ARG(cond, msg)                                   \
    do {                                                           \
        if (!(cond)) {                                             \
            spk_set_error("%s: invalid argument: %s", __func__, msg); \
            return SPK_EINVAL;                                     \
        }                                                          \
    } while (0)

#define SPK_CUDA(call)                                                                   \
    do {                                                                                 \
        cudaError_t _e = (call);                                                         \
        if (_e != cudaSuccess) {                                                         \
            spk_set_error("%s: CUDA error %s at %s:%d", __func__, cudaGetErrorString(_e), \
                          __FILE__, __LINE__);                                           \
            return SPK_ECUDA;                                                            \
        }                                                                                \
    } while (0)

// every kernel launch of the library passes through here: count it (bench.py reports gpu_launches)
extern unsigned long long g_spk_launches;
#define SPK_LAUNCH_CHECK()            \
    do {                              \
        g_spk_launches++;             \
        SPK_CUDA(cudaGetLastError()); \
    } while (0)

// Tile geometry shared by the sequence-streaming kernels (count, map): one CTA tile is TILE_BASES
// bases = 1 KiB of 2-bit codes + 512 B of validity bits; the halo covers k-1 <= 31 more bases.
constexpr int SPK_TILE_BASES = 4096;
constexpr int SPK_TILE_PACKED_BYTES = SPK_TILE_BASES / 4;         // 1024
constexpr int SPK_TILE_VALID_BYTES = SPK_TILE_BASES / 8;          // 512
constexpr int SPK_HALO_PACKED_BYTES = 16;                          // 64 bases
constexpr int SPK_HALO_VALID_BYTES = 16;                           // 128 bases
constexpr uint64_t SPK_EMPTY_KEY = ~0ull;

// murmur3 fmix64: the avalanche hash used for every open-addressed table.
__host__ __device__ __forceinline__ uint64_t spk_hash64(uint64_t x) {
    x ^= x >> 33;
    x *= 0xff51afd7ed558ccdULL;
    x ^= x >> 33;
    x *= 0xc4ceb9fe1a85ec53ULL;
    x ^= x >> 33;
    return x;
}

#ifdef __CUDACC__
// Map a 64-bit hash onto [0, slots) without a modulo.
__device__ __forceinline__ uint64_t spk_slot_of(uint64_t h, uint64_t slots) {
    return __umul64hi(h, slots);
}

__device__ __forceinline__ uint64_t spk_warp_sum_u64(uint64_t v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ uint32_t spk_warp_sum_u32(uint32_t v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---- mbarrier + 1-D bulk async copy (TMA engine; SASS: UBLKCP) -----------------------------------
__device__ __forceinline__ uint32_t spk_smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void spk_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(spk_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void spk_fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void spk_fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void spk_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(spk_smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void spk_bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes,
                                             uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
            "r"(spk_smem_u32(smem_dst)),
        "l"(gsrc), "r"(bytes), "r"(spk_smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void spk_mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "SPK_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra SPK_DONE_%=;\n"
        "bra SPK_WAIT_%=;\n"
        "SPK_DONE_%=:\n"
        "}\n" ::"r"(spk_smem_u32(bar)),
        "r"(parity)
        : "memory");
}
#endif
