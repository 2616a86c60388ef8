// Sequence tile pipeline shared by the count (K2) and map (K9) kernels.
//
// A CTA of 256 threads walks 4096-base tiles of one chromosome.  The 2-bit codes (1 KiB + 16 B halo)
// and validity bits (512 B + 16 B halo) of a tile are fetched with two 1-D bulk async copies (TMA
// engine, SASS UBLKCP) that complete on an mbarrier; two buffers alternate so the copy of tile t+1
// overlaps the hashing of tile t.  Thread i then owns the 16 k-mers starting at bases 16i..16i+15.
#pragma once
#include "spk_common.cuh"

constexpr int SPK_TILE_THREADS = 256;
constexpr int SPK_KMERS_PER_THREAD = 16;
static_assert(SPK_TILE_THREADS * SPK_KMERS_PER_THREAD == SPK_TILE_BASES, "tile geometry");

struct __align__(16) SpkTileSmem {
    uint32_t packed[2][(SPK_TILE_PACKED_BYTES + SPK_HALO_PACKED_BYTES) / 4];
    uint32_t valid[2][(SPK_TILE_VALID_BYTES + SPK_HALO_VALID_BYTES) / 4];
    uint64_t bar[2];
};

struct SpkKmerParams {
    uint64_t kmask;  // 2k low bits
    uint64_t vmask;  // k low bits
    int k;
    int top_shift;   // 2(k-1)
};

__device__ __forceinline__ SpkKmerParams spk_kmer_params(int k) {
    SpkKmerParams p;
    p.k = k;
    p.kmask = (k == 32) ? ~0ull : ((1ull << (2 * k)) - 1);
    p.vmask = (1ull << k) - 1;
    p.top_shift = 2 * (k - 1);
    return p;
}

__device__ __forceinline__ void spk_tile_init(SpkTileSmem& s) {
    if (threadIdx.x == 0) {
        spk_mbar_init(&s.bar[0], 1);
        spk_mbar_init(&s.bar[1], 1);
        spk_fence_mbar_init();
    }
    __syncthreads();
}

// one thread issues both copies of a tile
__device__ __forceinline__ void spk_tile_issue(SpkTileSmem& s, const uint8_t* __restrict__ packed,
                                               const uint8_t* __restrict__ valid, uint64_t tile,
                                               int buf) {
    constexpr uint32_t PB = SPK_TILE_PACKED_BYTES + SPK_HALO_PACKED_BYTES;
    constexpr uint32_t VB = SPK_TILE_VALID_BYTES + SPK_HALO_VALID_BYTES;
    spk_mbar_expect_tx(&s.bar[buf], PB + VB);
    spk_bulk_g2s(s.packed[buf], packed + tile * SPK_TILE_PACKED_BYTES, PB, &s.bar[buf]);
    spk_bulk_g2s(s.valid[buf], valid + tile * SPK_TILE_VALID_BYTES, VB, &s.bar[buf]);
}

// reverse the order of the 32 2-bit groups of x
__device__ __forceinline__ uint64_t spk_rev2(uint64_t x) {
    x = __brevll(x);
    return ((x & 0xAAAAAAAAAAAAAAAAull) >> 1) | ((x & 0x5555555555555555ull) << 1);
}

// Canonical k-mers of this thread's 16 start positions; bit j of okmask = window j has k valid bases.
// Forward word: first base most significant (integer order == lexicographic A<C<G<T); the reverse
// complement is rolled alongside; canonical = min of the two.
// `pk` / `vd`: shared-memory words of one tile (260 / 132 words incl. halo).
__device__ __forceinline__ void spk_kmers_from(const uint32_t* pk, const uint32_t* vd, const SpkKmerParams& p,
                                               uint64_t (&key)[SPK_KMERS_PER_THREAD],
                                               uint32_t& okmask) {
    const int tid = threadIdx.x;
    const uint32_t w0 = pk[tid], w1 = pk[tid + 1], w2 = pk[tid + 2];
    const uint32_t v0 = vd[tid >> 1], v1 = vd[(tid >> 1) + 1];
    const uint64_t vbits = (((uint64_t)v1 << 32) | v0) >> ((tid & 1) * 16);
    const uint64_t lo = ((uint64_t)w1 << 32) | w0;

    // state before the k-th base is shifted in: the first k-1 bases (base 0 in the low bits of `lo`)
    const uint64_t le = lo & (p.kmask >> 2);
    uint64_t fwd = (p.k > 1) ? (spk_rev2(le) >> (64 - 2 * (p.k - 1))) : 0;
    uint64_t rc = (p.k > 1) ? (((~le) & (p.kmask >> 2)) << 2) : 0;
    // the next 16 bases (indices k-1 .. k+14)
    uint64_t st = lo >> p.top_shift;
    if (p.top_shift > 0) st |= (uint64_t)w2 << (64 - p.top_shift);
    const uint32_t nxt = (uint32_t)st;

    okmask = 0;
#pragma unroll
    for (int j = 0; j < SPK_KMERS_PER_THREAD; j++) {
        const uint64_t b = (nxt >> (2 * j)) & 3u;
        fwd = ((fwd << 2) | b) & p.kmask;
        rc = (rc >> 2) | ((3ull - b) << p.top_shift);
        key[j] = (fwd < rc) ? fwd : rc;
        if (((vbits >> j) & p.vmask) == p.vmask) okmask |= 1u << j;
    }
}

__device__ __forceinline__ void spk_tile_kmers(const SpkTileSmem& s, int buf, const SpkKmerParams& p,
                                               uint64_t (&key)[SPK_KMERS_PER_THREAD],
                                               uint32_t& okmask) {
    spk_kmers_from(s.packed[buf], s.valid[buf], p, key, okmask);
}
