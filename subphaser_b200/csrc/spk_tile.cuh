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
// reverse the order of the 16 2-bit groups of x
__device__ __forceinline__ uint32_t spk_rev2_32(uint32_t x) {
    x = __brev(x);
    return ((x & 0xAAAAAAAAu) >> 1) | ((x & 0x55555555u) << 1);
}

// Canonical k-mers of this thread's 16 start positions; bit j of okmask = window j has k valid bases.
// Forward word: first base most significant (integer order == lexicographic A<C<G<T); canonical = min of
// the forward word and its reverse complement.
// `pk` / `vd`: shared-memory words of one tile (260 / 132 words incl. halo).
//
// No rolling state: with W = the thread's 48-base window as packed (base 0 in the low bits),
//   revcomp_j = (~W >> 2j) & kmask                      (complementing = NOT, and the little-endian packing
//                                                         already is "last base most significant")
//   forward_j = (rev2(W) >> 2(48-k-j)) & kmask          (rev2 = order of the 2-bit groups reversed)
// rev2(W) is shifted once per thread by the k-dependent 2(33-k) bits, so every per-position shift is a
// compile-time constant: 2 funnel shifts + 2 masks per word instead of a 64-bit shift/or/and chain.
// (w0, w1, w2: the three packed words that hold the thread's 48-base window; vbits: its validity bits, bit 0 = base 0)
__device__ __forceinline__ void spk_kmers_from_words(const uint32_t w0, const uint32_t w1, const uint32_t w2,
                                                     const uint64_t vbits, const SpkKmerParams& p,
                                                     uint64_t (&key)[SPK_KMERS_PER_THREAD], uint32_t& okmask) {
    const uint32_t klo = (uint32_t)p.kmask, khi = (uint32_t)(p.kmask >> 32);

    const uint32_t n0 = ~w0, n1 = ~w1, n2 = ~w2;
    uint32_t a0 = spk_rev2_32(w2), a1 = spk_rev2_32(w1), a2 = spk_rev2_32(w0);
    const int s = 2 * (33 - p.k);            // 2 .. 64
    if (s >= 64) { a0 = a2; a1 = 0; a2 = 0; }
    else if (s >= 32) { a0 = a1; a1 = a2; a2 = 0; }
    const uint32_t sl = (uint32_t)s & 31u;
    const uint32_t f0 = __funnelshift_r(a0, a1, sl), f1 = __funnelshift_r(a1, a2, sl), f2 = a2 >> sl;

#pragma unroll
    for (int j = 0; j < SPK_KMERS_PER_THREAD; j++) {
        const uint32_t flo = __funnelshift_r(f0, f1, 2 * (15 - j)) & klo;
        const uint32_t fhi = __funnelshift_r(f1, f2, 2 * (15 - j)) & khi;
        const uint32_t rlo = __funnelshift_r(n0, n1, 2 * j) & klo;
        const uint32_t rhi = __funnelshift_r(n1, n2, 2 * j) & khi;
        const uint64_t fwd = ((uint64_t)fhi << 32) | flo, rc = ((uint64_t)rhi << 32) | rlo;
        key[j] = (fwd < rc) ? fwd : rc;
    }
    // validity: window j needs bits j .. j+k-1 of vbits set.  inv smeared down by k-1 marks the bad starts.
    const uint64_t need = (1ull << (15 + p.k)) - 1;   // 15 + k <= 47
    uint64_t x = ~vbits & need;
    if (x != 0) {
        int covered = 1;
        while (covered < p.k) {
            const int sh = min(covered, p.k - covered);
            x |= x >> sh;
            covered += sh;
        }
    }
    okmask = (uint32_t)(~x) & 0xffffu;
}

__device__ __forceinline__ void spk_kmers_from_t(const uint32_t* pk, const uint32_t* vd, const int tid,
                                                 const SpkKmerParams& p, uint64_t (&key)[SPK_KMERS_PER_THREAD],
                                                 uint32_t& okmask) {
    const uint32_t v0 = vd[tid >> 1], v1 = vd[(tid >> 1) + 1];
    spk_kmers_from_words(pk[tid], pk[tid + 1], pk[tid + 2], (((uint64_t)v1 << 32) | v0) >> ((tid & 1) * 16), p, key,
                         okmask);
}

// canonical k-mer starting at base o (< 16) of the window held in (w0, w1, w2)
__device__ __forceinline__ uint64_t spk_kmer_of_words(const uint32_t w0, const uint32_t w1, const uint32_t w2, int o,
                                                      const SpkKmerParams& p) {
    const int sh = 2 * o;
    const uint64_t lo = ((uint64_t)__funnelshift_r(w1, w2, sh) << 32) | __funnelshift_r(w0, w1, sh);
    const uint64_t le = lo & p.kmask;
    const uint64_t rc = (~lo) & p.kmask;
    const uint64_t fwd = spk_rev2(le) >> (64 - 2 * p.k);
    return fwd < rc ? fwd : rc;
}

__device__ __forceinline__ void spk_kmers_from(const uint32_t* pk, const uint32_t* vd, const SpkKmerParams& p,
                                               uint64_t (&key)[SPK_KMERS_PER_THREAD],
                                               uint32_t& okmask) {
    spk_kmers_from_t(pk, vd, threadIdx.x, p, key, okmask);
}

// Canonical k-mer starting at base `o` of a tile (shared-memory words `pk`), for out-of-line paths that need
// ONE k-mer of the thread's 16 again (cheaper than keeping all 16 addressable in local memory).
// `as_read` (optional): the k-mer as it stands in the sequence is the canonical one (or a palindrome)
__device__ __forceinline__ uint64_t spk_kmer_at(const uint32_t* pk, int o, const SpkKmerParams& p,
                                                bool* as_read = nullptr) {
    const int w = o >> 4, sh = 2 * (o & 15);
    const uint32_t w0 = pk[w], w1 = pk[w + 1], w2 = pk[w + 2];
    const uint64_t lo = ((uint64_t)__funnelshift_r(w1, w2, sh) << 32) | __funnelshift_r(w0, w1, sh);   // 32 bases from o
    const uint64_t le = lo & p.kmask;
    const uint64_t rc = (~lo) & p.kmask;
    const uint64_t fwd = spk_rev2(le) >> (64 - 2 * p.k);
    if (as_read) *as_read = fwd <= rc;
    return fwd < rc ? fwd : rc;
}

__device__ __forceinline__ void spk_tile_kmers(const SpkTileSmem& s, int buf, const SpkKmerParams& p,
                                               uint64_t (&key)[SPK_KMERS_PER_THREAD],
                                               uint32_t& okmask) {
    spk_kmers_from(s.packed[buf], s.valid[buf], p, key, okmask);
}
