// Bijective mixer on 2k-bit canonical k-mer words, shared by the partitioned counter (K2) and the
// bucketed specific-k-mer table (K9).  f permutes [0, 4^k): the top bits of f(u) pick a partition /
// bucket, the low `rbits` bits are the remainder that is stored; f^-1 recovers the exact k-mer.
#pragma once
#include <stdint.h>

constexpr uint64_t SPK_MIX_C1 = 0xff51afd7ed558ccdULL;

constexpr uint64_t spk_inv64(uint64_t a) {  // multiplicative inverse of odd a modulo 2^64 (Newton)
    uint64_t x = a;
    for (int i = 0; i < 6; i++) x *= 2 - a * x;
    return x;
}
constexpr uint64_t SPK_MIX_C1_INV = spk_inv64(SPK_MIX_C1);
static_assert(SPK_MIX_C1 * SPK_MIX_C1_INV == 1ull, "inverse constant");

struct Mixer {
    uint64_t mask;  // 2k low bits
    int s;          // xorshift distance k = (2k)/2: x ^= x >> s is an involution on 2k-bit words
    int rbits;      // remainder bits = 2k - (partition / bucket bits)
    // fold, one odd multiply, fold: the top bits of the product depend on every bit of the folded word, and the second
    // fold carries them into the low bits the shared-memory tables are indexed with.  (Round 1 used two multiply
    // rounds; on wheat-like and on low-complexity sequence — microsatellites, homopolymers with 2 % substitutions —
    // partition sizes, distinct keys per partition and home-slot collisions are the same with one, and the mixer is a
    // quarter of level 1's instructions.)
    __host__ __device__ uint64_t fwd(uint64_t u) const {
        uint64_t x = u;
        x ^= x >> s;
        x = (x * SPK_MIX_C1) & mask;
        x ^= x >> s;
        return x;
    }
    // one odd multiply: still a bijection on 2k-bit words, and its TOP bits depend on every input bit —
    // enough to spread keys over buckets when an occasional crowded bucket is handled gracefully (K9 table)
    __host__ __device__ uint64_t fwd_light(uint64_t u) const { return (u * SPK_MIX_C1) & mask; }
    __host__ __device__ uint64_t inv(uint64_t x) const {
        x ^= x >> s;
        x = (x * SPK_MIX_C1_INV) & mask;
        x ^= x >> s;
        return x;
    }
};

inline Mixer spk_make_mixer(int k, int top_bits) {
    Mixer m;
    m.mask = (k == 32) ? ~0ull : ((1ull << (2 * k)) - 1);
    m.s = k;
    m.rbits = 2 * k - top_bits;
    return m;
}
