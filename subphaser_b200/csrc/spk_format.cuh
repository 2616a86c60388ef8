// Shortest round-trip formatting of IEEE doubles exactly as Python prints them (`str(float)` / `repr(float)`),
// usable on the device and (for the table-driven unit check in tools/ryu_check.cpp) on the host.
//
// The text wire formats of the path (`.kmer.mat`, `.sig.kmer-subgenome.tsv`; Jellyfish.py:515-520,
// Cluster.py:158-172) hold Python's shortest repr of every float; a 3.6-million-row wheat matrix is 76 million of
// them.  The digits come from the Ryu algorithm (Adams, PLDI 2018: shortest decimal that rounds back to the same
// double, closest to the true value, ties to even) restated here from the paper with 128-bit powers of five
// (spk_ryu_tables.cuh, generated with exact integer arithmetic by tools/gen_ryu_tables.py); the layout rules are
// CPython's float_repr_style 'short': fixed notation for 1e-4 <= |x| < 1e16 with ".0" appended to integers,
// otherwise d.ddde±XX with at least two exponent digits; "inf", "nan", "-0.0".
#pragma once
#include <stdint.h>
#include <string.h>

#ifdef __CUDACC__
#define SPK_FMT_HD __host__ __device__ __forceinline__
#else
#define SPK_FMT_HD inline
#endif

#if defined(__CUDA_ARCH__)
#define SPK_RYU_TABLE_QUAL __device__ const
#else
#define SPK_RYU_TABLE_QUAL static const
#endif
#if defined(__CUDACC__) && !defined(__CUDA_ARCH__)
// host pass of a .cu file: the tables are needed as device symbols too; give the host pass its own copy
#endif
#include "spk_ryu_tables.cuh"

namespace spkfmt {

SPK_FMT_HD uint64_t umulh(uint64_t a, uint64_t b) {
#if defined(__CUDA_ARCH__)
    return __umul64hi(a, b);
#else
    return (uint64_t)(((unsigned __int128)a * b) >> 64);
#endif
}

// (m * mul) >> j for a 128-bit mul = {lo, hi}, 64 < j < 128 + 64
SPK_FMT_HD uint64_t mul_shift64(uint64_t m, const uint64_t* mul, int j) {
    const uint64_t high1 = umulh(m, mul[1]), low1 = m * mul[1];
    const uint64_t high0 = umulh(m, mul[0]);
    const uint64_t sum = high0 + low1;
    const uint64_t hi = high1 + (sum < high0 ? 1u : 0u);
    const int s = j - 64;                           // 0 < s < 64 for every double
    return (sum >> s) | (hi << (64 - s));
}

SPK_FMT_HD int pow5bits(int e) { return (int)(((uint32_t)e * 1217359u) >> 19) + 1; }
SPK_FMT_HD int log10_pow2(int e) { return (int)(((uint32_t)e * 78913u) >> 18); }
SPK_FMT_HD int log10_pow5(int e) { return (int)(((uint32_t)e * 732923u) >> 20); }
SPK_FMT_HD int pow5_factor(uint64_t v) {
    int c = 0;
    while (v > 0 && v % 5 == 0) {
        v /= 5;
        c++;
    }
    return c;
}
SPK_FMT_HD bool multiple_of_pow5(uint64_t v, int p) { return pow5_factor(v) >= p; }
SPK_FMT_HD bool multiple_of_pow2(uint64_t v, int p) { return (v & ((1ull << p) - 1)) == 0; }

// shortest decimal (digits, exponent) of a finite, non-zero double: value = digits * 10^exp
SPK_FMT_HD void d2d(uint64_t ieee_mantissa, uint32_t ieee_exponent, uint64_t& out_digits, int& out_exp) {
    int e2;
    uint64_t m2;
    if (ieee_exponent == 0) {
        e2 = 1 - 1023 - 52 - 2;
        m2 = ieee_mantissa;
    } else {
        e2 = (int)ieee_exponent - 1023 - 52 - 2;
        m2 = (1ull << 52) | ieee_mantissa;
    }
    const bool accept_bounds = (m2 & 1) == 0;
    const uint64_t mv = 4 * m2;
    const uint32_t mm_shift = (ieee_mantissa != 0 || ieee_exponent <= 1) ? 1u : 0u;
    uint64_t vr, vp, vm;
    int e10;
    bool vm_tz = false, vr_tz = false;
    if (e2 >= 0) {
        const int q = log10_pow2(e2) - (e2 > 3);
        e10 = q;
        const int k = SPK_RYU_POW5_INV_BITCOUNT + pow5bits(q) - 1;
        const int i = -e2 + q + k;
        const uint64_t* mul = SPK_RYU_POW5_INV_SPLIT[q];
        vr = mul_shift64(4 * m2, mul, i);
        vp = mul_shift64(4 * m2 + 2, mul, i);
        vm = mul_shift64(4 * m2 - 1 - mm_shift, mul, i);
        if (q <= 21) {
            const uint32_t mv_mod5 = (uint32_t)(mv % 5);
            if (mv_mod5 == 0) vr_tz = multiple_of_pow5(mv, q);
            else if (accept_bounds) vm_tz = multiple_of_pow5(mv - 1 - mm_shift, q);
            else vp -= multiple_of_pow5(mv + 2, q) ? 1u : 0u;
        }
    } else {
        const int q = log10_pow5(-e2) - (-e2 > 1);
        e10 = q + e2;
        const int i = -e2 - q;
        const int k = pow5bits(i) - SPK_RYU_POW5_BITCOUNT;
        const int j = q - k;
        const uint64_t* mul = SPK_RYU_POW5_SPLIT[i];
        vr = mul_shift64(4 * m2, mul, j);
        vp = mul_shift64(4 * m2 + 2, mul, j);
        vm = mul_shift64(4 * m2 - 1 - mm_shift, mul, j);
        if (q <= 1) {
            vr_tz = true;
            if (accept_bounds) vm_tz = mm_shift == 1;
            else --vp;
        } else if (q < 63) {
            vr_tz = multiple_of_pow2(mv, q);
        }
    }
    int removed = 0;
    uint32_t last = 0;
    uint64_t output;
    if (vm_tz || vr_tz) {
        while (vp / 10 > vm / 10) {
            vm_tz &= (vm % 10) == 0;
            vr_tz &= last == 0;
            last = (uint32_t)(vr % 10);
            vr /= 10;
            vp /= 10;
            vm /= 10;
            ++removed;
        }
        if (vm_tz) {
            while (vm % 10 == 0) {
                vr_tz &= last == 0;
                last = (uint32_t)(vr % 10);
                vr /= 10;
                vp /= 10;
                vm /= 10;
                ++removed;
            }
        }
        if (vr_tz && last == 5 && vr % 2 == 0) last = 4;      // exactly half-way: round to even
        output = vr + (((vr == vm && (!accept_bounds || !vm_tz)) || last >= 5) ? 1u : 0u);
    } else {
        bool round_up = false;
        while (vp / 10 > vm / 10) {
            round_up = (vr % 10) >= 5;
            vr /= 10;
            vp /= 10;
            vm /= 10;
            ++removed;
        }
        output = vr + ((vr == vm || round_up) ? 1u : 0u);
    }
    out_digits = output;
    out_exp = e10 + removed;
}

constexpr int PY_REPR_MAX = 25;      // "-1.2345678901234567e-308" is 24 characters

// Python repr(float) of v into out (no terminator); returns the length (<= PY_REPR_MAX)
SPK_FMT_HD int py_repr(double v, char* out) {
    uint64_t bits;
    memcpy(&bits, &v, 8);
    const bool neg = (bits >> 63) != 0;
    const uint64_t man = bits & ((1ull << 52) - 1);
    const uint32_t ex = (uint32_t)((bits >> 52) & 0x7ffu);
    int n = 0;
    if (ex == 0x7ffu) {
        if (man) { out[0] = 'n'; out[1] = 'a'; out[2] = 'n'; return 3; }
        if (neg) out[n++] = '-';
        out[n++] = 'i'; out[n++] = 'n'; out[n++] = 'f';
        return n;
    }
    if (neg) out[n++] = '-';
    if (ex == 0 && man == 0) {
        out[n++] = '0'; out[n++] = '.'; out[n++] = '0';
        return n;
    }
    uint64_t digits;
    int exp10;
    d2d(man, ex, digits, exp10);
    char ds[20];
    int nd = 0;
    while (digits) {                      // least significant first
        ds[nd++] = (char)('0' + (int)(digits % 10));
        digits /= 10;
    }
    const int decpt = nd + exp10;         // value = 0.d1d2... * 10^decpt
    if (decpt <= -4 || decpt > 16) {
        out[n++] = ds[nd - 1];
        if (nd > 1) {
            out[n++] = '.';
            for (int i = nd - 2; i >= 0; i--) out[n++] = ds[i];
        }
        out[n++] = 'e';
        int e = decpt - 1;
        out[n++] = e < 0 ? '-' : '+';
        if (e < 0) e = -e;
        if (e >= 100) {
            out[n++] = (char)('0' + e / 100);
            e %= 100;
            out[n++] = (char)('0' + e / 10);
            out[n++] = (char)('0' + e % 10);
        } else {
            out[n++] = (char)('0' + e / 10);
            out[n++] = (char)('0' + e % 10);
        }
    } else if (decpt <= 0) {
        out[n++] = '0';
        out[n++] = '.';
        for (int i = 0; i < -decpt; i++) out[n++] = '0';
        for (int i = nd - 1; i >= 0; i--) out[n++] = ds[i];
    } else if (decpt >= nd) {
        for (int i = nd - 1; i >= 0; i--) out[n++] = ds[i];
        for (int i = nd; i < decpt; i++) out[n++] = '0';
        out[n++] = '.';
        out[n++] = '0';
    } else {
        for (int i = nd - 1; i >= 0; i--) {
            out[n++] = ds[i];
            if (nd - i == decpt) out[n++] = '.';
        }
    }
    return n;
}

}  // namespace spkfmt
