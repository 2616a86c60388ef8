// K1: FASTA bytes -> 2-bit packed bases + validity bits, entirely on the device.
//
// Replaces the text ingest the reference delegates to `cat f | jellyfish count ... /dev/stdin`
// (Jellyfish.py:697) and to SeqIO.parse + upper() (Seqs.py:121-139).
//
// Three streaming passes over 4-KiB byte tiles (256 threads x 16 B, one coalesced uint4 each):
//   A  last '\n' of every tile                      -> exclusive max-scan  (which line does a tile start in)
//   B  kept bases / valid bases / headers per tile  -> exclusive sum-scan  (where does a tile write)
//   C  classify again and OR the 2-bit codes / validity bits of each thread's <= 16 kept bases straight
//      into the zero-initialised output words (<= 4 RED.OR per thread; no byte stores, no code array)
// A byte is in a header iff the first byte of its line is '>'.
#include <stdlib.h>
#include "spk_common.cuh"

namespace {

constexpr int PK_THREADS = 256;
constexpr int PK_BYTES_PER_THREAD = 16;
constexpr int PK_TILE = PK_THREADS * PK_BYTES_PER_THREAD;  // 4096

struct PackWs {
    int64_t* tile_last_nl;   // [ntiles]  global position of last '\n' in tile (or -1); scanned in place
    uint64_t* tile_off;      // [ntiles+1] kept bases before tile (after scan)
    uint32_t* tile_kept;     // [ntiles]
    uint32_t* tile_skip;     // [ntiles] pass A: number of '\n' + '\r' bytes; bit 31: the tile contains a '>'
    uint64_t* totals;        // [4] scratch: valid count, records
};

__host__ __device__ inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

__host__ PackWs carve_ws(void* ws, size_t nbytes, size_t* total) {
    size_t ntiles = (nbytes + PK_TILE - 1) / PK_TILE;
    if (ntiles == 0) ntiles = 1;
    char* p = (char*)ws;
    size_t off = 0;
    PackWs w;
    w.tile_last_nl = (int64_t*)(p + off);
    off += align_up(ntiles * 8, 256);
    w.tile_off = (uint64_t*)(p + off);
    off += align_up((ntiles + 1) * 8, 256);
    w.tile_kept = (uint32_t*)(p + off);
    off += align_up(ntiles * 4, 256);
    w.tile_skip = (uint32_t*)(p + off);
    off += align_up(ntiles * 4, 256);
    w.totals = (uint64_t*)(p + off);
    off += 256;
    if (total) *total = off;
    return w;
}

__device__ __forceinline__ uint4 load_tile_bytes(const uint8_t* __restrict__ in, size_t nbytes,
                                                 size_t pos) {
    // 16 bytes starting at pos (16-aligned); bytes past the end read as '\n'
    uint4 v;
    if (pos + 16 <= nbytes) {
        v = *reinterpret_cast<const uint4*>(in + pos);
    } else {
        uint8_t tmp[16];
#pragma unroll
        for (int i = 0; i < 16; i++) tmp[i] = (pos + i < nbytes) ? in[pos + i] : (uint8_t)'\n';
        v = *reinterpret_cast<uint4*>(tmp);
    }
    return v;
}

__device__ __forceinline__ uint32_t word_of(const uint4& v, int w) {
    return w == 0 ? v.x : w == 1 ? v.y : w == 2 ? v.z : v.w;
}

// index (0..15) of the last '\n' among the 16 bytes, or -1
__device__ __forceinline__ int last_newline16(const uint4& v) {
    int last = -1;
#pragma unroll
    for (int w = 0; w < 4; w++) {
        const uint32_t x = word_of(v, w) ^ 0x0a0a0a0au;                       // zero byte where '\n'
        const uint32_t z = (x - 0x01010101u) & ~x & 0x80808080u;              // exact for the first zero byte
        if (z) {
            // scan the four bytes explicitly (borrow propagation makes higher flags unreliable)
#pragma unroll
            for (int b = 0; b < 4; b++)
                if (((word_of(v, w) >> (8 * b)) & 0xffu) == 0x0au) last = w * 4 + b;
        }
    }
    return last;
}

// ---- pass A -------------------------------------------------------------------------------------
// 0x80 in every byte of y that is zero (exact: no borrow between bytes)
__device__ __forceinline__ uint32_t zero_bytes(uint32_t y) {
    return ~(((y & 0x7f7f7f7fu) + 0x7f7f7f7fu) | y) & 0x80808080u;
}
__device__ __forceinline__ uint32_t eq_bytes(uint32_t w, uint32_t c) { return zero_bytes(w ^ (c * 0x01010101u)); }

// Per tile: position of the last '\n' (for the line-start carry), and — so that the kept-base count of an
// ordinary tile needs no second look at its bytes — the number of '\n' / '\r' bytes and whether it has a '>'.
__global__ void __launch_bounds__(PK_THREADS) k_tile_last_newline(const uint8_t* __restrict__ in,
                                                                   size_t nbytes, size_t ntiles,
                                                                   int64_t* __restrict__ tile_last_nl,
                                                                   uint32_t* __restrict__ tile_skip,
                                                                   const uint32_t* __restrict__ run_if) {
    __shared__ int s_max[PK_THREADS / 32];
    if (run_if && *run_if == 0) return;
    __shared__ uint32_t s_skip[PK_THREADS / 32];
    for (size_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const size_t pos = tile * PK_TILE + (size_t)threadIdx.x * PK_BYTES_PER_THREAD;
        int last = -1;
        uint32_t skip = 0;       // low bits: count, bit 31: '>' seen
        if (pos < nbytes) {
            const uint4 v = load_tile_bytes(in, nbytes, pos);
            const int l = last_newline16(v);
            if (l >= 0) {
                // bytes past the end were padded with '\n': clamp to real bytes
                int ll = l;
                while (ll >= 0 && pos + ll >= nbytes) ll--;
                while (ll >= 0 && ((word_of(v, ll >> 2) >> (8 * (ll & 3))) & 0xffu) != 0x0au) ll--;
                if (ll >= 0) last = threadIdx.x * 16 + ll;
            }
            uint32_t gt = 0;
#pragma unroll
            for (int w = 0; w < 4; w++) {
                const uint32_t x = word_of(v, w);
                skip += __popc(eq_bytes(x, '\n') | eq_bytes(x, '\r'));
                gt |= eq_bytes(x, '>');
            }
            if (pos + 16 > nbytes) skip -= (uint32_t)(pos + 16 - nbytes);   // the '\n' padding past the end
            if (gt) skip |= 0x80000000u;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) last = max(last, __shfl_xor_sync(0xffffffffu, last, o));
        uint32_t cnt = skip & 0x7fffffffu, flag = skip >> 31;
        cnt = spk_warp_sum_u32(cnt);
        flag = __any_sync(0xffffffffu, flag) ? 1u : 0u;
        if ((threadIdx.x & 31) == 0) {
            s_max[threadIdx.x >> 5] = last;
            s_skip[threadIdx.x >> 5] = cnt | (flag << 31);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int m = -1;
            uint32_t c = 0, f = 0;
            for (int w = 0; w < PK_THREADS / 32; w++) {
                m = max(m, s_max[w]);
                c += s_skip[w] & 0x7fffffffu;
                f |= s_skip[w] >> 31;
            }
            tile_last_nl[tile] = (m < 0) ? -1 : (int64_t)(tile * PK_TILE + m);
            tile_skip[tile] = c | (f << 31);
        }
        __syncthreads();
    }
}

// single-CTA exclusive max-scan (int64) / sum-scan (u32 -> u64) over the per-tile values (bytes/4096 of
// them); 8 consecutive elements per thread per round keep the number of latency-bound rounds small
constexpr int SCAN_PER = 8;

__global__ void __launch_bounds__(1024) k_scan_max_excl(int64_t* data, size_t n, const uint32_t* run_if) {
    __shared__ int64_t s_warp[32];
    __shared__ int64_t s_carry;
    if (run_if && *run_if == 0) return;
    if (threadIdx.x == 0) s_carry = -1;
    __syncthreads();
    for (size_t base = 0; base < n; base += 1024 * SCAN_PER) {
        const size_t i0 = base + (size_t)threadIdx.x * SCAN_PER;
        int64_t loc[SCAN_PER];
        int64_t mine = -1;
#pragma unroll
        for (int q = 0; q < SCAN_PER; q++) {
            loc[q] = (i0 + q < n) ? data[i0 + q] : -1;
            mine = max(mine, loc[q]);
        }
        int64_t incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int64_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if ((threadIdx.x & 31) >= o) incl = max(incl, t);
        }
        if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = incl;
        __syncthreads();
        int64_t prefix = s_carry;
        for (int w = 0; w < (int)(threadIdx.x >> 5); w++) prefix = max(prefix, s_warp[w]);
        int64_t excl = __shfl_up_sync(0xffffffffu, incl, 1);
        if ((threadIdx.x & 31) == 0) excl = -1;
        int64_t run = max(excl, prefix);
#pragma unroll
        for (int q = 0; q < SCAN_PER; q++) {
            if (i0 + q < n) data[i0 + q] = run;
            run = max(run, loc[q]);
        }
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = max(prefix, incl);
        __syncthreads();
    }
}

__global__ void __launch_bounds__(1024) k_scan_sum_excl(const uint32_t* in, uint64_t* out, size_t n,
                                                         const uint32_t* run_if) {
    __shared__ uint64_t s_warp[32];
    __shared__ uint64_t s_carry;
    if (run_if && *run_if == 0) return;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (size_t base = 0; base < n; base += 1024 * SCAN_PER) {
        const size_t i0 = base + (size_t)threadIdx.x * SCAN_PER;
        uint32_t loc[SCAN_PER];
        uint64_t mine = 0;
#pragma unroll
        for (int q = 0; q < SCAN_PER; q++) {
            loc[q] = (i0 + q < n) ? in[i0 + q] : 0u;
            mine += loc[q];
        }
        uint64_t incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint64_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if ((threadIdx.x & 31) >= o) incl += t;
        }
        if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = incl;
        __syncthreads();
        uint64_t prefix = s_carry;
        for (int w = 0; w < (int)(threadIdx.x >> 5); w++) prefix += s_warp[w];
        uint64_t run = prefix + incl - mine;
#pragma unroll
        for (int q = 0; q < SCAN_PER; q++) {
            if (i0 + q < n) out[i0 + q] = run;
            run += loc[q];
        }
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = prefix + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) out[n] = s_carry;
}

// Classify the 16 bytes of one thread.  Kept bases are appended, in order, to `bits` (2 bits each, first
// kept base lowest) and `vbits` (1 bit each).  `bol`: the first byte is at the beginning of a line;
// `hdr`: the line the first byte belongs to is a header (only meaningful when !bol).
__device__ __forceinline__ void classify16(const uint4& v, size_t pos, size_t nbytes, bool bol, bool hdr,
                                           uint32_t& bits, uint32_t& vbits, int& kept, int& valid,
                                           int& headers) {
    bits = vbits = 0;
    kept = valid = headers = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) {
        const uint32_t c = (word_of(v, i >> 2) >> (8 * (i & 3))) & 0xffu;
        if (pos + i >= nbytes) break;
        if (c == '\n') {
            bol = true;
            hdr = false;
            continue;
        }
        if (bol && c == '>') {
            hdr = true;
            headers++;
            bol = false;
            if (pos + i != 0) kept++;  // record separator: one invalid base (bits stay 0)
            continue;
        }
        bol = false;
        if (hdr || c == '\r') continue;
        const uint32_t u = (c & 0xdfu) - 'A';                                // upper-case letter index
        const bool ok = u < 26u && ((0x00080045u >> u) & 1u);                // A, C, G, T
        if (ok) {
            const uint32_t x = (c >> 1) & 3u;                                // A0 C1 G3 T2
            bits |= (x ^ (x >> 1)) << (2 * kept);                            // A0 C1 G2 T3
            vbits |= 1u << kept;
            valid++;
        }
        kept++;
    }
}

// ---- SWAR fast path --------------------------------------------------------------------------------------
// Same result as classify16 for 16 in-range bytes that are not inside a header line and contain no '>':
// four bytes per 32-bit operation instead of a 16-step byte loop (the byte loop made K1 ALU-bound).
// Returns false when the bytes need the general path.
__device__ __forceinline__ bool classify16_fast(const uint4& v, uint32_t& bits, uint32_t& vbits, int& kept,
                                                int& valid) {
    uint32_t skm = 0, gt = 0;
    bits = vbits = 0;
#pragma unroll
    for (int wi = 0; wi < 4; wi++) {
        const uint32_t w = word_of(v, wi);
        gt |= eq_bytes(w, '>');
        const uint32_t sk = eq_bytes(w, '\n') | eq_bytes(w, '\r');
        const uint32_t u = w & 0xdfdfdfdfu;                                       // upper-case
        const uint32_t va = eq_bytes(u, 'A') | eq_bytes(u, 'C') | eq_bytes(u, 'G') | eq_bytes(u, 'T');
        const uint32_t vb = va >> 7;                                              // 0/1 per byte
        const uint32_t x = (w >> 1) & 0x03030303u;                                // A0 C1 G3 T2
        const uint32_t code = (x ^ ((x >> 1) & 0x01010101u)) & (vb | (vb << 1));  // A0 C1 G2 T3, 0 if invalid
        bits |= ((code * 0x01041040u) >> 24) << (8 * wi);                         // 4 x 2 bits
        vbits |= (((vb * 0x01020408u) >> 24) & 0xfu) << (4 * wi);                 // 4 x 1 bit
        skm |= ((((sk >> 7) * 0x01020408u) >> 24) & 0xfu) << (4 * wi);
    }
    if (gt) return false;
    valid = __popc(vbits);
    kept = 16 - __popc(skm);
    while (skm) {                                   // drop the skipped bytes ('\n', '\r'), highest first
        const int q = 31 - __clz(skm);
        skm &= ~(1u << q);
        const uint32_t lo2 = bits & ((1u << (2 * q)) - 1u), lo1 = vbits & ((1u << q) - 1u);
        const uint32_t hi2 = (q == 15) ? 0u : (bits >> (2 * q + 2)), hi1 = vbits >> (q + 1);
        bits = lo2 | (hi2 << (2 * q));
        vbits = lo1 | (hi1 << q);
    }
    return true;
}

// exclusive max-scan of per-thread last-newline offsets (int, tile-local; -1 none) inside a CTA
__device__ __forceinline__ int block_excl_max(int v, int* s_warp) {
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if ((threadIdx.x & 31) >= o) incl = max(incl, t);
    }
    if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = incl;
    __syncthreads();
    int prefix = -1;
    for (int w = 0; w < (int)(threadIdx.x >> 5); w++) prefix = max(prefix, s_warp[w]);
    int excl = __shfl_up_sync(0xffffffffu, incl, 1);
    if ((threadIdx.x & 31) == 0) excl = -1;
    __syncthreads();
    return max(excl, prefix);
}

__device__ __forceinline__ uint32_t block_excl_sum(uint32_t v, uint32_t* s_warp, uint32_t* total) {
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if ((threadIdx.x & 31) >= o) incl += t;
    }
    if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = incl;
    __syncthreads();
    uint32_t prefix = 0, tot = 0;
    for (int w = 0; w < PK_THREADS / 32; w++) {
        if (w < (int)(threadIdx.x >> 5)) prefix += s_warp[w];
        tot += s_warp[w];
    }
    __syncthreads();
    if (total) *total = tot;
    return prefix + incl - v;
}

// ---- passes B and C share one body ---------------------------------------------------------------
template <bool EMIT>
__global__ void __launch_bounds__(PK_THREADS) k_classify(const uint8_t* __restrict__ in, size_t nbytes,
                                                          size_t ntiles,
                                                          const int64_t* __restrict__ tile_carry_nl,
                                                          const uint32_t* __restrict__ tile_skip,
                                                          uint32_t* __restrict__ tile_kept,
                                                          const uint64_t* __restrict__ tile_off,
                                                          uint32_t* __restrict__ packed,
                                                          uint32_t* __restrict__ valid_out,
                                                          uint64_t* __restrict__ totals,
                                                          const uint32_t* __restrict__ run_if) {
    __shared__ int s_wi[PK_THREADS / 32];
    __shared__ uint32_t s_w32[PK_THREADS / 32];
    __shared__ int s_carry_hdr;
    __shared__ uint32_t s_pk[EMIT ? PK_TILE / 16 + 4 : 1];
    if (run_if && *run_if == 0) return;
    __shared__ uint32_t s_vd[EMIT ? PK_TILE / 32 + 2 : 1];
    uint64_t my_valid = 0, my_hdr = 0;
    for (size_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const size_t tbase = tile * PK_TILE;
        const size_t pos = tbase + (size_t)threadIdx.x * PK_BYTES_PER_THREAD;
        const int64_t carry = tile_carry_nl[tile];       // last '\n' before this tile (-1: none)
        if (threadIdx.x == 0) {
            // is the line that straddles into this tile a header?  (one byte read per tile)
            const size_t ls = (size_t)(carry + 1);
            s_carry_hdr = (ls < tbase) ? (in[ls] == '>') : 0;
        }
        if (!EMIT) {
            // an ordinary tile (no '>' in it, not entered inside a header line) keeps every byte that is not a
            // line break: pass A already counted those, so the tile's bytes are not read a second time here
            __syncthreads();
            const uint32_t sk = tile_skip[tile];
            if (!(sk >> 31) && !s_carry_hdr) {
                if (threadIdx.x == 0) {
                    const size_t nb = (nbytes - tbase < (size_t)PK_TILE) ? nbytes - tbase : (size_t)PK_TILE;
                    tile_kept[tile] = (uint32_t)nb - (sk & 0x7fffffffu);
                }
                __syncthreads();   // s_carry_hdr is rewritten by the next iteration
                continue;
            }
        }
        uint4 v = make_uint4(0x0a0a0a0au, 0x0a0a0a0au, 0x0a0a0a0au, 0x0a0a0a0au);
        int my_last = -1;
        if (pos < nbytes) {
            v = load_tile_bytes(in, nbytes, pos);
            const int l = last_newline16(v);
            if (l >= 0) my_last = threadIdx.x * 16 + l;   // padded '\n' past the end are harmless here
        }
        const int prev_nl = block_excl_max(my_last, s_wi);   // tile-local offset of the last '\n' before my bytes
        // state at my first byte
        bool bol = false, hdr = false;
        const int my_off = threadIdx.x * 16;
        if (pos >= nbytes) {
            // nothing to classify (keeps every global read inside the buffer)
        } else if (prev_nl >= 0) {
            bol = (prev_nl + 1 == my_off);
            // first byte of my line lives in this tile at offset prev_nl+1 (< my_off when !bol)
            hdr = bol ? false : (in[tbase + prev_nl + 1] == '>');
        } else {
            bol = ((int64_t)tbase + my_off == carry + 1);
            hdr = bol ? false : (((size_t)(carry + 1) >= tbase) ? (in[carry + 1] == '>') : (s_carry_hdr != 0));
        }
        uint32_t bits, vbits;
        int kept, nvalid, headers = 0;
        if (!(pos + 16 <= nbytes && !hdr && classify16_fast(v, bits, vbits, kept, nvalid)))
            classify16(v, pos, nbytes, bol, hdr, bits, vbits, kept, nvalid, headers);
        if (!EMIT) {
            uint32_t tot;
            block_excl_sum((uint32_t)kept, s_w32, &tot);
            if (threadIdx.x == 0) tile_kept[tile] = tot;
            my_hdr += headers;
        } else {
            my_valid += nvalid;
            // Stage the tile's output words in shared memory (OR of <= 4 words per thread), then write whole
            // words with coalesced plain stores; only the two words a tile may share with its neighbours
            // go through a global atomicOr (the outputs are zero-initialised).
            const uint32_t off = block_excl_sum((uint32_t)kept, s_w32, nullptr);
            const uint64_t o0 = tile_off[tile];
            const uint32_t T = (uint32_t)(tile_off[tile + 1] - o0);      // kept bases of this tile
            const uint32_t lb = (uint32_t)(o0 & 31);                     // local bit origin: valid word o0/32
            for (int i = threadIdx.x; i < PK_TILE / 16 + 4; i += PK_THREADS) s_pk[i] = 0;
            for (int i = threadIdx.x; i < PK_TILE / 32 + 2; i += PK_THREADS) s_vd[i] = 0;
            __syncthreads();
            if (kept) {
                const uint32_t l = lb + off;
                const int sh = 2 * (int)(l & 15);
                if (bits) {
                    if (bits << sh) atomicOr(&s_pk[l >> 4], bits << sh);
                    if (sh && (bits >> (32 - sh))) atomicOr(&s_pk[(l >> 4) + 1], bits >> (32 - sh));
                }
                if (vbits) {
                    const int vs = (int)(l & 31);
                    atomicOr(&s_vd[l >> 5], vbits << vs);
                    if (vs > 16 && (vbits >> (32 - vs))) atomicOr(&s_vd[(l >> 5) + 1], vbits >> (32 - vs));
                }
            }
            __syncthreads();
            if (T) {
                const uint32_t npw = ((lb + T - 1) >> 4) + 1;            // packed words touched (from word 2*(o0/32))
                const uint32_t nvw = ((lb + T - 1) >> 5) + 1;
                uint32_t* gp = packed + (o0 >> 5) * 2;
                uint32_t* gv = valid_out + (o0 >> 5);
                for (uint32_t i = threadIdx.x; i < npw; i += PK_THREADS) {
                    const uint32_t wv = s_pk[i];
                    if (i < 2 || i + 1 >= npw) { if (wv) atomicOr(&gp[i], wv); }
                    else gp[i] = wv;
                }
                for (uint32_t i = threadIdx.x; i < nvw; i += PK_THREADS) {
                    const uint32_t wv = s_vd[i];
                    if (i == 0 || i + 1 == nvw) { if (wv) atomicOr(&gv[i], wv); }
                    else gv[i] = wv;
                }
            }
        }
        __syncthreads();   // s_carry_hdr is rewritten by the next iteration
    }
    my_valid = spk_warp_sum_u64(my_valid);     // valid bases: counted where they are emitted (pass C)
    my_hdr = spk_warp_sum_u64(my_hdr);         // records: only tiles with a '>' reach the classifier in pass B
    if ((threadIdx.x & 31) == 0) {
        if (my_valid) atomicAdd((unsigned long long*)&totals[0], (unsigned long long)my_valid);
        if (my_hdr) atomicAdd((unsigned long long*)&totals[1], (unsigned long long)my_hdr);
    }
}


// ---- single pass (default) -----------------------------------------------------------------------------------------
// The three-pass scheme above reads the FASTA twice and scans two per-tile arrays with a single CTA.  For ordinary
// (line-wrapped) FASTA all of that can be decided locally: the state a tile starts in ("is the line that contains
// the tile's first byte a header line?") follows from the last '\n' within the 512 bytes before the tile, and the
// only global quantity left — how many kept bases precede the tile — is a prefix sum that the tiles resolve among
// themselves with decoupled look-back (tiles are handed out in order by a ticket; each publishes its kept count,
// then the inclusive prefix, in one 64-bit status word).  One read of the input, no scan kernels.
// A tile that finds no line break in its look-back window (lines longer than 512 bytes: unwrapped FASTA) raises
// *need_slow; the three-pass kernels then run (they return immediately otherwise) — decided on the device, no
// host synchronisation.
constexpr int PK_LOOKBACK = 512;
constexpr uint64_t PK_FLAG_AGG = 1ull << 62, PK_FLAG_INC = 2ull << 62, PK_VAL_MASK = (1ull << 62) - 1;

__device__ __forceinline__ uint64_t ld_status(const uint64_t* p) {
    uint64_t v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_status(uint64_t* p, uint64_t v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// exclusive prefix of tile `tile` (sum of the kept counts of all earlier tiles); called by warp 0
__device__ __forceinline__ uint64_t pk_lookback(const uint64_t* status, int64_t tile) {
    const int lane = threadIdx.x & 31;
    uint64_t excl = 0;
    int64_t j = tile - 1;
    while (true) {
        const int64_t idx = j - lane;
        uint64_t sv = PK_FLAG_INC;                       // before the first tile: inclusive prefix 0
        if (idx >= 0) {
            sv = ld_status(status + idx);
            while ((sv >> 62) == 0) sv = ld_status(status + idx);
        }
        const uint32_t inc = __ballot_sync(0xffffffffu, (sv >> 62) == 2);
        uint64_t v = sv & PK_VAL_MASK;
        if (inc) {
            const int first = __ffs(inc) - 1;            // nearest tile with a complete prefix
            if (lane > first) v = 0;
            excl += spk_warp_sum_u64(v);
            break;
        }
        excl += spk_warp_sum_u64(v);
        j -= 32;
    }
    return excl;
}

constexpr int PK_SUB = 4;                       // 4-KiB sub-tiles per look-back unit (one ticket, one status word per 16 KiB)

__global__ void __launch_bounds__(PK_THREADS)
k_pack_single(const uint8_t* __restrict__ in, size_t nbytes, size_t nunits, uint64_t* __restrict__ status,
              uint32_t* __restrict__ ticket, uint32_t* __restrict__ packed, uint32_t* __restrict__ valid_out,
              uint64_t* __restrict__ totals, uint32_t* __restrict__ need_slow, const uint32_t* __restrict__ run_if) {
    if (run_if && *run_if == 0) return;              // (the regular-layout kernel took the call)
    __shared__ int s_wi[PK_THREADS / 32];
    __shared__ uint32_t s_w32[PK_THREADS / 32];
    __shared__ int s_carry_hdr;
    __shared__ long long s_carry;
    __shared__ uint32_t s_unit;
    __shared__ unsigned long long s_o0;
    __shared__ uint32_t s_T[PK_SUB];
    __shared__ uint32_t s_pk[PK_TILE / 16 + 4];
    __shared__ uint32_t s_vd[PK_TILE / 32 + 2];
    __shared__ __align__(16) uint8_t s_bytes[PK_TILE];
    uint64_t my_valid = 0, my_hdr = 0;
    const int lane = threadIdx.x & 31;
    while (true) {
        if (threadIdx.x == 0) s_unit = atomicAdd(ticket, 1u);
        __syncthreads();
        const size_t unit = s_unit;
        if (unit >= nunits) break;
        uint32_t bits_s[PK_SUB], vbits_s[PK_SUB], off_s[PK_SUB];
        // ---- phase 1: classify the unit's sub-tiles (state at a sub-tile's first byte from its look-back window) ----
#pragma unroll
        for (int sub = 0; sub < PK_SUB; sub++) {
            const size_t tbase = (unit * PK_SUB + sub) * PK_TILE;
            const size_t pos = tbase + (size_t)threadIdx.x * PK_BYTES_PER_THREAD;
            uint4 v = make_uint4(0x0a0a0a0au, 0x0a0a0a0au, 0x0a0a0a0au, 0x0a0a0a0au);
            int my_last = -1;
            if (pos < nbytes) {
                v = load_tile_bytes(in, nbytes, pos);
                const int l = last_newline16(v);
                if (l >= 0) my_last = threadIdx.x * 16 + l;   // padded '\n' past the end are harmless here
            }
            *reinterpret_cast<uint4*>(s_bytes + threadIdx.x * 16) = v;
            if (threadIdx.x < 32) {
                long long found = -1;
                int hdr = 0;
                if (tbase < nbytes) {
                    const size_t w0 = tbase >= (size_t)PK_LOOKBACK ? tbase - PK_LOOKBACK : 0;
                    const size_t p = w0 + (size_t)lane * 16;       // w0 is a multiple of 16
                    if (p < tbase) {
                        const uint4 b = *reinterpret_cast<const uint4*>(in + p);
                        const int l = last_newline16(b);
                        if (l >= 0) found = (long long)(p + l);
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) found = max(found, __shfl_xor_sync(0xffffffffu, found, o));
                    if (lane == 0) {
                        if (found < 0 && w0 > 0) {
                            atomicOr(need_slow, 1u);           // a line longer than the window: the three-pass path decides
                        } else {
                            const size_t ls = (size_t)(found + 1);
                            hdr = (ls < tbase) ? (in[ls] == '>') : 0;
                        }
                    }
                }
                if (lane == 0) {
                    s_carry = found;
                    s_carry_hdr = hdr;
                }
            }
            const int prev_nl = block_excl_max(my_last, s_wi);   // (barriers inside: s_bytes, s_carry visible)
            const long long carry = s_carry;
            bool bol = false, hdr = false;
            const int my_off = threadIdx.x * 16;
            if (pos >= nbytes) {
            } else if (prev_nl >= 0) {
                bol = (prev_nl + 1 == my_off);
                hdr = bol ? false : (s_bytes[prev_nl + 1] == '>');
            } else {
                bol = ((long long)tbase + my_off == carry + 1);
                hdr = bol ? false : (((size_t)(carry + 1) >= tbase) ? (s_bytes[(size_t)(carry + 1) - tbase] == '>')
                                                                    : (s_carry_hdr != 0));
            }
            uint32_t bits, vbits;
            int kept, nvalid, headers = 0;
            if (!(pos + 16 <= nbytes && !hdr && classify16_fast(v, bits, vbits, kept, nvalid)))
                classify16(v, pos, nbytes, bol, hdr, bits, vbits, kept, nvalid, headers);
            my_valid += nvalid;
            my_hdr += headers;
            uint32_t T;
            off_s[sub] = block_excl_sum((uint32_t)kept, s_w32, &T) | ((uint32_t)kept << 16);   // offset < 4096, kept <= 16
            bits_s[sub] = bits;
            vbits_s[sub] = vbits;
            if (threadIdx.x == 0) s_T[sub] = T;
            __syncthreads();       // s_bytes / s_carry are rewritten by the next sub-tile
        }
        // ---- phase 2: publish the kept count, resolve the prefix by look-back, publish the inclusive prefix ----
        if (threadIdx.x < 32) {
            uint32_t T = 0;
#pragma unroll
            for (int sub = 0; sub < PK_SUB; sub++) T += s_T[sub];
            if (lane == 0) st_status(status + unit, PK_FLAG_AGG | (uint64_t)T);
            const uint64_t excl = pk_lookback(status, (int64_t)unit);
            if (lane == 0) {
                st_status(status + unit, PK_FLAG_INC | (excl + T));
                s_o0 = excl;
                if (unit + 1 == nunits) totals[2] = excl + T;     // number of bases
            }
        }
        __syncthreads();
        // ---- phase 3: emit ----
        uint64_t o0 = s_o0;
#pragma unroll
        for (int sub = 0; sub < PK_SUB; sub++) {
            const uint32_t T = s_T[sub];
            const uint32_t off = off_s[sub] & 0xffffu, kept = off_s[sub] >> 16;
            const uint32_t bits = bits_s[sub], vbits = vbits_s[sub];
            for (int i = threadIdx.x; i < PK_TILE / 16 + 4; i += PK_THREADS) s_pk[i] = 0;
            for (int i = threadIdx.x; i < PK_TILE / 32 + 2; i += PK_THREADS) s_vd[i] = 0;
            __syncthreads();
            const uint32_t lb = (uint32_t)(o0 & 31);
            if (kept) {
                const uint32_t l = lb + off;
                const int sh = 2 * (int)(l & 15);
                if (bits) {
                    if (bits << sh) atomicOr(&s_pk[l >> 4], bits << sh);
                    if (sh && (bits >> (32 - sh))) atomicOr(&s_pk[(l >> 4) + 1], bits >> (32 - sh));
                }
                if (vbits) {
                    const int vs = (int)(l & 31);
                    atomicOr(&s_vd[l >> 5], vbits << vs);
                    if (vs > 16 && (vbits >> (32 - vs))) atomicOr(&s_vd[(l >> 5) + 1], vbits >> (32 - vs));
                }
            }
            __syncthreads();
            if (T) {
                const uint32_t npw = ((lb + T - 1) >> 4) + 1;
                const uint32_t nvw = ((lb + T - 1) >> 5) + 1;
                uint32_t* gp = packed + (o0 >> 5) * 2;
                uint32_t* gv = valid_out + (o0 >> 5);
                for (uint32_t i = threadIdx.x; i < npw; i += PK_THREADS) {
                    const uint32_t wv = s_pk[i];
                    if (i < 2 || i + 1 >= npw) { if (wv) atomicOr(&gp[i], wv); }
                    else gp[i] = wv;
                }
                for (uint32_t i = threadIdx.x; i < nvw; i += PK_THREADS) {
                    const uint32_t wv = s_vd[i];
                    if (i == 0 || i + 1 == nvw) { if (wv) atomicOr(&gv[i], wv); }
                    else gv[i] = wv;
                }
            }
            o0 += T;
            __syncthreads();   // staging words are cleared by the next sub-tile
        }
    }
    my_valid = spk_warp_sum_u64(my_valid);
    my_hdr = spk_warp_sum_u64(my_hdr);
    if (lane == 0) {
        if (my_valid) atomicAdd((unsigned long long*)&totals[0], (unsigned long long)my_valid);
        if (my_hdr) atomicAdd((unsigned long long*)&totals[1], (unsigned long long)my_hdr);
    }
}

// ---- regular layout (default first attempt) ---------------------------------------------------------------------------
// One record whose sequence lines all hold exactly W bases and end in '\n' (the last line may be shorter) — what
// BioPython, samtools faidx and the reference's own split_genomes write.  Base j then sits at byte h + j + j / W
// (h = length of the header line), so nothing has to be scanned: every thread turns the <= 35 bytes that hold its 32
// bases into one validity word and two packed words, written once, coalesced, no atomics, no zero-fill beforehand.
// The layout is not trusted: every thread checks that the line breaks of its byte span are exactly the predicted
// ones and that the span holds no '\r' and no '>'; any deviation raises *irregular and the general single-pass
// kernel (and behind it the three-pass one) redoes the call — decided on the device.
constexpr int PKR_THREADS = 256;
constexpr int PKR_BASES = PKR_THREADS * 32;                    // 8192 bases per CTA
constexpr int PKR_MIN_W = 16;
constexpr int PKR_STAGE = PKR_BASES + PKR_BASES / PKR_MIN_W + 64 + 16;   // bytes staged per CTA (<= 8784)
constexpr int PKR_PROBE = 1 << 16;                             // the header line and the first sequence line end within 64 KiB

struct PackReg {               // written by k_pack_probe
    unsigned long long h, W, n;
};

// first '\n' at or after `from` within [from, lim): 4-KiB rounds of one 16-byte load per thread, stopping at the first
// round with a hit (the header and the first sequence line normally end within the first round)
__device__ unsigned long long pkr_find_newline(const uint8_t* __restrict__ in, size_t nbytes, size_t from, size_t lim,
                                               unsigned long long* s_min) {
    for (size_t base = from & ~(size_t)15; base < lim; base += 256 * 16) {
        const size_t pos = base + (size_t)threadIdx.x * 16;
        unsigned long long f = ~0ull;
        if (pos < lim) {
            const uint4 v = load_tile_bytes(in, nbytes, pos);      // (bytes past the end read as '\n': bounded by lim below)
#pragma unroll
            for (int i = 15; i >= 0; i--) {
                const uint32_t c = (word_of(v, i >> 2) >> (8 * (i & 3))) & 0xffu;
                if (c == '\n' && pos + i >= from && pos + i < lim) f = pos + i;
            }
        }
        if (f != ~0ull) atomicMin(s_min, f);
        __syncthreads();
        const unsigned long long cur = *s_min;
        __syncthreads();                     // (everyone has read it before the next round may lower it)
        if (cur != ~0ull) break;
    }
    __syncthreads();
    return *s_min;
}

__global__ void __launch_bounds__(256) k_pack_probe(const uint8_t* __restrict__ in, size_t nbytes, PackReg* reg,
                                                    uint32_t* irregular, uint64_t* totals) {
    __shared__ unsigned long long s_first, s_second;
    if (threadIdx.x == 0) {
        s_first = ~0ull;
        s_second = ~0ull;
    }
    __syncthreads();
    const unsigned long long first = pkr_find_newline(in, nbytes, 0, min(nbytes, (size_t)PKR_PROBE), &s_first);
    const unsigned long long h = first == ~0ull ? ~0ull : first + 1;
    const size_t lim2 = h == ~0ull ? 0 : min(nbytes, (size_t)h + PKR_PROBE);
    if (h != ~0ull) pkr_find_newline(in, nbytes, (size_t)h, lim2, &s_second);
    __syncthreads();
    if (threadIdx.x == 0) {
        bool ok = nbytes > 0 && in[0] == '>' && h != ~0ull;
        unsigned long long W = 0, n = 0;
        if (ok) {
            const unsigned long long B = nbytes - h;
            bool only_line = false;                               // the first sequence line is also the last
            if (s_second != ~0ull) {
                W = s_second - h;
                only_line = s_second + 1 == nbytes;
            } else if (lim2 == nbytes) {                          // a single line without a final '\n' (or no sequence)
                W = B;
                only_line = true;
            } else ok = false;                                    // the first line is longer than the probe window
            if (only_line && W < (unsigned long long)PKR_MIN_W) W = PKR_MIN_W;   // (any width >= its length describes it)
            if (ok && B > 0) {
                if (W < (unsigned long long)PKR_MIN_W) ok = false;
                else {
                    const unsigned long long L = W + 1, full = B / L, rem = B % L;
                    n = full * W + rem - ((rem && in[nbytes - 1] == '\n') ? 1 : 0);
                }
            }
        }
        reg->h = ok ? h : 0;
        reg->W = ok ? (W ? W : 1) : 1;
        reg->n = ok ? n : 0;
        if (!ok) *irregular = 1;
        else {
            totals[1] = 1;       // records
            totals[2] = n;       // bases
        }
    }
}

// 4 bytes -> 2-bit codes (8 bits), valid flags (4 bits), "special" flags (4 bits: byte < 0x40 — every line break,
// '\r' and '>' is one; the caller looks at those bytes individually, there is about one per thread).
// x = bits 2..1 of a letter tell A, C, T, G apart (0, 1, 2, 3); a byte is a valid base iff its upper-case form equals
// the letter x stands for — one byte-permute looks the four expected letters up, another the four codes.
__device__ __forceinline__ void pkr_word(uint32_t w, uint32_t& code8, uint32_t& v4, uint32_t& sp4) {
    const uint32_t x = (w >> 1) & 0x03030303u;
    const uint32_t t = x | (x >> 4);
    const uint32_t sel = __byte_perm(t, 0u, 0x4420u);                     // nibble i = x of byte i
    const uint32_t expect = __byte_perm(0x47544341u, 0u, sel);            // 'A' 'C' 'T' 'G' by x
    const uint32_t vb = zero_bytes(expect ^ (w & 0xdfdfdfdfu)) >> 7;      // 1 per valid byte
    const uint32_t code = __byte_perm(0x02030100u, 0u, sel) & (vb | (vb << 1));   // A0 C1 G2 T3, 0 if invalid
    code8 = (code * 0x01041040u) >> 24;
    v4 = ((vb * 0x01020408u) >> 24) & 0xfu;
    const uint32_t sp = (~(w | (w << 1)) & 0x80808080u) >> 7;
    sp4 = ((sp * 0x01020408u) >> 24) & 0xfu;
}

__global__ void __launch_bounds__(PKR_THREADS)
k_pack_regular(const uint8_t* __restrict__ in, size_t nbytes, const PackReg* __restrict__ reg, size_t nwords,
               uint32_t* __restrict__ packed, uint32_t* __restrict__ valid_out, uint64_t* __restrict__ totals,
               uint32_t* __restrict__ irregular) {
    __shared__ __align__(16) uint8_t s_in[PKR_STAGE];
    // Persistent CTAs: the parameters are read once, and the bytes of the CTA's NEXT tile are loaded into registers
    // before the current tile is classified (a tile per CTA spent its life waiting: parameters, then bytes, then a
    // barrier — 19 stalled warps per issued instruction on the load scoreboard).
    const uint32_t irr0 = *irregular;
    const unsigned long long h = reg->h, W = reg->W, n = reg->n;
    if (irr0) return;                                             // (set by the probe: uniform)
    constexpr int NV = (PKR_STAGE / 16 + PKR_THREADS - 1) / PKR_THREADS;   // 16-byte loads per thread and tile (3)
    const size_t n_tiles = (nwords + PKR_THREADS - 1) / PKR_THREADS;
    uint32_t nv_total = 0, anybad = 0;
    uint4 pre[NV];
    auto tile_origin = [&](size_t tile, unsigned long long& j0, unsigned long long& q0, size_t& a0) {
        j0 = (unsigned long long)tile * PKR_BASES;
        q0 = j0 / W;
        a0 = (size_t)(h + j0 + q0) & ~(size_t)15;
    };
    auto prefetch = [&](size_t tile) {
        unsigned long long pj0, pq0;
        size_t pa0;
        tile_origin(tile, pj0, pq0, pa0);
#pragma unroll
        for (int q = 0; q < NV; q++) {
            const int i = threadIdx.x + q * PKR_THREADS;
            const size_t pos = pa0 + (size_t)i * 16;
            pre[q] = make_uint4(0x0a0a0a0au, 0x0a0a0a0au, 0x0a0a0a0au, 0x0a0a0a0au);
            if (tile < n_tiles && pj0 < n && i < PKR_STAGE / 16 && pos < nbytes) pre[q] = load_tile_bytes(in, nbytes, pos);
        }
    };
    prefetch(blockIdx.x);
    for (size_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    unsigned long long j0, q0;
    size_t a0;
    tile_origin(tile, j0, q0, a0);
    const size_t t = tile * PKR_THREADS + threadIdx.x;
    if (j0 >= n) {                                                // padding words past the sequence (block-uniform)
        if (t < nwords) {
            valid_out[t] = 0;
            *reinterpret_cast<uint2*>(packed + 2 * t) = make_uint2(0u, 0u);
        }
        continue;
    }
    // ---- stage the tile's bytes (aligned 16-byte loads, already in registers; past the end of the input: '\n') ----
#pragma unroll
    for (int q = 0; q < NV; q++) {
        const int i = threadIdx.x + q * PKR_THREADS;
        if (i < PKR_STAGE / 16) *reinterpret_cast<uint4*>(s_in + i * 16) = pre[q];
    }
    __syncthreads();
    prefetch(tile + gridDim.x);
    const unsigned long long j = j0 + 32ull * threadIdx.x;
    uint32_t vword = 0, bad = 0;
    uint64_t bits = 0;
    if (j < n) {
        // (j / W, j % W) from the CTA's quotient: j = j0 + d, d < 8192, W >= 16
        const uint32_t d = 32u * threadIdx.x;
        const unsigned long long r0 = j0 - q0 * W + d;            // < W + 8192
        const uint32_t dq = (uint32_t)(r0 / W), c0 = (uint32_t)(r0 - (unsigned long long)dq * W);
        const size_t s = (size_t)(h + j + q0 + dq);               // byte of base j
        const uint32_t nb = (uint32_t)min((unsigned long long)32, n - j);          // bases of this thread
        // bytes of this thread: its bases, the line breaks after those that end a line, and (last thread) the final '\n'
        const uint32_t nbrk = (c0 + nb) / (uint32_t)min(W, (unsigned long long)0xffffffffu);
        uint32_t span = nb + nbrk;
        uint64_t expect = 0;
        {
            unsigned long long pb = W - c0;                       // span index of the first predicted break
            for (uint32_t b = 0; b < nbrk; b++, pb += W + 1) expect |= 1ull << pb;
        }
        if (j + nb == n && s + span < nbytes) {                   // bytes after the last base: only one final '\n' is regular
            if (s + span + 1 == nbytes) { expect |= 1ull << span; span++; }
            else bad = 1;
        }
        const uint32_t lo = (uint32_t)(s - a0);
        const uint32_t* wsrc = reinterpret_cast<const uint32_t*>(s_in + (lo & ~3u));
        const uint32_t sh = 8u * (lo & 3u);
        uint32_t raw[10];
#pragma unroll
        for (int i = 0; i < 10; i++) raw[i] = wsrc[i];
        uint64_t code = 0, code_hi = 0;                            // 2 bits per byte: bytes 0..31, bytes 32..35
        uint64_t vm = 0, spm = 0;                                  // 1 bit per byte (36 bytes)
#pragma unroll
        for (int i = 0; i < 9; i++) {
            const uint32_t w = __funnelshift_r(raw[i], raw[i + 1], sh);
            uint32_t c8, v4, s4;
            pkr_word(w, c8, v4, s4);
            if (i < 8) code |= (uint64_t)c8 << (8 * i);
            else code_hi = c8;
            vm |= (uint64_t)v4 << (4 * i);
            spm |= (uint64_t)s4 << (4 * i);
        }
        const uint64_t smask = (1ull << span) - 1;                // span <= 35
        uint64_t nl = 0;
        spm &= smask;
        while (spm) {                                             // the bytes below 0x40 of the span, one by one
            const int q = __ffsll((long long)spm) - 1;
            spm &= spm - 1;
            const uint8_t c = s_in[lo + q];
            if (c == '\n') nl |= 1ull << q;
            else if (c == '\r' || c == '>') bad = 1;
        }
        if (nl != expect) bad = 1;                                // (expect has no bit at or above span)
        // drop the line-break bytes (<= 3 inside the span), highest first, then keep the first nb bases
        uint64_t drop = nl & smask;
        uint64_t vbits = vm & smask;
        while (drop) {
            const int q = 63 - __clzll(drop);
            drop &= ~(1ull << q);
            vbits = (vbits & ((1ull << q) - 1)) | ((vbits >> (q + 1)) << q);
            if (q < 32) {
                const uint64_t lo2 = code & ((1ull << (2 * q)) - 1);
                const uint64_t hi2 = (q == 31) ? 0ull : (code >> (2 * q + 2));
                code = lo2 | (hi2 << (2 * q)) | ((code_hi & 3ull) << 62);
                code_hi >>= 2;
            } else {
                const int qq = q - 32;
                code_hi = (code_hi & ((1ull << (2 * qq)) - 1)) | ((code_hi >> (2 * qq + 2)) << (2 * qq));
            }
        }
        const uint64_t keep = nb >= 32 ? ~0ull : ((1ull << (2 * nb)) - 1);
        bits = code & keep;
        vword = (uint32_t)(vbits & (nb >= 32 ? 0xffffffffull : ((1ull << nb) - 1)));
    }
    if (t < nwords) {
        valid_out[t] = vword;
        *reinterpret_cast<uint2*>(packed + 2 * t) = make_uint2((uint32_t)bits, (uint32_t)(bits >> 32));
    }
    nv_total += (uint32_t)__popc(vword);
    anybad |= bad;
    __syncthreads();                                              // s_in is rewritten by the next tile
    }
    // ---- totals ----
    uint64_t nv = spk_warp_sum_u64((uint64_t)nv_total);
    anybad = __ballot_sync(0xffffffffu, anybad != 0);
    if ((threadIdx.x & 31) == 0) {
        if (nv) atomicAdd((unsigned long long*)&totals[0], (unsigned long long)nv);
        if (anybad) atomicOr(irregular, 1u);
    }
}

// the three-pass path runs only when the single pass gave up: its outputs and totals are reset on the device
__global__ void __launch_bounds__(256) k_zero_if(const uint32_t* __restrict__ flag, uint4* __restrict__ p, size_t n16) {
    if (*flag == 0) return;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x)
        p[i] = make_uint4(0u, 0u, 0u, 0u);
}
__global__ void k_reset_totals_if(const uint32_t* flag, uint64_t* totals) {
    if (*flag) totals[0] = totals[1] = totals[2] = 0;
}

__global__ void k_write_info(const uint64_t* tile_off, size_t ntiles, const uint64_t* totals, const uint32_t* slow,
                             const uint32_t* irregular, uint64_t* info) {
    info[0] = (slow && *slow == 0) ? totals[2] : tile_off[ntiles];
    info[1] = totals[0];
    info[2] = totals[1];
    // which kernel produced the result: 0 regular-layout, 2 general single pass, 1 three-pass
    info[3] = (!slow || *slow) ? 1 : ((irregular && *irregular == 0) ? 0 : 2);
}

}  // namespace

extern "C" size_t spk_packed_words(uint64_t n_bases) {
    const uint64_t tiles = (n_bases + SPK_TILE_BASES - 1) / SPK_TILE_BASES + 1;
    return (size_t)(tiles * (SPK_TILE_BASES / 16) + 64);
}
extern "C" size_t spk_valid_words(uint64_t n_bases) {
    const uint64_t tiles = (n_bases + SPK_TILE_BASES - 1) / SPK_TILE_BASES + 1;
    return (size_t)(tiles * (SPK_TILE_BASES / 32) + 32);
}
extern "C" size_t spk_pack_workspace_bytes(size_t nbytes) {
    size_t total = 0;
    carve_ws(nullptr, nbytes, &total);
    return total;
}

extern "C" int spk_pack_fasta(const uint8_t* d_ascii, size_t nbytes, uint32_t* d_packed,
                              uint32_t* d_valid, uint64_t cap_bases, uint64_t* d_info, void* d_ws,
                              size_t ws_bytes, void* stream) {
    SPK_CHECK_ARG(d_packed && d_valid && d_info && d_ws, "null pointer");
    SPK_CHECK_ARG(nbytes == 0 || d_ascii, "null input");
    SPK_CHECK_ARG(((uintptr_t)d_ascii & 15) == 0, "d_ascii must be 16-byte aligned");
    SPK_CHECK_ARG(cap_bases >= nbytes, "cap_bases must be >= nbytes (upper bound on bases)");
    size_t need = 0;
    PackWs w = carve_ws(d_ws, nbytes, &need);
    if (ws_bytes < need) {
        spk_set_error("spk_pack_fasta: workspace %zu < %zu bytes", ws_bytes, need);
        return SPK_ECAP;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const size_t ntiles = (nbytes + PK_TILE - 1) / PK_TILE;
    const int sms = spk_num_sms();
    SPK_CUDA(cudaMemsetAsync(w.totals, 0, 128, st));
    uint32_t* ticket = (uint32_t*)(w.totals + 4);
    uint32_t* need_slow = ticket + 1;
    uint32_t* irregular = ticket + 2;
    PackReg* reg = (PackReg*)(w.totals + 8);
    const char* mode = getenv("SPK_PACK_MODE");            // tests: "3pass" = three-pass only, "single" = skip the regular-layout kernel
    const bool single = !(mode && mode[0] == '3');
    const bool regular = single && !(mode && mode[0] == 's') && ntiles > 0;
    // outputs of the general kernels are OR-accumulated: the whole tile-aligned extent must be zero (this is also the
    // padding); the regular-layout kernel writes every word of the extent itself
    const size_t pbytes = spk_packed_words(cap_bases) * 4, vbytes = spk_valid_words(cap_bases) * 4;
    const uint32_t* run_single = nullptr;
    if (regular) {
        const size_t nwords = spk_valid_words(cap_bases);           // (packed words = 2 x validity words)
        k_pack_probe<<<1, 256, 0, st>>>(d_ascii, nbytes, reg, irregular, w.totals);
        SPK_LAUNCH_CHECK();
        k_pack_regular<<<(unsigned)min((size_t)sms * 5, (nwords + PKR_THREADS - 1) / PKR_THREADS), PKR_THREADS, 0, st>>>(
            d_ascii, nbytes, reg, nwords, d_packed, d_valid, w.totals, irregular);
        SPK_LAUNCH_CHECK();
        run_single = irregular;
        k_zero_if<<<sms * 4, 256, 0, st>>>(irregular, (uint4*)d_packed, pbytes / 16);
        SPK_LAUNCH_CHECK();
        k_zero_if<<<sms * 4, 256, 0, st>>>(irregular, (uint4*)d_valid, vbytes / 16);
        SPK_LAUNCH_CHECK();
        k_reset_totals_if<<<1, 1, 0, st>>>(irregular, w.totals);
        SPK_LAUNCH_CHECK();
    } else {
        SPK_CUDA(cudaMemsetAsync(d_packed, 0, pbytes, st));
        SPK_CUDA(cudaMemsetAsync(d_valid, 0, vbytes, st));
    }
    const uint32_t* run_if = nullptr;
    if (single && ntiles > 0) {
        uint64_t* status = (uint64_t*)w.tile_last_nl;      // (one 64-bit word per PK_SUB tiles fits the per-tile array)
        const size_t nunits = (ntiles + PK_SUB - 1) / PK_SUB;
        SPK_CUDA(cudaMemsetAsync(status, 0, nunits * 8, st));
        const unsigned grid1 = (unsigned)min((size_t)sms * 6, nunits);
        k_pack_single<<<grid1, PK_THREADS, 0, st>>>(d_ascii, nbytes, nunits, status, ticket, d_packed, d_valid, w.totals,
                                                    need_slow, run_single);
        SPK_LAUNCH_CHECK();
        // fallback (lines longer than the look-back window), decided on the device: every kernel below returns at once
        // unless *need_slow is set
        run_if = need_slow;
        k_zero_if<<<sms * 4, 256, 0, st>>>(need_slow, (uint4*)d_packed, pbytes / 16);
        SPK_LAUNCH_CHECK();
        k_zero_if<<<sms * 4, 256, 0, st>>>(need_slow, (uint4*)d_valid, vbytes / 16);
        SPK_LAUNCH_CHECK();
        k_reset_totals_if<<<1, 1, 0, st>>>(need_slow, w.totals);
        SPK_LAUNCH_CHECK();
    }
    if (ntiles > 0) {
        const unsigned grid = (unsigned)min((size_t)sms * 8, ntiles);
        k_tile_last_newline<<<grid, PK_THREADS, 0, st>>>(d_ascii, nbytes, ntiles, w.tile_last_nl, w.tile_skip, run_if);
        SPK_LAUNCH_CHECK();
        k_scan_max_excl<<<1, 1024, 0, st>>>(w.tile_last_nl, ntiles, run_if);
        SPK_LAUNCH_CHECK();
        k_classify<false><<<grid, PK_THREADS, 0, st>>>(d_ascii, nbytes, ntiles, w.tile_last_nl, w.tile_skip,
                                                       w.tile_kept, nullptr, nullptr, nullptr, w.totals, run_if);
        SPK_LAUNCH_CHECK();
    }
    k_scan_sum_excl<<<1, 1024, 0, st>>>(w.tile_kept, w.tile_off, ntiles, ntiles > 0 ? run_if : nullptr);
    SPK_LAUNCH_CHECK();
    if (ntiles > 0) {
        const unsigned grid = (unsigned)min((size_t)sms * 8, ntiles);
        k_classify<true><<<grid, PK_THREADS, 0, st>>>(d_ascii, nbytes, ntiles, w.tile_last_nl, w.tile_skip, nullptr,
                                                      w.tile_off, d_packed, d_valid, w.totals, run_if);
        SPK_LAUNCH_CHECK();
    }
    k_write_info<<<1, 1, 0, st>>>(w.tile_off, ntiles, w.totals, ntiles > 0 ? run_if : nullptr, run_single, d_info);
    SPK_LAUNCH_CHECK();
    return SPK_OK;
}
