// K1: FASTA bytes -> 2-bit packed bases + validity bits, entirely on the device.
//
// Replaces the text ingest the reference delegates to `cat f | jellyfish count ... /dev/stdin`
// (Jellyfish.py:697) and to SeqIO.parse + upper() (Seqs.py:121-139).
//
// Four streaming passes over 4-KiB byte tiles (256 threads x 16 B, one coalesced uint4 each):
//   A  last '\n' of every tile                      -> exclusive max-scan  (which line does a tile start in)
//   B  kept bases / valid bases / headers per tile  -> exclusive sum-scan  (where does a tile write)
//   C  emit one code byte per kept base (0..3 = ACGT, 4 = invalid) at its final index
//   D  codes -> 2-bit words + validity bits (zero padded to the tile-aligned extent)
// A byte is in a header iff the first byte of its line is '>'.
#include "spk_common.cuh"

namespace {

constexpr int PK_THREADS = 256;
constexpr int PK_BYTES_PER_THREAD = 16;
constexpr int PK_TILE = PK_THREADS * PK_BYTES_PER_THREAD;  // 4096

struct PackWs {
    int64_t* tile_last_nl;   // [ntiles]  global position of last '\n' in tile (or -1); scanned in place
    uint64_t* tile_off;      // [ntiles+1] kept bases before tile (after scan)
    uint32_t* tile_kept;     // [ntiles]
    uint8_t* codes;          // [nbytes]
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
    w.totals = (uint64_t*)(p + off);
    off += 256;
    w.codes = (uint8_t*)(p + off);
    off += align_up(nbytes + 64, 256);
    if (total) *total = off;
    return w;
}

__device__ __forceinline__ uint4 load_tile_bytes(const uint8_t* __restrict__ in, size_t nbytes,
                                                 size_t pos) {
    // 16 bytes starting at pos (pos is 16-aligned relative to the buffer start, which is >=16-aligned)
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

__device__ __forceinline__ uint8_t byte_of(const uint4& v, int i) {
    const uint32_t w = (i < 4) ? v.x : (i < 8) ? v.y : (i < 12) ? v.z : v.w;
    return (uint8_t)(w >> (8 * (i & 3)));
}

// ---- pass A -------------------------------------------------------------------------------------
__global__ void __launch_bounds__(PK_THREADS) k_tile_last_newline(const uint8_t* __restrict__ in,
                                                                   size_t nbytes, size_t ntiles,
                                                                   int64_t* __restrict__ tile_last_nl) {
    __shared__ int s_max[PK_THREADS / 32];
    for (size_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const size_t pos = tile * PK_TILE + (size_t)threadIdx.x * PK_BYTES_PER_THREAD;
        int last = -1;
        if (pos < nbytes) {
            const uint4 v = load_tile_bytes(in, nbytes, pos);
#pragma unroll
            for (int i = 0; i < 16; i++)
                if (byte_of(v, i) == '\n' && pos + i < nbytes) last = threadIdx.x * 16 + i;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) last = max(last, __shfl_xor_sync(0xffffffffu, last, o));
        if ((threadIdx.x & 31) == 0) s_max[threadIdx.x >> 5] = last;
        __syncthreads();
        if (threadIdx.x == 0) {
            int m = -1;
            for (int w = 0; w < PK_THREADS / 32; w++) m = max(m, s_max[w]);
            tile_last_nl[tile] = (m < 0) ? -1 : (int64_t)(tile * PK_TILE + m);
        }
        __syncthreads();
    }
}

// single-CTA exclusive max-scan (int64) / sum-scan (u32 -> u64); tile counts are small (bytes/4096)
__global__ void __launch_bounds__(1024) k_scan_max_excl(int64_t* data, size_t n) {
    __shared__ int64_t s_warp[32];
    __shared__ int64_t s_carry;
    if (threadIdx.x == 0) s_carry = -1;
    __syncthreads();
    for (size_t base = 0; base < n; base += 1024) {
        const size_t i = base + threadIdx.x;
        const int64_t v = (i < n) ? data[i] : -1;
        int64_t incl = v;
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
        excl = max(excl, prefix);
        if (i < n) data[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = max(prefix, incl);
        __syncthreads();
    }
}

__global__ void __launch_bounds__(1024) k_scan_sum_excl(const uint32_t* in, uint64_t* out, size_t n) {
    __shared__ uint64_t s_warp[32];
    __shared__ uint64_t s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (size_t base = 0; base < n; base += 1024) {
        const size_t i = base + threadIdx.x;
        const uint64_t v = (i < n) ? in[i] : 0;
        uint64_t incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint64_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if ((threadIdx.x & 31) >= o) incl += t;
        }
        if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = incl;
        __syncthreads();
        uint64_t prefix = s_carry;
        for (int w = 0; w < (int)(threadIdx.x >> 5); w++) prefix += s_warp[w];
        if (i < n) out[i] = prefix + incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = prefix + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) out[n] = s_carry;
}

// Classify the 16 bytes of one thread.  codes[i]: 0..3 base, 4 invalid-but-kept, 255 dropped.
// `ls` = global position of the last '\n' before this thread's first byte (-1: none).
__device__ __forceinline__ void classify16(const uint8_t* __restrict__ in, size_t nbytes, size_t pos,
                                           const uint4& v, int64_t ls, uint8_t* codes, int& kept,
                                           int& valid, int& headers) {
    kept = valid = headers = 0;
    bool hdr = false;
    {
        const size_t line_start = (size_t)(ls + 1);
        hdr = (line_start < pos) ? (in[line_start] == '>') : false;  // line_start == pos handled below
    }
#pragma unroll
    for (int i = 0; i < 16; i++) {
        const size_t p = pos + i;
        const uint8_t c = byte_of(v, i);
        uint8_t code = 255;
        if (p < nbytes) {
            const bool at_line_start = ((int64_t)p == ls + 1);
            if (c == '\n') {
                ls = (int64_t)p;
                hdr = false;
            } else if (at_line_start && c == '>') {
                hdr = true;
                headers++;
                if (p != 0) code = 4;  // record separator
            } else if (!hdr && c != '\r') {
                switch (c) {
                    case 'A': case 'a': code = 0; break;
                    case 'C': case 'c': code = 1; break;
                    case 'G': case 'g': code = 2; break;
                    case 'T': case 't': code = 3; break;
                    default: code = 4;
                }
            }
        }
        codes[i] = code;
        kept += (code != 255);
        valid += (code < 4);
    }
}

// exclusive max-scan of per-thread last-newline positions inside a CTA, seeded with the tile carry
__device__ __forceinline__ int64_t block_excl_max(int64_t v, int64_t carry, int64_t* s_warp) {
    int64_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int64_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if ((threadIdx.x & 31) >= o) incl = max(incl, t);
    }
    if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = incl;
    __syncthreads();
    int64_t prefix = carry;
    for (int w = 0; w < (int)(threadIdx.x >> 5); w++) prefix = max(prefix, s_warp[w]);
    int64_t excl = __shfl_up_sync(0xffffffffu, incl, 1);
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
                                                          uint32_t* __restrict__ tile_kept,
                                                          const uint64_t* __restrict__ tile_off,
                                                          uint8_t* __restrict__ codes_out,
                                                          uint64_t* __restrict__ totals) {
    __shared__ int64_t s_w64[PK_THREADS / 32];
    __shared__ uint32_t s_w32[PK_THREADS / 32];
    uint64_t my_valid = 0, my_hdr = 0;
    for (size_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const size_t pos = tile * PK_TILE + (size_t)threadIdx.x * PK_BYTES_PER_THREAD;
        uint4 v = make_uint4(0x0a0a0a0a, 0x0a0a0a0a, 0x0a0a0a0a, 0x0a0a0a0a);
        int64_t my_last = -1;
        if (pos < nbytes) {
            v = load_tile_bytes(in, nbytes, pos);
#pragma unroll
            for (int i = 0; i < 16; i++)
                if (byte_of(v, i) == '\n' && pos + i < nbytes) my_last = (int64_t)(pos + i);
        }
        const int64_t ls = block_excl_max(my_last, tile_carry_nl[tile], s_w64);
        uint8_t codes[16];
        int kept, valid, headers;
        classify16(in, nbytes, pos, v, ls, codes, kept, valid, headers);
        if (!EMIT) {
            uint32_t tot;
            block_excl_sum((uint32_t)kept, s_w32, &tot);
            if (threadIdx.x == 0) tile_kept[tile] = tot;
            my_valid += valid;
            my_hdr += headers;
        } else {
            const uint32_t off = block_excl_sum((uint32_t)kept, s_w32, nullptr);
            uint8_t* dst = codes_out + tile_off[tile] + off;
            int j = 0;
#pragma unroll
            for (int i = 0; i < 16; i++)
                if (codes[i] != 255) dst[j++] = codes[i];
        }
    }
    if (!EMIT) {
        my_valid = spk_warp_sum_u64(my_valid);
        my_hdr = spk_warp_sum_u64(my_hdr);
        if ((threadIdx.x & 31) == 0) {
            if (my_valid) atomicAdd((unsigned long long*)&totals[0], (unsigned long long)my_valid);
            if (my_hdr) atomicAdd((unsigned long long*)&totals[1], (unsigned long long)my_hdr);
        }
    }
}

__global__ void k_write_info(const uint64_t* tile_off, size_t ntiles, const uint64_t* totals,
                             uint64_t* info) {
    info[0] = tile_off[ntiles];
    info[1] = totals[0];
    info[2] = totals[1];
    info[3] = 0;
}

// ---- pass D --------------------------------------------------------------------------------------
// one thread = 32 bases -> 2 packed words + 1 validity word
__global__ void __launch_bounds__(256) k_pack_codes(const uint8_t* __restrict__ codes,
                                                     const uint64_t* __restrict__ info,
                                                     uint32_t* __restrict__ packed,
                                                     uint32_t* __restrict__ valid, size_t n_groups) {
    const uint64_t n = info[0];
    for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < n_groups;
         g += (size_t)gridDim.x * blockDim.x) {
        const uint64_t base = (uint64_t)g * 32;
        uint32_t w0 = 0, w1 = 0, vm = 0;
        if (base < n) {
            uint4 a, b;
            if (base + 32 <= n) {
                a = *reinterpret_cast<const uint4*>(codes + base);
                b = *reinterpret_cast<const uint4*>(codes + base + 16);
            } else {
                uint8_t tmp[32];
#pragma unroll
                for (int i = 0; i < 32; i++) tmp[i] = (base + i < n) ? codes[base + i] : (uint8_t)4;
                a = *reinterpret_cast<uint4*>(tmp);
                b = *reinterpret_cast<uint4*>(tmp + 16);
            }
#pragma unroll
            for (int i = 0; i < 16; i++) {
                const uint32_t c = byte_of(a, i);
                if (c < 4) {
                    w0 |= c << (2 * i);
                    vm |= 1u << i;
                }
            }
#pragma unroll
            for (int i = 0; i < 16; i++) {
                const uint32_t c = byte_of(b, i);
                if (c < 4) {
                    w1 |= c << (2 * i);
                    vm |= 1u << (16 + i);
                }
            }
        }
        packed[2 * g] = w0;
        packed[2 * g + 1] = w1;
        valid[g] = vm;
    }
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
    SPK_CUDA(cudaMemsetAsync(w.totals, 0, 32, st));
    if (ntiles > 0) {
        const unsigned grid = (unsigned)min((size_t)sms * 8, ntiles);
        k_tile_last_newline<<<grid, PK_THREADS, 0, st>>>(d_ascii, nbytes, ntiles, w.tile_last_nl);
        SPK_LAUNCH_CHECK();
        k_scan_max_excl<<<1, 1024, 0, st>>>(w.tile_last_nl, ntiles);
        SPK_LAUNCH_CHECK();
        k_classify<false><<<grid, PK_THREADS, 0, st>>>(d_ascii, nbytes, ntiles, w.tile_last_nl,
                                                       w.tile_kept, nullptr, nullptr, w.totals);
        SPK_LAUNCH_CHECK();
    }
    k_scan_sum_excl<<<1, 1024, 0, st>>>(w.tile_kept, w.tile_off, ntiles);
    SPK_LAUNCH_CHECK();
    if (ntiles > 0) {
        const unsigned grid = (unsigned)min((size_t)sms * 8, ntiles);
        k_classify<true><<<grid, PK_THREADS, 0, st>>>(d_ascii, nbytes, ntiles, w.tile_last_nl,
                                                      nullptr, w.tile_off, w.codes, nullptr);
        SPK_LAUNCH_CHECK();
    }
    k_write_info<<<1, 1, 0, st>>>(w.tile_off, ntiles, w.totals, d_info);
    SPK_LAUNCH_CHECK();
    // pack the whole tile-aligned extent of cap_bases so padding is zero
    const size_t n_groups = spk_valid_words(cap_bases);  // 32 bases per group; packed has 2x words
    {
        const size_t blocks = (n_groups + 255) / 256;
        const unsigned grid = (unsigned)min((size_t)sms * 16, max(blocks, (size_t)1));
        k_pack_codes<<<grid, 256, 0, st>>>(w.codes, d_info, d_packed, d_valid, n_groups);
        SPK_LAUNCH_CHECK();
    }
    return SPK_OK;
}
