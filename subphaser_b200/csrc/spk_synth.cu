// Synthetic chromosome generator — TEST / BENCH INFRASTRUCTURE, not part of the product path.
//
// Real genomes cannot be downloaded here, and 14 Gb of FASTA is impractical to build on the host, so
// bench.py and the GPU tests synthesise wheat-shaped chromosomes directly in device memory: a
// background of iid bases with copies of subgenome-specific / shared repeat families pasted in
// (per-copy substitutions at rate `div`), runs of N, soft-masked (lower-case) blocks, FASTA line
// wrapping and a header — exactly the byte stream a chromosome file would hold.  The segment table is
// drawn on the host (numpy) and everything per-base is a counter-based hash, so any byte can be
// regenerated independently.
#include "spk_common.cuh"

namespace {

__device__ __forceinline__ uint64_t mix(uint64_t a, uint64_t b) {
    return spk_hash64(a * 0x9E3779B97F4A7C15ull + b + 0x632BE59BD9B4E019ull);
}

__global__ void __launch_bounds__(256)
k_synth_fasta(uint8_t* __restrict__ out, uint64_t nbytes, uint64_t header_len, uint64_t n_bases,
              int lw, const uint64_t* __restrict__ seg_start, const int64_t* __restrict__ seg_src,
              const uint32_t* __restrict__ seg_seed, uint64_t n_segs,
              const uint8_t* __restrict__ library, uint64_t lib_len,
              const uint64_t* __restrict__ nrun_start, const uint64_t* __restrict__ nrun_end,
              uint64_t n_nruns, double div, double soft_frac, uint64_t seed) {
    const uint64_t div_thr = (uint64_t)(div * 18446744073709551615.0);
    const uint64_t soft_thr = (uint64_t)(soft_frac * 18446744073709551615.0);
    for (uint64_t o = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x + header_len; o < nbytes;
         o += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t q = o - header_len;
        const uint64_t line = q / (uint64_t)(lw + 1);
        const uint64_t col = q - line * (uint64_t)(lw + 1);
        const uint64_t i = line * lw + col;
        uint8_t ch;
        if (col == (uint64_t)lw || i >= n_bases) {
            ch = '\n';
        } else {
            // N runs
            bool isn = false;
            if (n_nruns) {
                uint64_t lo = 0, hi = n_nruns;
                while (lo < hi) {  // first run with start > i
                    const uint64_t mid = (lo + hi) >> 1;
                    if (nrun_start[mid] <= i) lo = mid + 1;
                    else hi = mid;
                }
                if (lo > 0 && i < nrun_end[lo - 1]) isn = true;
            }
            if (isn) {
                ch = 'N';
            } else {
                uint64_t lo = 0, hi = n_segs;
                while (lo < hi) {  // first segment with start > i
                    const uint64_t mid = (lo + hi) >> 1;
                    if (seg_start[mid] <= i) lo = mid + 1;
                    else hi = mid;
                }
                uint32_t b;
                if (lo == 0 || seg_src[lo - 1] < 0) {
                    b = (uint32_t)(mix(seed, i) & 3);
                } else {
                    const uint64_t s = lo - 1;
                    const uint64_t within = i - seg_start[s];
                    uint64_t src = (uint64_t)seg_src[s] + within;
                    if (src >= lib_len) src = lib_len - 1;
                    b = library[src] & 3;
                    const uint64_t h = mix(seed ^ ((uint64_t)seg_seed[s] << 32), within);
                    if (h < div_thr) b = (b + 1 + (uint32_t)((h >> 7) % 3)) & 3;
                }
                ch = "ACGT"[b];
                if (mix(seed ^ 0xABCDEFull, i >> 9) < soft_thr) ch |= 0x20;
            }
        }
        out[o] = ch;
    }
}

}  // namespace

extern "C" int spk_synth_fasta(uint8_t* d_out, uint64_t nbytes, uint64_t header_len, uint64_t n_bases,
                               int line_width, const uint64_t* d_seg_start, const int64_t* d_seg_src,
                               const uint32_t* d_seg_seed, uint64_t n_segs, const uint8_t* d_library,
                               uint64_t lib_len, const uint64_t* d_nrun_start,
                               const uint64_t* d_nrun_end, uint64_t n_nruns, double div,
                               double soft_frac, uint64_t seed, void* stream) {
    SPK_CHECK_ARG(d_out, "null output");
    SPK_CHECK_ARG(line_width >= 1, "line_width must be >= 1");
    SPK_CHECK_ARG(n_segs == 0 || (d_seg_start && d_seg_src && d_seg_seed), "null segment table");
    SPK_CHECK_ARG(n_nruns == 0 || (d_nrun_start && d_nrun_end), "null N-run table");
    if (nbytes <= header_len) return SPK_OK;
    const uint64_t blocks = (nbytes - header_len + 255) / 256;
    const unsigned grid = (unsigned)min(blocks, (uint64_t)spk_num_sms() * 32);
    k_synth_fasta<<<grid, 256, 0, (cudaStream_t)stream>>>(
        d_out, nbytes, header_len, n_bases, line_width, d_seg_start, d_seg_src, d_seg_seed, n_segs,
        d_library, lib_len, d_nrun_start, d_nrun_end, n_nruns, div, soft_frac, seed);
    SPK_LAUNCH_CHECK();
    return SPK_OK;
}
