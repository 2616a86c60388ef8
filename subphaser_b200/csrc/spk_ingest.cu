// Genome ingest on the device: the reference splits multi-record genome files into one FASTA per
// chromosome with BioPython on the host (Seqs.split_genomes, Seqs.py:27-71).  Here the file's bytes go to the GPU
// once; these kernels find the records and tell whether a record's body is already laid out the way BioPython
// would write it (60 columns, '\n'), in which case the per-chromosome file is a verbatim byte range of the input.
// The records themselves are packed by K1 (spk_pack_fasta) straight from the same device buffer.
#include "spk_common.cuh"

namespace {

// record starts: '>' at byte 0 or right after a '\n'
__global__ void __launch_bounds__(256)
k_fasta_record_starts(const uint8_t* __restrict__ in, uint64_t nbytes, uint64_t* __restrict__ pos, uint64_t cap,
                      unsigned long long* __restrict__ count) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nbytes; i += (uint64_t)gridDim.x * blockDim.x) {
        if (in[i] == '>' && (i == 0 || in[i - 1] == '\n')) {
            const unsigned long long at = atomicAdd(count, 1ull);
            if (at < cap) pos[at] = i;
        }
    }
}

// body [beg, end): bit 0 = a line break that is not at a multiple of (width + 1) bytes (other than the last byte),
// or a full-width position without one; bit 1 = a byte BioPython would drop or that is not plain sequence text
// ('\r', ' ', '\t', '>'); out[1] = number of '\n' bytes (body length - that = sequence length)
__global__ void __launch_bounds__(256)
k_fasta_wrap_check(const uint8_t* __restrict__ in, uint64_t beg, uint64_t end, uint32_t width,
                   unsigned long long* __restrict__ out) {
    uint32_t flags = 0;
    unsigned long long nl = 0;
    const uint64_t len = end - beg;
    for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < len; j += (uint64_t)gridDim.x * blockDim.x) {
        const uint8_t c = in[beg + j];
        const bool expect = (j % (width + 1)) == width;
        const bool is_nl = c == '\n';
        nl += is_nl;
        if (is_nl != expect && !(is_nl && j + 1 == len)) flags |= 1u;
        if (c == '\r' || c == ' ' || c == '\t' || c == '>') flags |= 2u;
    }
    flags = __reduce_or_sync(0xffffffffu, flags);
    nl = spk_warp_sum_u64(nl);
    if ((threadIdx.x & 31) == 0) {
        if (flags) atomicOr(&out[0], (unsigned long long)flags);
        if (nl) atomicAdd(&out[1], nl);
    }
}

}  // namespace

extern "C" int spk_fasta_record_starts(const uint8_t* d_ascii, uint64_t nbytes, uint64_t* d_pos, uint64_t cap,
                                       uint64_t* d_count, void* stream) {
    SPK_CHECK_ARG(d_count && (cap == 0 || d_pos), "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    SPK_CUDA(cudaMemsetAsync(d_count, 0, 8, st));
    if (nbytes == 0) return SPK_OK;
    SPK_CHECK_ARG(d_ascii, "null input");
    const unsigned grid = (unsigned)min((nbytes + 255) / 256, (uint64_t)spk_num_sms() * 32);
    k_fasta_record_starts<<<grid, 256, 0, st>>>(d_ascii, nbytes, d_pos, cap, (unsigned long long*)d_count);
    SPK_LAUNCH_CHECK();
    return SPK_OK;
}

extern "C" int spk_fasta_wrap_check(const uint8_t* d_ascii, uint64_t beg, uint64_t end, uint32_t width, uint64_t* d_out,
                                    void* stream) {
    SPK_CHECK_ARG(d_out && width >= 1 && end >= beg, "bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    SPK_CUDA(cudaMemsetAsync(d_out, 0, 16, st));
    if (end == beg) return SPK_OK;
    SPK_CHECK_ARG(d_ascii, "null input");
    const unsigned grid = (unsigned)min((end - beg + 255) / 256, (uint64_t)spk_num_sms() * 32);
    k_fasta_wrap_check<<<grid, 256, 0, st>>>(d_ascii, beg, end, width, (unsigned long long*)d_out);
    SPK_LAUNCH_CHECK();
    return SPK_OK;
}
