// Text wire formats on the device: the rows of `.kmer.mat` (JellyfishDumps.write_matrix, Jellyfish.py:515-520) and of
// `.sig.kmer-subgenome.tsv` (Cluster.output_kmers, Cluster.py:158-172) are formatted by the GPU, byte for byte what
// Python's `'\t'.join(map(str, ...))` writes — k-mer string from the 2-bit key, floats by shortest round-trip repr
// (spk_format.cuh) — instead of 10^7..10^8 `repr()` calls on the host.
//
// One thread formats one row; the same code runs twice through a counting and a writing sink (pass 1: row lengths,
// exclusive scan by the caller, pass 2: bytes at their final offsets), so the output is one contiguous text buffer
// that goes to the file in a single write.
#include "spk_common.cuh"
#include "spk_format.cuh"

namespace {

struct CountSink {
    uint32_t n = 0;
    __device__ __forceinline__ void put(char) { n++; }
    __device__ __forceinline__ void put_n(const char*, int len) { n += len; }
};
struct WriteSink {
    char* p;
    __device__ __forceinline__ void put(char c) { *p++ = c; }
    __device__ __forceinline__ void put_n(const char* s, int len) {
        for (int i = 0; i < len; i++) *p++ = s[i];
    }
};

struct RowArgs {
    const uint64_t* keys;       // [M] canonical k-mers (first base most significant)
    const double* vals;         // [M, n]
    const uint32_t* rows;       // optional subset: row ids (ascending); nullptr = all rows
    uint64_t n_rows;
    int n, k, kind;             // kind 0: KMER (\t v)*n \n ; kind 1: KMER \t LABEL \t p \t v1,v2,..,vn \n
    const int32_t* label;       // kind 1: [M] label index of the row
    const char* label_text;     // kind 1: [L][16] label strings
    const int32_t* label_len;   // kind 1: [L]
    const double* pval;         // kind 1: [M]
};

template <typename Sink>
__device__ __forceinline__ void emit_row(const RowArgs& a, uint64_t r, Sink& s) {
    const uint64_t key = a.keys[r];
    for (int i = a.k - 1; i >= 0; i--) s.put("ACGT"[(key >> (2 * i)) & 3]);
    char buf[spkfmt::PY_REPR_MAX];
    if (a.kind == 1) {
        s.put('\t');
        const int l = a.label[r];
        s.put_n(a.label_text + 16 * l, a.label_len[l]);
        s.put('\t');
        s.put_n(buf, spkfmt::py_repr(a.pval[r], buf));
    }
    const double* v = a.vals + r * (uint64_t)a.n;
    for (int c = 0; c < a.n; c++) {
        s.put((a.kind == 1 && c > 0) ? ',' : '\t');
        s.put_n(buf, spkfmt::py_repr(v[c], buf));
    }
    s.put('\n');
}

__global__ void __launch_bounds__(128) k_format_sizes(RowArgs a, uint32_t* __restrict__ row_len) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n_rows; i += (uint64_t)gridDim.x * blockDim.x) {
        CountSink s;
        emit_row(a, a.rows ? a.rows[i] : i, s);
        row_len[i] = s.n;
    }
}

__global__ void __launch_bounds__(128)
k_format_write(RowArgs a, const uint64_t* __restrict__ row_off, char* __restrict__ out) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n_rows; i += (uint64_t)gridDim.x * blockDim.x) {
        WriteSink s{out + row_off[i]};
        emit_row(a, a.rows ? a.rows[i] : i, s);
    }
}

}  // namespace

extern "C" int spk_format_rows(const uint64_t* d_keys, const double* d_vals, uint64_t M, int n, int k, int kind,
                               const uint32_t* d_rows, uint64_t n_rows, const int32_t* d_label, const char* d_label_text,
                               const int32_t* d_label_len, const double* d_pval, const uint64_t* d_row_off,
                               uint32_t* d_row_len, char* d_out, void* stream) {
    SPK_CHECK_ARG(k >= 1 && k <= 32 && n >= 1, "bad shape");
    SPK_CHECK_ARG(kind == 0 || kind == 1, "kind: 0 matrix rows, 1 specific-k-mer rows");
    SPK_CHECK_ARG(kind == 0 || (d_label && d_label_text && d_label_len && d_pval), "kind 1 needs labels and p-values");
    SPK_CHECK_ARG((d_row_len != nullptr) != (d_out != nullptr), "pass d_row_len (size pass) or d_row_off + d_out (write pass)");
    SPK_CHECK_ARG(!d_out || d_row_off, "write pass needs the row offsets");
    if (d_rows == nullptr) n_rows = M;
    if (n_rows == 0) return SPK_OK;
    SPK_CHECK_ARG(d_keys && d_vals, "null pointer");
    RowArgs a{d_keys, d_vals, d_rows, n_rows, n, k, kind, d_label, d_label_text, d_label_len, d_pval};
    const unsigned grid = (unsigned)min((n_rows + 127) / 128, (uint64_t)spk_num_sms() * 32);
    if (d_row_len) k_format_sizes<<<grid, 128, 0, (cudaStream_t)stream>>>(a, d_row_len);
    else k_format_write<<<grid, 128, 0, (cudaStream_t)stream>>>(a, d_row_off, d_out);
    SPK_LAUNCH_CHECK();
    return SPK_OK;
}
