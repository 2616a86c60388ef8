// Host-buffer entry point: one chromosome's FASTA bytes in host memory -> counted and dumped on the
// device.  This is the call a file-level binding makes in place of the shell-out of
// Jellyfish.run_jellyfish_dump (Jellyfish.py:681-704): H2D copy, K1 pack, K2 count, K3 scan + extract.
#include "spk_common.cuh"

extern "C" int spk_count_fasta_host(const uint8_t* h_fasta, size_t nbytes, int k, uint32_t lower_count,
                                    uint8_t* d_ascii, uint32_t* d_packed, uint32_t* d_valid,
                                    uint64_t cap_bases, void* d_ws, size_t ws_bytes, void* d_table,
                                    size_t table_bytes, uint32_t* d_block_counts, uint64_t* d_keys,
                                    uint32_t* d_counts, uint64_t cap_out, uint64_t* d_info,
                                    uint64_t* h_out, void* stream) {
    SPK_CHECK_ARG(h_fasta && d_ascii && d_info && h_out && d_table, "null pointer");
    SPK_CHECK_ARG(k >= 1 && k <= 32, "k must be in [1, 32]");
    SPK_CHECK_ARG(cap_bases >= nbytes, "cap_bases must be >= nbytes");
    cudaStream_t st = (cudaStream_t)stream;
    const int layout = spk_count_layout(cap_bases, k);
    // d_info: [0..3] pack info, [4..7] count stats, [8..11] table stats
    SPK_CUDA(cudaMemsetAsync(d_info, 0, 12 * sizeof(uint64_t), st));
    SPK_CUDA(cudaMemcpyAsync(d_ascii, h_fasta, nbytes, cudaMemcpyHostToDevice, st));
    int rc = spk_pack_fasta(d_ascii, nbytes, d_packed, d_valid, cap_bases, d_info, d_ws, ws_bytes, stream);
    if (rc) return rc;
    rc = spk_count_table_init(d_table, table_bytes, k, layout, stream);
    if (rc) return rc;
    // the exact number of bases sizes the count grid
    uint64_t info[4];
    SPK_CUDA(cudaMemcpyAsync(info, d_info, sizeof(info), cudaMemcpyDeviceToHost, st));
    SPK_CUDA(cudaStreamSynchronize(st));
    const uint64_t n_bases = info[0];
    rc = spk_count_canonical(d_packed, d_valid, n_bases, k, d_table, table_bytes, layout, d_info + 4, stream);
    if (rc) return rc;
    rc = spk_table_stats(d_table, table_bytes, k, layout, lower_count, d_info + 8, d_block_counts, nullptr,
                         0, stream);
    if (rc) return rc;
    uint64_t all[12];
    SPK_CUDA(cudaMemcpyAsync(all, d_info, sizeof(all), cudaMemcpyDeviceToHost, st));
    SPK_CUDA(cudaStreamSynchronize(st));
    h_out[0] = all[0];   // bases
    h_out[1] = all[1];   // valid bases
    h_out[2] = all[2];   // records
    h_out[3] = all[4];   // valid k-mer occurrences
    h_out[4] = all[8];   // distinct
    h_out[5] = all[9];   // k-mers with count >= lower_count
    h_out[6] = all[10];  // sum of their counts (lengths[i])
    h_out[7] = all[5];   // failed inserts
    if (all[5] != 0) {
        spk_set_error("spk_count_fasta_host: hash table full (%llu inserts failed)", (unsigned long long)all[5]);
        return SPK_EOVERFLOW;
    }
    if (all[9] > cap_out) {
        spk_set_error("spk_count_fasta_host: %llu k-mers to dump but output capacity is %llu",
                      (unsigned long long)all[9], (unsigned long long)cap_out);
        return SPK_ECAP;
    }
    if (d_keys && d_counts && all[9] > 0) {
        rc = spk_table_extract(d_table, table_bytes, k, layout, lower_count, d_block_counts, d_keys, d_counts,
                               cap_out, stream);
        if (rc) return rc;
    }
    return SPK_OK;
}
