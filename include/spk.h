/*
 * spk.h — C ABI of libspk.so, the sm_100a implementation of SubPhaser's k-mer hot path.
 *
 * The reference (zhangrengang/SubPhaser v1.2.7) is pure Python and has no FFI layer; its boundary for
 * this path is the module surface `subphaser/__main__.py:11-21` imports.  The Python modules in
 * `subphaser_b200/` mirror that surface and call the entry points below through ctypes.  Each entry
 * point cites the reference code (path:line under /root/reference) whose arithmetic it replaces.
 *
 * Conventions
 *   - every function returns 0 on success and a negative SPK_E* code on failure; the message of the
 *     last failure on the calling thread is returned by spk_last_error();
 *   - pointers named d_* are CALLER-OWNED DEVICE pointers (e.g. torch.Tensor.data_ptr()); pointers
 *     named h_* are host pointers; sizes are in elements unless the name says bytes;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); all work is enqueued
 *     on it and nothing synchronises unless stated ("syncs");
 *   - the library allocates no device memory: scratch is passed in (`*_workspace_bytes` tell how much);
 *   - no global state except the thread-local error string and cached device attributes.
 */
#ifndef SPK_H
#define SPK_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPK_OK 0
#define SPK_EINVAL (-1)   /* bad argument */
#define SPK_ECUDA (-2)    /* CUDA runtime error */
#define SPK_ECAP (-3)     /* caller-supplied buffer too small */
#define SPK_EOVERFLOW (-4) /* counter overflow / table full */

const char* spk_last_error(void);
int spk_version(void);
/* Number of SMs of the current device (persistent grids are sized from it). syncs: no. */
int spk_sm_count(void);
/* Number of kernels this library has launched since it was loaded (host-side counter). */
uint64_t spk_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * K1  FASTA bytes -> 2-bit packed bases + validity mask
 * replaces: the `cat/zcat f | jellyfish count ... /dev/stdin` ingest of Jellyfish.py:697 and the
 *           SeqIO.parse + str(rc.seq).upper() of Seqs.py:121-139.
 * Semantics: header lines (from '>' at a line start to the next '\n') and '\n' / '\r' are dropped;
 * every other byte is one base.  A/C/G/T (either case) -> code 0/1/2/3 with valid=1; any other byte ->
 * code 0 with valid=0 (it breaks every k-mer window that covers it, as jellyfish does).  Every header
 * other than one at byte 0 additionally emits ONE invalid separator base so that k-mers never span
 * records.  Base i lives in bits [2*(i%16), 2*(i%16)+2) of d_packed[i/16]; its validity in bit (i%32)
 * of d_valid[i/32].  Both arrays must be zero-padded by the caller up to spk_packed_words(cap)/
 * spk_valid_words(cap) — the kernels write every word of that extent.
 * d_info (uint64[4], device): [0] = number of bases emitted, [1] = number of valid bases,
 *                             [2] = number of records (headers seen), [3] = which kernel produced the result
 *                             (0 regular-layout, 2 general single pass, 1 three-pass; diagnostic).
 * Three kernels, each handing the call to the next on the device when the input is not what it handles: one record
 * with lines of a constant width (positions by arithmetic, layout verified byte by byte); any FASTA whose lines are
 * shorter than 512 bytes (decoupled look-back); anything else (three passes).  SPK_PACK_MODE=single|3pass skips the
 * earlier ones (tests).
 * ---------------------------------------------------------------------------------------------- */
size_t spk_packed_words(uint64_t n_bases); /* uint32 words incl. tile padding */
size_t spk_valid_words(uint64_t n_bases);
size_t spk_pack_workspace_bytes(size_t nbytes);
int spk_pack_fasta(const uint8_t* d_ascii, size_t nbytes, uint32_t* d_packed, uint32_t* d_valid,
                   uint64_t cap_bases, uint64_t* d_info, void* d_ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * K2  canonical k-mer counting of one chromosome
 * replaces: `jellyfish count -m K -s 100000000 --canonical` (Jellyfish.py:697; jellyfish 2.2.10).
 * Every window of k valid bases is counted under min(kmer, revcomp(kmer)) (A<C<G<T, first base most
 * significant), exactly, into a caller-zeroed open-addressed table:
 *   layout 0 (packed): one uint64 per slot, (canon << cbits) | count, cbits = 64-2k; used when
 *                      n_bases < 2^cbits so the count field cannot overflow;
 *   layout 1 (split):  uint64 key[slots] (empty = ~0) followed by uint32 count[slots].
 * spk_count_layout picks the layout from an upper bound of the chromosome length (pass the SAME
 * layout to every call on one table); spk_count_table_bytes the recommended size (load <= 0.7 even
 * if every k-mer is distinct).  The table must be initialised with spk_count_table_init.
 * d_stats (uint64[4], device, accumulated — zero it first): [0] valid k-mer occurrences inserted,
 *   [1] failed inserts (table full; must be 0), [2..3] reserved.
 * 1 <= k <= 32.
 * ---------------------------------------------------------------------------------------------- */
int spk_count_layout(uint64_t n_bases, int k);            /* 0 or 1 for a chromosome of <= n_bases */
size_t spk_count_table_bytes(uint64_t n_bases, int k);     /* recommended table size */
uint64_t spk_count_table_slots(size_t table_bytes, int layout);
int spk_count_table_init(void* d_table, size_t table_bytes, int k, int layout, void* stream);
int spk_count_canonical(const uint32_t* d_packed, const uint32_t* d_valid, uint64_t n_bases, int k,
                        void* d_table, size_t table_bytes, int layout, uint64_t* d_stats,
                        void* stream);

/* ------------------------------------------------------------------------------------------------
 * K2+K3 (v2)  partitioned counting: same semantics and outputs as spk_count_canonical +
 * spk_table_stats + spk_table_extract in ONE asynchronous call, but the chromosome is first split
 * into hash partitions whose tables stay resident in L2 (see subphaser_b200/csrc/spk_pcount.cu).
 * Requires n_bases < 2^32 - 1.  d_ws: spk_pcount_workspace_bytes(n_bases, k) bytes, 256-B aligned.
 * d_keys/d_counts receive every k-mer with count >= lower_count in arbitrary order (like a
 * jellyfish dump); entries beyond `cap` are dropped — compare d_stats[5] with cap.
 * d_stats (uint64[8], overwritten): [0] valid k-mer occurrences, [1] failed inserts (must be 0),
 *   [4] distinct, [5] k-mers >= lower_count, [6] sum of their counts (lengths[i]), [7] sum of all counts.
 * ---------------------------------------------------------------------------------------------- */
size_t spk_pcount_workspace_bytes(uint64_t n_bases, int k);
int spk_pcount_canonical(const uint32_t* d_packed, const uint32_t* d_valid, uint64_t n_bases, int k,
                         uint32_t lower_count, void* d_ws, size_t ws_bytes, uint64_t* d_keys,
                         uint32_t* d_counts, uint64_t cap, uint64_t* d_stats, uint64_t* d_histo,
                         uint32_t histo_len, void* stream);
/* _ex: the same counter with (i) a caller-fixed number of partition bits `pbits` (0 = automatic;
 * otherwise >= spk_pcount_pbits(n_bases, k), <= min(2k, 22)), so that every chromosome of a genome is split
 * by the SAME function of the k-mer, and (ii) the partition index of the dump: the entries of partition p are
 * contiguous, d_pindex (uint32[2 << pbits], optional) receives [2p] = index of the first entry, [2p+1] =
 * number of entries.  This is what lets spk_pmatrix_filter merge the chromosomes partition by partition. */
int spk_pcount_pbits(uint64_t n_bases, int k);
size_t spk_pcount_workspace_bytes_ex(uint64_t n_bases, int k, int pbits);
int spk_pcount_canonical_ex(const uint32_t* d_packed, const uint32_t* d_valid, uint64_t n_bases, int k,
                            uint32_t lower_count, void* d_ws, size_t ws_bytes, uint64_t* d_keys,
                            uint32_t* d_counts, uint64_t cap, uint64_t* d_stats, uint64_t* d_histo,
                            uint32_t histo_len, int pbits, uint32_t* d_pindex, void* stream);

/* ------------------------------------------------------------------------------------------------
 * K3  table scan: `jellyfish dump -c -L lower_count` + the dump parse of Jellyfish.py:90-98
 * spk_table_stats: one pass over the table.  d_out (uint64[4]): [0] distinct k-mers, [1] k-mers with
 *   count >= lower_count, [2] sum of those counts (this is `lengths[i]`, Jellyfish.py:97,449),
 *   [3] sum of all counts.  d_block_counts (uint32[3*spk_table_scan_blocks()+2]) receives the per-block
 *   number of surviving entries; spk_table_extract turns it into offsets and writes the survivors in
 *   slot order (deterministic for a given table) as d_keys[i] (canonical k-mer, 2 bits/base, first
 *   base most significant) and d_counts[i].  d_histo (optional, uint64[histo_len], zeroed by the
 *   caller) accumulates the count histogram like `jellyfish histo -h` (counts >= histo_len-1 go to
 *   the last bin).
 * ---------------------------------------------------------------------------------------------- */
int spk_table_scan_blocks(void);
int spk_table_stats(const void* d_table, size_t table_bytes, int k, int layout, uint32_t lower_count,
                    uint64_t* d_out, uint32_t* d_block_counts, uint64_t* d_histo, uint32_t histo_len,
                    void* stream);
int spk_table_extract(const void* d_table, size_t table_bytes, int k, int layout, uint32_t lower_count,
                      uint32_t* d_block_counts, uint64_t* d_keys, uint32_t* d_counts, uint64_t cap,
                      void* stream);

/* ------------------------------------------------------------------------------------------------
 * K3b/K4  union of the per-chromosome dumps -> count matrix -> differential filter
 * replaces: JellyfishDumps.to_matrix (Jellyfish.py:439-460) and .filter/_filter_kmer
 *           (Jellyfish.py:462-512, 611-648).
 * The union table is an open-addressed map canonical k-mer -> row id.  d_ukeys (uint64[uslots]) must
 * be filled with 0xFF bytes and d_urows (uint32[uslots]) needs no initialisation; *d_nrows (uint32,
 * zeroed) is the row counter.  spk_union_insert adds keys (rows are numbered in arrival order);
 * spk_matrix_fill stores counts[i] into d_matrix[row(keys[i]) * ncol + col] (matrix zeroed by caller;
 * row-major uint32 [nrows x ncol]) and d_row_keys[row] = key.  With nparts > 1 only the keys whose
 * hash falls in partition `part` are inserted / filled: rank r of an N-GPU job builds rows part = r
 * (rows are independent, so the filter shards with no collective).
 *
 * spk_filter_differential evaluates _filter_kmer for every row (outfig is always set by
 * __main__.py:421, so the fold test runs before the frequency gate):
 *   tot = sum(row); for every homoeologous set with >= 2 groups: f_g = sum(count)/sum(length) over the
 *   group's columns (IEEE fp64), sort descending, pass if f[0]/(f[baseline]+1e-20) >= min_fold
 *   (baseline < 0 indexes from the end); include/all < ratio -> reject; then min_freq <= tot <=
 *   max_freq.  d_flags[row]: bit0 = fold test passed, bit1 = kept (fold && frequency gate).
 * Config arrays (device int32): set_off[n_sets+1] -> group range, grp_off[n_groups+1] -> member range,
 * members[] = column indices.  Sets with < 2 groups are skipped as in Jellyfish.py:622-623.
 * by_count != 0 uses the (summed) raw count instead of count/length (Jellyfish.py:632,636).
 * d_counters (uint64[4], zeroed): [0] rows passing the fold test, [1] rows kept.
 * spk_filter_select compacts the kept rows in row order into (d_out_keys[j], d_out_rows[j]);
 * d_scan_ws: uint32[nrows + 2 + nrows/8192], d_scan_ws[nrows] ends up holding the number kept.  The caller may then
 * sort the pairs by key (spk_sort_pairs_u64) for a deterministic row order.  spk_filter_emit writes,
 * for the m rows listed in d_rows: d_out_norm[j*ncol + c] = (double)count / (double)length[c]
 * (Jellyfish.py:648) and d_out_tot[j].
 * ---------------------------------------------------------------------------------------------- */
int spk_union_insert(const uint64_t* d_keys, uint64_t n, uint64_t* d_ukeys, uint32_t* d_urows,
                     uint64_t uslots, uint32_t* d_nrows, uint64_t* d_fail, uint32_t nparts,
                     uint32_t part, void* stream);
int spk_matrix_fill(const uint64_t* d_keys, const uint32_t* d_counts, uint64_t n,
                    const uint64_t* d_ukeys, const uint32_t* d_urows, uint64_t uslots,
                    uint32_t* d_matrix, uint64_t* d_row_keys, int ncol, int col, uint32_t nparts,
                    uint32_t part, void* stream);
int spk_filter_differential(const uint32_t* d_matrix, uint64_t nrows, int ncol,
                            const uint64_t* d_lengths, const int32_t* d_set_off, int n_sets,
                            const int32_t* d_grp_off, int n_groups, const int32_t* d_members, int n_members,
                            double min_fold, int baseline, int by_count, double ratio,
                            double min_freq, double max_freq, uint8_t* d_flags, uint64_t* d_tot,
                            uint64_t* d_counters, void* stream);
int spk_filter_select(const uint64_t* d_row_keys, const uint8_t* d_flags, uint64_t nrows,
                      uint32_t* d_scan_ws, uint64_t* d_out_keys, uint32_t* d_out_rows, uint64_t cap,
                      void* stream);
int spk_filter_emit(const uint32_t* d_matrix, const uint64_t* d_tot, const uint32_t* d_rows,
                    uint64_t m, int ncol, const uint64_t* d_lengths, double* d_out_norm,
                    uint64_t* d_out_tot, void* stream);
/* spk_pmatrix_filter: JellyfishDumps.to_matrix + the first, integer stage of filter (Jellyfish.py:439-512,
 * 611-648) for dumps that carry a partition index with a common `pbits` (spk_pcount_canonical_ex): the union
 * of partition p over the n chromosomes is merged in shared memory and every row goes through the exact
 * integer pre-screen of the differential test (a homoeologous set whose counts are all zero cannot pass the
 * fold test; if too few sets remain to reach `ratio` the row is rejected).  The candidates (~2 % of a real
 * union) are written as a compact count matrix d_out_keys[i], d_out_counts[i*n + c] (arbitrary row order; rows
 * beyond `cap` are dropped) for spk_filter_differential / _select / _emit; every rejected row fails the fold
 * test, so fold-pass and kept counts of the compact matrix are those of the whole union.
 * d_keys / d_counts / d_pindex: device arrays of n device pointers.  nparts/part: only partitions
 * p % nparts == part (multi-GPU row sharding).  n_sets == 0 with cap == 0: count the union rows only.
 * total_entries: sum of the dump sizes (sizes the shared-memory table so an average partition takes one round).
 * d_counters (uint64[8], overwritten): [0] union rows (= len(d_mat)), [2] candidate rows, [3] table overflows
 * (must be 0; otherwise use the plain path), [4] rows written. */
int spk_pmatrix_filter(const uint64_t* const* d_keys, const uint32_t* const* d_counts,
                       const uint32_t* const* d_pindex, int n, int pbits, uint32_t nparts, uint32_t part,
                       const uint64_t* d_lengths, const int32_t* d_set_off, int n_sets, const int32_t* d_grp_off,
                       int n_groups, const int32_t* d_members, int n_members, double min_fold, int baseline,
                       int by_count, double ratio, double min_freq, double max_freq, uint64_t* d_out_keys,
                       uint32_t* d_out_counts, uint64_t cap, uint64_t total_entries, uint64_t* d_counters,
                       void* stream);
/* spk_dump_regroup: move the contiguous run of every hash partition of a dump (d_pindex, uint32[2 << pbits])
 * to d_new_start[p] (uint32[1 << pbits]) in d_out_keys / d_out_counts — the multi-GPU exchange regroups a dump by
 * destination rank (partition class p mod world) before the all-to-all. */
int spk_dump_regroup(const uint64_t* d_keys, const uint32_t* d_counts, const uint32_t* d_pindex,
                     const uint32_t* d_new_start, int pbits, uint64_t* d_out_keys, uint32_t* d_out_counts,
                     void* stream);
/* spk_dump_scatter_peers: the exchange step without a collective — write every hash-partition run of a dump
 * (d_pindex) directly into the receive buffers of the rank that merges its class (p mod world) over peer-mapped
 * memory (NVLink P2P stores; replaces the reference's single-process dict merge of Jellyfish.py:447-458 across
 * GPUs).  d_peer_keys / d_peer_counts / d_peer_psize: device arrays of `world` pointers (one per rank, symmetric
 * allocations).  The run of partition p lands at region_off + (exclusive sum of the earlier partitions of its class)
 * in the destination's key / count buffers and its length at psize_off + p / world in the destination's size table;
 * a class that would exceed region_cap entries is not written and *d_overflow is incremented (caller falls back to
 * the collective exchange).  d_dst_off: uint32[1 << pbits] scratch; d_class_tot: uint64[world] out, entries per
 * class.  syncs: no. */
int spk_dump_scatter_peers(const uint64_t* d_keys, const uint32_t* d_counts, const uint32_t* d_pindex, int pbits,
                           uint32_t world, uint64_t region_off, uint64_t region_cap, uint64_t psize_off,
                           uint64_t* const* d_peer_keys, uint32_t* const* d_peer_counts,
                           uint32_t* const* d_peer_psize, uint32_t* d_dst_off, uint64_t* d_class_tot,
                           uint64_t* d_overflow, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Histogram of the totals of the fold-passing k-mers (`.kmer_freq.pdf`, Jellyfish.py:499-511,650-666) without
 * materialising them on the host.  d_tot / d_flags: outputs of spk_filter_differential (bit 0 of a flag = the row
 * passed the fold test).  spk_tot_minmax: d_out uint64[3] = count, min, max.  spk_tot_histogram: numpy.histogram
 * semantics for `nbins` uniform bins over [mn, mx].  spk_tot_select_pass: one 16-bit pass of a radix select (counts of
 * bits [shift, shift+16) among the values whose higher bits equal `prefix`) — the caller walks the passes to the
 * order statistics np.percentile needs.  syncs: no. */
int spk_tot_minmax(const uint64_t* d_tot, const uint8_t* d_flags, uint64_t n, uint64_t* d_out, void* stream);
int spk_tot_histogram(const uint64_t* d_tot, const uint8_t* d_flags, uint64_t n, double mn, double mx, uint32_t nbins,
                      uint64_t* d_hist, void* stream);
int spk_tot_select_pass(const uint64_t* d_tot, const uint8_t* d_flags, uint64_t n, int shift, uint64_t prefix,
                        uint64_t* d_hist65536, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Genome ingest (Seqs.split_genomes, Seqs.py:27-71, without BioPython): the genome file's bytes are on the device.
 * spk_fasta_record_starts: positions of every record start ('>' at byte 0 or after '\n') appended, unordered, to
 *   d_pos (uint64[cap]); *d_count = number of records (may exceed cap: call again with a larger array).
 * spk_fasta_wrap_check: is the body [beg, end) of a record laid out as BioPython writes it (lines of `width` bytes
 *   ending in '\n')?  d_out (uint64[2]): [0] bit 0 = irregular line breaks, bit 1 = bytes that need rewriting
 *   ('\r', blanks, '>'); [1] = number of '\n' bytes.  A regular body is written to `<id>.fasta` verbatim.
 * syncs: no. */
int spk_fasta_record_starts(const uint8_t* d_ascii, uint64_t nbytes, uint64_t* d_pos, uint64_t cap, uint64_t* d_count,
                            void* stream);
int spk_fasta_wrap_check(const uint8_t* d_ascii, uint64_t beg, uint64_t end, uint32_t width, uint64_t* d_out,
                         void* stream);


/* ------------------------------------------------------------------------------------------------
 * Text wire formats on the device.  spk_format_rows writes the rows of `.kmer.mat` (kind 0:
 * `KMER\tv1\t...\tvn\n`, JellyfishDumps.write_matrix, Jellyfish.py:515-520) or of `.sig.kmer-subgenome.tsv`
 * (kind 1: `KMER\tLABEL\tp\tm1,...,mn\n`, Cluster.output_kmers, Cluster.py:158-172) exactly as Python's str()
 * prints them (floats: shortest round-trip repr).  d_rows (optional, uint32[n_rows]): the rows to write, else all M.
 * Call it twice: with d_row_len (uint32[n_rows] out) and d_out = NULL to get the row lengths, then — after an
 * exclusive scan into d_row_off (uint64[n_rows]) — with d_row_off and d_out to write the bytes.
 * kind 1: d_label int32[M] (index into d_label_text, 16 bytes per label, lengths in d_label_len), d_pval [M]. */
int spk_format_rows(const uint64_t* d_keys, const double* d_vals, uint64_t M, int n, int k, int kind,
                    const uint32_t* d_rows, uint64_t n_rows, const int32_t* d_label, const char* d_label_text,
                    const int32_t* d_label_len, const double* d_pval, const uint64_t* d_row_off,
                    uint32_t* d_row_len, char* d_out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Sorting helper (stable LSD radix sort of uint64 keys with a uint32 payload, ascending).
 * Used to give the matrix a deterministic (sorted) row order and by the BH step.
 * d_keys_tmp / d_vals_tmp: scratch of n elements each; the result ends in d_keys / d_vals.
 * key_bits: number of significant low bits (rounded up to a multiple of 4).
 * ---------------------------------------------------------------------------------------------- */
size_t spk_sort_workspace_bytes(uint64_t n);
int spk_sort_pairs_u64(uint64_t* d_keys, uint32_t* d_vals, uint64_t* d_keys_tmp,
                       uint32_t* d_vals_tmp, uint64_t n, int key_bits, void* d_ws, size_t ws_bytes,
                       void* stream);

/* ------------------------------------------------------------------------------------------------
 * K9  per-position lookup of subgenome-specific k-mers, counted into bins
 * replaces: Seqs.map_kmer3 / map_kmer_each4 / _get_kmer (Seqs.py:74-119, 209-244).
 * The specific-k-mer table maps CANONICAL k-mer -> subgenome index (the reference's dict holds each
 * k-mer and its reverse complement with the same value, Cluster.py:174-175, which is the same map).
 * spk_sig_table_build fills an open-addressed table d_skeys (uint64[sslots], pre-filled with 0xFF) /
 * d_svals (uint8[sslots]) and, optionally, a one-hash membership bitmap d_filter (uint32[filter_bits/32],
 * zeroed, filter_bits a power of two ~16x the key count) that lets the ~97 % of positions that are not
 * specific k-mers finish after a single 4-byte load.  With pack_vals != 0 (k <= 28) the subgenome id is
 * stored in the top byte of the key slot and d_svals is not touched.  spk_map_bins then looks up the k-mer starting at every position i of the
 * packed chromosome and, on a hit with value sg, increments d_line_counts[line(i) * S + sg] where
 *   line(i) = i / bin_size + (chunk_size ? (i + k - 1) / chunk_size : 0)
 * i.e. one counter row per (bin, 10-Mb chunk) pair, reproducing the duplicate border lines of
 * Seqs.py:131-137,229-236 (chunk_size = 0: no chunking, `chunk=False`).  d_line_counts is uint32
 * [n_lines x S], zeroed by the caller, n_lines = spk_map_num_lines(...).
 * d_hit_flags (optional, uint8[sslots], zeroed) is set for every table slot that was hit (the
 * reference's "mapped kmers" set, Seqs.py:109,227); d_nhits (uint64[1], zeroed) counts hits.
 * ---------------------------------------------------------------------------------------------- */
 /* spk_stack_windows: Circos.stack_matrix / _bed_density(stack=True) (Circos.py:709-742,831-842):
  * d_out[d_line_window[l] * S + c] += d_line_counts[l * S + c]  (d_out int64 [W x S], zeroed). */
int spk_stack_windows(const int64_t* d_line_counts, const uint32_t* d_line_window, uint64_t n_lines,
                      int S, int64_t* d_out, void* stream);
/* spk_stack_lines: the same stacking for the line-count array of ONE chromosome straight from
 * spk_map_bins (no text round trip): line l is assigned to window (its bin start) / window_size. */
int spk_stack_lines(const uint32_t* d_line_counts, uint64_t n_lines, int S, int k, uint64_t bin_size,
                    uint64_t chunk_size, uint64_t window_size, uint64_t n_bases, int64_t* d_out,
                    uint64_t n_windows, void* stream);
int spk_sig_table_build(const uint64_t* d_keys, const uint8_t* d_vals, uint64_t n, uint64_t* d_skeys,
                        uint8_t* d_svals, uint64_t sslots, uint32_t* d_filter, uint64_t filter_bits,
                        int pack_vals, uint64_t* d_fail, void* stream);
uint64_t spk_map_num_lines(uint64_t n_bases, int k, uint64_t bin_size, uint64_t chunk_size);
int spk_map_bins(const uint32_t* d_packed, const uint32_t* d_valid, uint64_t n_bases, int k,
                 const uint64_t* d_skeys, const uint8_t* d_svals, uint64_t sslots, int S,
                 const uint32_t* d_filter, uint64_t filter_bits, int pack_vals, uint64_t bin_size,
                 uint64_t chunk_size, uint32_t* d_line_counts, uint64_t n_lines,
                 uint8_t* d_hit_flags, uint64_t* d_nhits, void* stream);
/* Bucketed quotient table — the shipped layout of the same map (Seqs.py:209-244 lookups), sized to stay
 * L2-resident and to answer a position with ONE 16- or 32-byte load.  f = bijective mixer on the 2k-bit
 * canonical word; bucket = top bucket_bits of f(u); a slot holds (remainder << sgbits) | subgenome id
 * (uint16 or uint32 slots, all-ones = empty); a bucket = 7 entry slots + 1 overflow marker.  Keys whose
 * bucket is full go to the open-addressed stash d_skeys/d_svals (as in spk_sig_table_build, pre-filled
 * with 0xFF), probed only for marked buckets; membership is exact.
 * spk_qtable_plan picks slot_bits (16|32) and bucket_bits for n_keys, k, S (SPK_EINVAL: no such layout,
 * use spk_sig_table_build / spk_map_bins).  d_buckets: (8 << bucket_bits) slots pre-filled with 0xFF.
 * spk_map_bins_q: same contract as spk_map_bins; d_hit_flags (optional) is uint8[(8 << bucket_bits) + sslots].
 * Multi-record FASTA (Seqs.py:121-153 map every record on its own coordinates): d_rec_start (uint64[n_rec+1],
 * packed position where each record starts, separator bases included, last = n_bases) and d_rec_line0
 * (uint64[n_rec], first counter row of each record; a record of L bases owns spk_map_num_lines(L, ...) rows);
 * a hit at packed position i of record r then counts in row rec_line0[r] + line(i - rec_start[r]).
 * NULL: the whole input is one record. */
int spk_qtable_plan(uint64_t n_keys, int k, int S, int* slot_bits, int* bucket_bits);
int spk_qtable_build(const uint64_t* d_keys, const uint8_t* d_vals, uint64_t n, int k, int S, void* d_buckets,
                     int slot_bits, int bucket_bits, uint64_t* d_skeys, uint8_t* d_svals, uint64_t sslots,
                     int pack_vals, uint64_t* d_fail, void* stream);
int spk_map_bins_q(const uint32_t* d_packed, const uint32_t* d_valid, uint64_t n_bases, int k,
                   const void* d_buckets, int slot_bits, int bucket_bits, const uint64_t* d_skeys,
                   const uint8_t* d_svals, uint64_t sslots, int pack_vals, int S, uint64_t bin_size,
                   uint64_t chunk_size, uint32_t* d_line_counts, uint64_t n_lines, uint8_t* d_hit_flags,
                   uint64_t* d_nhits, const uint64_t* d_rec_start, const uint64_t* d_rec_line0, uint32_t n_rec,
                   void* stream);

/* ------------------------------------------------------------------------------------------------
 * K10  per-window Fisher exact test (right tail) + Benjamini-Hochberg
 * replaces: Stats.fisher_test (Stats.py:14-31; fisher 0.1.9 `pvalue(...).right_tail`) and
 *           Stats.correct_pvals (Stats.py:11-12; statsmodels multipletests fdr_bh).
 * For row r, column i: x11=c[r,i]; x12=sum(row)-x11; x21=tot[i]-x11; x22=sum(tot)-x21-x12 (as coded,
 * Stats.py:20-23); x21,x22 clamped to 214748364; p = P(X >= x11), X ~ Hypergeom(N=x11+x12+x21+x22,
 * K=x11+x21, n=x11+x12).  d_counts int64 [W x S] row-major, d_totals int64 [S], d_pvals fp64 [W x S].
 * spk_enrich_rows applies Pvalues.get_enriched + _enrich (Stats.py:150-192): per row the stable
 * arg-min, significance, ratios; outputs d_idx int32[W], d_sig uint8[W], d_ratios fp64[W x S],
 * d_pmin fp64[W] (the row's smallest p-value, the one BH corrects).  spk_colsum_i64: the genome-wide
 * column totals `arr.sum(axis=0)` (Stats.py:145).
 * spk_bh_adjust: q-values of n p-values (needs ws from spk_bh_workspace_bytes).
 * ---------------------------------------------------------------------------------------------- */
int spk_fisher_right_tail(const int64_t* d_counts, const int64_t* d_totals, uint64_t W, int S,
                          double* d_pvals, void* stream);
int spk_colsum_i64(const int64_t* d_counts, uint64_t W, int S, int64_t* d_totals, void* stream);
int spk_enrich_rows(const int64_t* d_counts, const int64_t* d_totals, const double* d_pvals,
                    uint64_t W, int S, double max_pval, double cutoff, double min_ratio,
                    int32_t* d_idx, uint8_t* d_sig, double* d_ratios, double* d_pmin, void* stream);
/* Self-check: total mass of Hypergeom(N,K,n) for `count` triples d_NKn[3*i..] summed with the kernel's
 * own point-mass routine and recurrences (must be 1 within ~1e-13). */
int spk_debug_hypergeom_mass(const int64_t* d_NKn, uint64_t count, double* d_out, void* stream);
size_t spk_bh_workspace_bytes(uint64_t n);
int spk_bh_adjust(const double* d_p, double* d_q, uint64_t n, void* d_ws, size_t ws_bytes,
                  void* stream);

/* ------------------------------------------------------------------------------------------------
 * K5-K8  statistics over the chromosome x k-mer matrix (all fp64)
 * Xraw is the M x n row-major matrix of `.kmer.mat` (row = k-mer, column = chromosome).
 *
 * spk_zscore_rows: Z[m, c] = (X[m,c] - mean_m) / std_m with population std over the n chromosomes —
 *   Cluster.normalize_data on data.T, axis=0 (Cluster.py:25-26,76-80).
 * spk_gram: G[n x n] = sum_m Z[m,:]^T Z[m,:] (optionally only over rows listed in d_idx) — all point-
 *   to-point geometry K-Means and PCA need for n points in M dimensions.
 * spk_kmeans_gram: Lloyd K-Means on n points given G (k-means++ seeding from `seed`, n_init restarts,
 *   max_iter; best inertia wins) -> labels int32[n], inertia.  Replaces sklearn KMeans(n_clusters)
 *   as called by Cluster.fit (Cluster.py:114-118).
 * spk_gram_batched + spk_kmeans_gram with R > 1: the bootstrap loop of Cluster.bootstrap
 *   (Cluster.py:82-112): replicate r clusters the B resampled k-mers d_idx[r*B..] (indices supplied
 *   by the caller, drawn with replacement).
 * spk_cluster_scores: adjusted Rand index and V-measure of each replicate vs d_ref_labels
 *   (Cluster.py:97-100; sklearn.metrics).
 * spk_centroids: C[s, m] = mean of Z[m, c] over c with label s  (sklearn cluster_centers_).
 * spk_ttest_groups: Cluster._output_kmers (Cluster.py:178-194) with scipy.stats.ttest_ind: group the
 *   n values of each raw row by subgenome label, order groups by descending mean (stable), Student
 *   t-test (pooled variance, two-sided) of the top two; outputs best group, p-value, group means.
 * spk_pca_gram: eigen-decomposition of G (Jacobi) -> eigenvalues (descending) and scores U*sqrt(l)
 *   [n x ncomp], explained variance ratio; equals sklearn PCA(svd_solver='full') up to sign
 *   (Cluster.py:48-51).
 * ---------------------------------------------------------------------------------------------- */
int spk_zscore_rows(const double* d_X, uint64_t M, int n, double* d_Z, void* stream);
size_t spk_gram_workspace_bytes(int n);
int spk_gram(const double* d_Z, uint64_t M, int n, const uint32_t* d_idx, uint64_t n_idx,
             double* d_G, void* d_ws, size_t ws_bytes, void* stream);
/* R Gram matrices, replicate r over the B rows d_idx[r*B .. r*B+B): d_G is fp64 [R x n x n] */
int spk_gram_batched(const double* d_Z, uint64_t M, int n, const uint32_t* d_idx, int R, int B,
                     double* d_G, void* stream);
/* R independent K-Means problems on Gram matrices d_G [R x n x n].  d_order (int32[n], optional):
 * chromosome indices sorted by name — labels are renumbered by first appearance in that order
 * (Cluster.sort_subgenomes, Cluster.py:119-126).  d_labels int32 [R x n], d_inertia fp64 [R]. */
size_t spk_kmeans_workspace_bytes(int R);
int spk_kmeans_gram(const double* d_G, int R, int n, int S, int n_init, int max_iter, uint64_t seed,
                    const int32_t* d_order, int32_t* d_labels, double* d_inertia, void* d_ws,
                    size_t ws_bytes, void* stream);
/* the same for replicates r0 .. r0+R-1 of a larger batch (the random stream of a replicate is a function of its
 * global number): ranks that split the bootstrap replicates of Cluster.py:82-112 reproduce the one-rank result */
int spk_kmeans_gram_at(const double* d_G, int R, int r0, int n, int S, int n_init, int max_iter, uint64_t seed,
                       const int32_t* d_order, int32_t* d_labels, double* d_inertia, void* d_ws,
                       size_t ws_bytes, void* stream);
int spk_cluster_scores(const int32_t* d_ref_labels, const int32_t* d_labels, int R, int n,
                       double* d_ari, double* d_vmeasure, void* stream);
int spk_centroids(const double* d_Z, uint64_t M, int n, const int32_t* d_labels, int S,
                  double* d_C, void* stream);
int spk_ttest_groups(const double* d_X, uint64_t M, int n, const int32_t* d_col_group, int S,
                     int32_t* d_best, double* d_pval, double* d_means, void* stream);
/* spk_ranktest_groups: the other `test_method` choices of Cluster.output_kmers (Cluster.py:160,191): method 1 =
 * scipy.stats.kruskal, 2 = mannwhitneyu, 3 = wilcoxon (default arguments; wilcoxon with the mode rules of the pinned
 * scipy 1.7.1) on the two groups with the largest means.  d_pair_off (int32[S*S], index g_top * S + g_second): offset
 * of that pair's exact null distribution in d_tables (built by the caller with integer arithmetic) or -1.
 * d_flags (uint32, caller zeroes): bit 0 wilcoxon on groups of different size, bit 1 kruskal on identical values —
 * scipy raises ValueError for those; the rows get NaN. */
int spk_ranktest_groups(const double* d_X, uint64_t M, int n, const int32_t* d_col_group, int S, int method,
                        const int32_t* d_pair_off, const double* d_tables, int32_t* d_best, double* d_pval,
                        double* d_means, uint32_t* d_flags, void* stream);
size_t spk_pca_workspace_bytes(int n);
int spk_pca_gram(const double* d_G, int n, int ncomp, double* d_eigvals, double* d_scores,
                 double* d_ratio, void* d_ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Host-buffer entry point (what a file-level caller binds): FASTA bytes in HOST memory -> H2D copy ->
 * K1 -> K2 -> K3 on `stream`.  d_* are device scratch/outputs as above; d_ascii must hold nbytes.
 * d_info: uint64[12] device scratch.  d_block_counts as for spk_table_stats.  Sizes derive from
 * cap_bases (>= nbytes): packed/valid words, workspace, table bytes, layout.
 * h_out (uint64[8], host): [0] n_bases, [1] n_valid_bases, [2] n_records, [3] valid k-mers,
 *   [4] distinct, [5] k-mers >= lower_count, [6] sum of their counts (lengths[i]), [7] failed inserts.
 * syncs: yes (it returns the numbers).  The extracted (key,count) list is left in d_keys/d_counts.
 * ---------------------------------------------------------------------------------------------- */
int spk_count_fasta_host(const uint8_t* h_fasta, size_t nbytes, int k, uint32_t lower_count,
                         uint8_t* d_ascii, uint32_t* d_packed, uint32_t* d_valid, uint64_t cap_bases,
                         void* d_ws, size_t ws_bytes, void* d_table, size_t table_bytes,
                         uint32_t* d_block_counts, uint64_t* d_keys, uint32_t* d_counts,
                         uint64_t cap_out, uint64_t* d_info, uint64_t* h_out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Synthetic genome generator (TEST/BENCH INFRASTRUCTURE, not part of the product path): writes the
 * FASTA bytes of one chromosome directly in device memory.  See subphaser_b200/synth.py.
 * ---------------------------------------------------------------------------------------------- */
int spk_synth_fasta(uint8_t* d_out, uint64_t nbytes, uint64_t header_len, uint64_t n_bases,
                    int line_width, const uint64_t* d_seg_start, const int64_t* d_seg_src,
                    const uint32_t* d_seg_seed, uint64_t n_segs, const uint8_t* d_library,
                    uint64_t lib_len, const uint64_t* d_nrun_start, const uint64_t* d_nrun_end,
                    uint64_t n_nruns, double div, double soft_frac, uint64_t seed, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SPK_H */
