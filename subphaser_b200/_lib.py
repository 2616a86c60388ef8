"""ctypes binding of libspk.so (include/spk.h).  There is NO CPU fallback: if the library is missing
or no CUDA device is visible the product path raises."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libspk.so")

c_p = ctypes.c_void_p
c_u64 = ctypes.c_uint64
c_u32 = ctypes.c_uint32
c_i = ctypes.c_int
c_sz = ctypes.c_size_t
c_d = ctypes.c_double

# name -> (restype, argtypes); must list every symbol include/spk.h declares
SIGNATURES = {
    "spk_last_error": (ctypes.c_char_p, []),
    "spk_version": (c_i, []),
    "spk_sm_count": (c_i, []),
    "spk_launch_count": (c_u64, []),
    "spk_packed_words": (c_sz, [c_u64]),
    "spk_valid_words": (c_sz, [c_u64]),
    "spk_pack_workspace_bytes": (c_sz, [c_sz]),
    "spk_pack_fasta": (c_i, [c_p, c_sz, c_p, c_p, c_u64, c_p, c_p, c_sz, c_p]),
    "spk_count_layout": (c_i, [c_u64, c_i]),
    "spk_count_table_bytes": (c_sz, [c_u64, c_i]),
    "spk_count_table_slots": (c_u64, [c_sz, c_i]),
    "spk_count_table_init": (c_i, [c_p, c_sz, c_i, c_i, c_p]),
    "spk_count_canonical": (c_i, [c_p, c_p, c_u64, c_i, c_p, c_sz, c_i, c_p, c_p]),
    "spk_pcount_workspace_bytes": (c_sz, [c_u64, c_i]),
    "spk_pcount_canonical": (c_i, [c_p, c_p, c_u64, c_i, c_u32, c_p, c_sz, c_p, c_p, c_u64, c_p, c_p, c_u32, c_p]),
    "spk_pcount_pbits": (c_i, [c_u64, c_i]),
    "spk_pcount_workspace_bytes_ex": (c_sz, [c_u64, c_i, c_i]),
    "spk_pcount_canonical_ex": (c_i, [c_p, c_p, c_u64, c_i, c_u32, c_p, c_sz, c_p, c_p, c_u64, c_p, c_p, c_u32,
                                      c_i, c_p, c_p]),
    "spk_pmatrix_filter": (c_i, [c_p, c_p, c_p, c_i, c_i, c_u32, c_u32, c_p, c_p, c_i, c_p, c_i, c_p, c_i, c_d, c_i, c_i,
                                 c_d, c_d, c_d, c_p, c_p, c_u64, c_u64, c_p, c_p]),
    "spk_dump_regroup": (c_i, [c_p, c_p, c_p, c_p, c_i, c_p, c_p, c_p]),
    "spk_dump_scatter_peers": (c_i, [c_p, c_p, c_p, c_i, c_u32, c_u64, c_u64, c_u64, c_p, c_p, c_p, c_p, c_p, c_p, c_p]),
    "spk_table_scan_blocks": (c_i, []),
    "spk_table_stats": (c_i, [c_p, c_sz, c_i, c_i, c_u32, c_p, c_p, c_p, c_u32, c_p]),
    "spk_table_extract": (c_i, [c_p, c_sz, c_i, c_i, c_u32, c_p, c_p, c_p, c_u64, c_p]),
    "spk_union_insert": (c_i, [c_p, c_u64, c_p, c_p, c_u64, c_p, c_p, c_u32, c_u32, c_p]),
    "spk_matrix_fill": (c_i, [c_p, c_p, c_u64, c_p, c_p, c_u64, c_p, c_p, c_i, c_i, c_u32, c_u32, c_p]),
    "spk_filter_differential": (c_i, [c_p, c_u64, c_i, c_p, c_p, c_i, c_p, c_i, c_p, c_i, c_d, c_i, c_i,
                                      c_d, c_d, c_d, c_p, c_p, c_p, c_p]),
    "spk_filter_select": (c_i, [c_p, c_p, c_u64, c_p, c_p, c_p, c_u64, c_p]),
    "spk_filter_emit": (c_i, [c_p, c_p, c_p, c_u64, c_i, c_p, c_p, c_p, c_p]),
    "spk_tot_minmax": (c_i, [c_p, c_p, c_u64, c_p, c_p]),
    "spk_tot_histogram": (c_i, [c_p, c_p, c_u64, c_d, c_d, c_u32, c_p, c_p]),
    "spk_tot_select_pass": (c_i, [c_p, c_p, c_u64, c_i, c_u64, c_p, c_p]),
    "spk_fasta_record_starts": (c_i, [c_p, c_u64, c_p, c_u64, c_p, c_p]),
    "spk_fasta_wrap_check": (c_i, [c_p, c_u64, c_u64, c_u32, c_p, c_p]),
    "spk_format_rows": (c_i, [c_p, c_p, c_u64, c_i, c_i, c_i, c_p, c_u64, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p]),
    "spk_sort_workspace_bytes": (c_sz, [c_u64]),
    "spk_sort_pairs_u64": (c_i, [c_p, c_p, c_p, c_p, c_u64, c_i, c_p, c_sz, c_p]),
    "spk_stack_windows": (c_i, [c_p, c_p, c_u64, c_i, c_p, c_p]),
    "spk_stack_lines": (c_i, [c_p, c_u64, c_i, c_i, c_u64, c_u64, c_u64, c_u64, c_p, c_u64, c_p]),
    "spk_sig_table_build": (c_i, [c_p, c_p, c_u64, c_p, c_p, c_u64, c_p, c_u64, c_i, c_p, c_p]),
    "spk_map_num_lines": (c_u64, [c_u64, c_i, c_u64, c_u64]),
    "spk_map_bins": (c_i, [c_p, c_p, c_u64, c_i, c_p, c_p, c_u64, c_i, c_p, c_u64, c_i, c_u64, c_u64, c_p,
                           c_u64, c_p, c_p, c_p]),
    "spk_qtable_plan": (c_i, [c_u64, c_i, c_i, c_p, c_p]),
    "spk_qtable_build": (c_i, [c_p, c_p, c_u64, c_i, c_i, c_p, c_i, c_i, c_p, c_p, c_u64, c_i, c_p, c_p]),
    "spk_map_bins_q": (c_i, [c_p, c_p, c_u64, c_i, c_p, c_i, c_i, c_p, c_p, c_u64, c_i, c_i, c_u64, c_u64,
                             c_p, c_u64, c_p, c_p, c_p, c_p, c_u32, c_p]),
    "spk_fisher_right_tail": (c_i, [c_p, c_p, c_u64, c_i, c_p, c_p]),
    "spk_colsum_i64": (c_i, [c_p, c_u64, c_i, c_p, c_p]),
    "spk_enrich_rows": (c_i, [c_p, c_p, c_p, c_u64, c_i, c_d, c_d, c_d, c_p, c_p, c_p, c_p, c_p]),
    "spk_debug_hypergeom_mass": (c_i, [c_p, c_u64, c_p, c_p]),
    "spk_bh_workspace_bytes": (c_sz, [c_u64]),
    "spk_bh_adjust": (c_i, [c_p, c_p, c_u64, c_p, c_sz, c_p]),
    "spk_zscore_rows": (c_i, [c_p, c_u64, c_i, c_p, c_p]),
    "spk_gram_workspace_bytes": (c_sz, [c_i]),
    "spk_gram": (c_i, [c_p, c_u64, c_i, c_p, c_u64, c_p, c_p, c_sz, c_p]),
    "spk_gram_batched": (c_i, [c_p, c_u64, c_i, c_p, c_i, c_i, c_p, c_p]),
    "spk_kmeans_workspace_bytes": (c_sz, [c_i]),
    "spk_kmeans_gram": (c_i, [c_p, c_i, c_i, c_i, c_i, c_i, c_u64, c_p, c_p, c_p, c_p, c_sz, c_p]),
    "spk_kmeans_gram_at": (c_i, [c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_u64, c_p, c_p, c_p, c_p, c_sz, c_p]),
    "spk_cluster_scores": (c_i, [c_p, c_p, c_i, c_i, c_p, c_p, c_p]),
    "spk_centroids": (c_i, [c_p, c_u64, c_i, c_p, c_i, c_p, c_p]),
    "spk_ttest_groups": (c_i, [c_p, c_u64, c_i, c_p, c_i, c_p, c_p, c_p, c_p]),
    "spk_ranktest_groups": (c_i, [c_p, c_u64, c_i, c_p, c_i, c_i, c_p, c_p, c_p, c_p, c_p, c_p, c_p]),
    "spk_pca_workspace_bytes": (c_sz, [c_i]),
    "spk_pca_gram": (c_i, [c_p, c_i, c_i, c_p, c_p, c_p, c_p, c_sz, c_p]),
    "spk_count_fasta_host": (c_i, [c_p, c_sz, c_i, c_u32, c_p, c_p, c_p, c_u64, c_p, c_sz, c_p, c_sz,
                                   c_p, c_p, c_p, c_u64, c_p, c_p, c_p]),
    "spk_synth_fasta": (c_i, [c_p, c_u64, c_u64, c_u64, c_i, c_p, c_p, c_p, c_u64, c_p, c_u64, c_p,
                              c_p, c_u64, c_d, c_d, c_u64, c_p]),
}

# functions returning an SPK_* status code
_STATUS = {n for n, (r, _) in SIGNATURES.items() if r is c_i} - {
    "spk_version", "spk_sm_count", "spk_count_layout", "spk_table_scan_blocks", "spk_qtable_plan", "spk_pcount_pbits"}


class SpkError(RuntimeError):
    pass


_lib = None


def load():
    """Load libspk.so (once).  Raises if it has not been built — no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SpkError(
                "libspk.so is missing (%s). Build it with `python -m subphaser_b200.build`; "
                "there is no CPU fallback." % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def call(name, *args):
    """Call a status-returning entry point; raise SpkError with spk_last_error() on failure."""
    lib = load()
    rc = getattr(lib, name)(*args)
    if name in _STATUS and rc != 0:
        msg = lib.spk_last_error().decode(errors="replace")
        code = {-1: ValueError, -3: SpkError, -4: OverflowError}.get(rc, SpkError)
        raise code("%s failed (%d): %s" % (name, rc, msg))
    return rc
