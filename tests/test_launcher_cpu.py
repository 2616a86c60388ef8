"""CPU (build container only): the reference's `__main__` imports and binds to the GPU modules when the
launcher swaps them in — `__main__.py` itself is untouched.  Skipped where /root/reference is absent."""
import sys

import pytest

from oracle import ref_shims


@pytest.mark.skipif(not ref_shims.available(), reason="reference checkout not present")
def test_reference_main_binds_to_gpu_modules():
    ref_shims.install()
    for m in [m for m in sys.modules if m == "subphaser" or m.startswith("subphaser.")]:
        del sys.modules[m]
    from subphaser_b200 import Circos, Cluster, Jellyfish, Seqs, Stats, launcher
    main_mod = launcher.install()
    assert main_mod.run_jellyfish_dumps is Jellyfish.run_jellyfish_dumps
    assert main_mod.JellyfishDumps is Jellyfish.JellyfishDumps
    assert main_mod.Cluster is Cluster.Cluster
    assert main_mod.Stats is Stats
    assert main_mod.Seqs.map_kmer3 is Seqs.map_kmer3
    assert main_mod.Circos.stack_matrix is Circos.stack_matrix
    assert hasattr(main_mod, "Pipeline") and hasattr(main_mod, "makeArgparse")
    for m in [m for m in sys.modules if m == "subphaser" or m.startswith("subphaser.")]:
        del sys.modules[m]
