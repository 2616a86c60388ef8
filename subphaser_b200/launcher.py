"""Run the UNMODIFIED SubPhaser pipeline (`subphaser.__main__.main`) on top of the GPU modules.

    python -m subphaser_b200.launcher -i genome.fa -c sg.config [SubPhaser options ...]

Requires the reference package `subphaser` to be importable; it is not modified — the hot-path modules
are swapped in `sys.modules` before `subphaser.__main__` is imported."""
import importlib
import sys


def install():
    from . import Circos as GpuCircos
    from . import Cluster, Data, Jellyfish, Seqs, Stats
    for name, mod in (("Jellyfish", Jellyfish), ("Cluster", Cluster), ("Stats", Stats), ("Data", Data)):
        sys.modules["subphaser." + name] = mod
    ref_seqs = importlib.import_module("subphaser.Seqs")
    ref_circos = importlib.import_module("subphaser.Circos")
    ref_seqs.map_kmer3 = Seqs.map_kmer3                 # Seqs.py:74
    ref_circos.stack_matrix = GpuCircos.stack_matrix    # Circos.py:831
    pkg = importlib.import_module("subphaser")
    for name, mod in (("Jellyfish", Jellyfish), ("Cluster", Cluster), ("Stats", Stats), ("Data", Data)):
        setattr(pkg, name, mod)
    return importlib.import_module("subphaser.__main__")


def main():
    install().main()


if __name__ == "__main__":
    main()
