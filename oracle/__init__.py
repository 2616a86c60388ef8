"""ORACLE — TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference's k-mer hot path, used as the checker by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.  Nothing under
subphaser_b200/ may import this package.

  oracle.kmers    — canonical k-mer counting (C, oracle/kmer_count.c; jellyfish 2.2.10 semantics —
                    PARITY UNPINNED against jellyfish itself, pinned to known-answer vectors) and a
                    string-level brute-force counter
  oracle.restate  — numpy / scipy / sklearn restatements of Jellyfish.py, Cluster.py, Seqs.py,
                    Circos.py, Stats.py arithmetic, each citing the reference lines it follows;
                    pinned to fixtures generated from the reference's own code (tests/golden/)
  oracle.ref_shims — loads the unmodified reference modules (build container only)
"""
