"""Drop-in for subphaser/stat_enrich.py (:4-37): summarise an enrichment table by annotation x subgenome.

Input: the TSV `Stats.enrich_ltr` writes (`#id subgenome p_value counts ...`; the reference unpacks exactly the four
columns of its older layout, :11 — any further columns are ignored here).  The annotation of a row is the part of
its id before the first `-`.  Output, one line per annotation in sorted order: the annotation, the number of rows
per subgenome (subgenomes in sorted order), then the column-wise sum of the `counts` vectors of all its rows.

The table is read into arrays once and reduced with two `np.add.at` scatters over (annotation, subgenome) codes;
importing the module does not touch sys.argv (the reference evaluates `sys.argv[1]` when it is imported, :4).
"""
import sys

import numpy as np


def _read(path):
    ids, sgs, rows = [], [], []
    with open(path) as f:
        for line in f:
            if line.startswith("#"):
                continue
            fields = line.split()
            if len(fields) < 4:
                raise ValueError("not enough values to unpack (expected 4, got {})".format(len(fields)))
            ids.append(fields[0].split("-")[0])
            sgs.append(fields[1])
            rows.append(fields[3].split(","))
    counts = np.array(rows, dtype=np.int64) if rows else np.zeros((0, 0), np.int64)
    return ids, sgs, counts


def summarise(ids, sgs, counts):
    """-> (sorted annotations, sorted subgenomes, rows per (annotation, subgenome), summed counts per annotation)"""
    anns, a_code = np.unique(np.array(ids, dtype=object).astype(str), return_inverse=True)
    names, s_code = np.unique(np.array(sgs, dtype=object).astype(str), return_inverse=True)
    num = np.zeros((len(anns), len(names)), dtype=np.int64)
    np.add.at(num, (a_code, s_code), 1)
    tot = np.zeros((len(anns), counts.shape[1] if counts.ndim == 2 else 0), dtype=np.int64)
    np.add.at(tot, a_code, counts)
    return anns, names, num, tot


def main(inTsv=None, outStat=sys.stdout):
    if inTsv is None:
        inTsv = sys.argv[1]
    ids, sgs, counts = _read(inTsv)
    if not ids:
        return
    anns, _, num, tot = summarise(ids, sgs, counts)
    table = np.concatenate([num, tot], axis=1)
    outStat.write("".join(a + "\t" + "\t".join(map(str, r)) + "\n" for a, r in zip(anns.tolist(), table.tolist())))


if __name__ == "__main__":
    main()
