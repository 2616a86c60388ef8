"""Drop-in for subphaser/stat_enrich.py (:4-37): aggregate an enrich TSV by annotation prefix x
subgenome.  Host-only, tiny.  Unlike the reference, importing this module does not evaluate
sys.argv[1] at definition time, and both the 4-column (old) and 6-column (current enrich_ltr,
Stats.py:67) layouts are accepted."""
import sys

import numpy as np


def main(inTsv=None, outStat=sys.stdout):
    if inTsv is None:
        inTsv = sys.argv[1]
    d_count = {}
    ids, sgs = set([]), set([])
    for line in open(inTsv):
        if line.startswith("#"):
            continue
        temp = line.strip().split()
        id, subgenome, p_value, counts = temp[:4]
        ann = id.split("-")[0]
        counts = np.array(list(map(int, counts.split(","))))
        key = (ann, subgenome)
        if key not in d_count:
            d_count[key] = [1, counts]
        else:
            d_count[key][0] += 1
            d_count[key][1] = d_count[key][1] + counts
        ids.add(key[0])
        sgs.add(key[1])
    for ann in sorted(ids):
        num = []
        count = None
        for i, sg in enumerate(sorted(sgs)):
            key = (ann, sg)
            if key in d_count:
                _num, _count = d_count[key]
            else:
                _num, _count = 0, np.array([0] * len(sgs))
            num += [_num]
            count = _count if i == 0 else count + _count
        line = [ann] + num + list(count)
        outStat.write("\t".join(map(str, line)) + "\n")


if __name__ == "__main__":
    main()
