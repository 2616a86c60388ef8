"""Aggregate an `ncu --page source --csv` (SASS) dump by opcode: executed warp-instructions, stall samples."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
iS, iSamp, iInst, iThr = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed'), hdr.index('Thread Instructions Executed')
agg = collections.defaultdict(lambda: [0, 0, 0])
tot = [0, 0, 0]
for r in rows[2:]:
    if r[0] == "Address" or r[0] == "Kernel Name": continue
    if len(r) <= iThr: continue
    ops = r[iS].split()
    op = ops[1] if ops and ops[0].startswith('@') else (ops[0] if ops else '?')
    op = '.'.join(op.split('.')[:2])
    v = [int(float(r[iSamp] or 0)), int(float(r[iInst] or 0)), int(float(r[iThr] or 0))]
    for j in range(3):
        agg[op][j] += v[j]; tot[j] += v[j]
print('total samples %d  warp-inst %d  thread-inst %d' % tuple(tot))
for op, v in sorted(agg.items(), key=lambda x: -x[1][0])[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    print('%-22s samples %6.2f%%  warp-inst %6.2f%%  avg-threads %5.1f' % (op, 100 * v[0] / tot[0], 100 * v[1] / tot[1], v[2] / max(v[1], 1)))
