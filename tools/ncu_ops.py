"""Opcode histogram (executed warp instructions, stall samples, shared wavefronts) from `ncu --page source --csv`."""
import csv
import sys
from collections import Counter

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 20]
hdr = rows[0]
ia, isamp, iex, iw = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed'), hdr.index('L1 Wavefronts Shared')
body = [r for r in rows[1:] if r[isamp].isdigit()]
tot_s = sum(int(r[isamp]) for r in body)
tot_e = sum(int(r[iex]) for r in body)
print('total samples', tot_s, 'instr', tot_e, 'rows', len(body))
ex, sm, wf = Counter(), Counter(), Counter()
for r in body:
    toks = r[ia].split()
    op = toks[1] if toks[0].startswith('@') else toks[0]
    op = '.'.join(op.split('.')[:2]) if op.startswith(('IMAD', 'LDS', 'STS', 'SYNCS', 'LD', 'ST')) else op.split('.')[0]
    ex[op] += int(r[iex]); sm[op] += int(r[isamp]); wf[op] += int(r[iw] or 0)
print('%-16s %10s %6s %8s %6s %10s' % ('op', 'executed', '%', 'samples', '%', 'smem_wf'))
for op, c in ex.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 30):
    print('%-16s %10d %6.1f %8d %6.1f %10d' % (op, c, 100 * c / tot_e, sm[op], 100 * sm[op] / tot_s, wf[op]))
