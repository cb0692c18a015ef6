"""Summarises `ncu --page source --print-source sass --csv` output: stall samples per opcode and top lines."""
import collections
import csv
import re
import sys

r = list(csv.reader(open(sys.argv[1])))
hi = [i for i, row in enumerate(r) if row and row[0] == 'Address']
inst = int(sys.argv[2]) if len(sys.argv) > 2 else 0
start = hi[inst]
end = hi[inst + 1] if len(hi) > inst + 1 else len(r)
hdr = r[start]
si = hdr.index('Warp Stall Sampling (All Samples)')
src = hdr.index('Source')
ie = hdr.index('Instructions Executed')
rows = [row for row in r[start + 1:end] if len(row) > max(si, ie, src)]
tot = sum(int(x[si] or 0) for x in rows) or 1
print('total samples', tot, 'n instr', len(rows))
ops = collections.Counter()
execs = collections.Counter()
for x in rows:
    m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)', x[src])
    op = m.group(2).split('.')[0] if m else '?'
    ops[op] += int(x[si] or 0)
    execs[op] += int(x[ie] or 0)
for op, c in ops.most_common(18):
    print(f'{c / tot * 100:6.2f}% stall  execs={execs[op]:>10}  {op}')
print('top instr lines:')
for x in sorted(rows, key=lambda x: -int(x[si] or 0))[:24]:
    print(f'{int(x[si]) / tot * 100:6.2f}%  exec={x[ie]:>9}  {x[src][:110]}')
