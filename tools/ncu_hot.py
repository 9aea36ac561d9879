"""Executed-instruction breakdown of an ncu source-page CSV (ncu -i X.ncu-rep --page source --csv):
opcode histogram weighted by execution count, and contiguous address ranges with their share."""
import csv
import sys
from collections import Counter

rows = list(csv.reader(open(sys.argv[1])))
hdr, data = rows[1], rows[2:]
ia, isrc, ismp = hdr.index('Instructions Executed'), hdr.index('Source'), hdr.index('# Samples')
tot = sum(int(r[ia]) for r in data)
tots = sum(int(r[ismp]) for r in data)
print(len(data), 'SASS instrs; executed', tot, 'samples', tots)
c, cs = Counter(), Counter()
for r in data:
    t = r[isrc].split()
    op = t[1] if t[0].startswith('@') else t[0]
    c[op.rstrip(';')] += int(r[ia])
    cs[op.rstrip(';')] += int(r[ismp])
for op, n in c.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 30):
    print(f'{op:22s} {n/tot*100:6.2f}%  samples {cs[op]/tots*100:6.2f}%')
start, prev = 0, None
for i, r in enumerate(data + [None]):
    n = int(r[ia]) if r else -1
    if prev is None:
        prev = n
    if r is None or abs(n - prev) > max(20000, 0.12 * prev):
        seg = data[start:i]
        s = sum(int(x[ia]) for x in seg)
        sm = sum(int(x[ismp]) for x in seg)
        if s / tot > 0.004:
            print(f'[{start:4d},{i:4d}) n={i-start:4d} exec/instr~{prev:8d} share {s/tot*100:5.1f}% samples {sm/tots*100:5.1f}%  {seg[0][isrc][:50]}')
        start, prev = i, n
