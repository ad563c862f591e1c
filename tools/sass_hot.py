#!/usr/bin/env python
"""Aggregates an `ncu --page source --csv` export: executed warp-instructions per opcode and per code region."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
# find header rows (there may be several kernels; take the first block)
hi = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
hdr = rows[hi]
ia, isrc, iex, ith, ismp = hdr.index('Address'), hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('Thread Instructions Executed'), hdr.index('# Samples')
ops = collections.Counter(); tot = 0; smp = collections.Counter()
body = []
for r in rows[hi + 1:]:
    if len(r) <= iex or r[0] == 'Address' or r[0] == 'Kernel Name': break
    try: n = int(r[iex]); t = int(r[ith]); s = int(r[ismp])
    except ValueError: continue
    src = r[isrc].strip()
    toks = src.split()
    op = toks[1] if toks and toks[0].startswith('@') and len(toks) > 1 else (toks[0] if toks else '?')
    op = op.split('.')[0]
    ops[op] += n; smp[op] += s; tot += n
    body.append((n, t, s, src))
print('total warp inst', tot)
for op, n in ops.most_common(28):
    print(f'{op:12s} {n:14d} {100.0*n/tot:6.2f}%   samples {smp[op]}')
if len(sys.argv) > 2:
    vox = float(sys.argv[2])
    print('warp-inst*32/voxel =', tot * 32 / vox)
