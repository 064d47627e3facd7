"""Histogram of executed warp instructions by opcode from `ncu --page source --csv --print-source sass`."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
hdr = rows[hi]
ci = hdr.index('Instructions Executed'); si = hdr.index('Source'); ss = hdr.index('# Samples')
ops = collections.Counter(); samp = collections.Counter(); total = 0; n = 0
for r in rows[hi + 1:]:
    if len(r) <= ci: continue
    try: c = int(r[ci])
    except ValueError: continue
    src = r[si].strip()
    toks = src.split()
    op = toks[1] if toks and toks[0].startswith('@') else (toks[0] if toks else '?')
    op = op.split('.')[0] + ('.' + op.split('.')[1] if op.startswith(('LDS', 'STS', 'LDG', 'STG', 'MUFU')) and '.' in op else '')
    ops[op] += c; total += c; n += 1
    try: samp[op] += int(r[ss])
    except ValueError: pass
print('static SASS instructions', n, 'executed warp instructions', total)
ts = sum(samp.values()) or 1
for op, c in ops.most_common(28):
    print(f'{op:14s} {c:12d} {100*c/total:5.1f}%   samples {100*samp[op]/ts:5.1f}%')
