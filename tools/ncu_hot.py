#!/usr/bin/env python3
"""Hot regions of a kernel from an .ncu-rep source page (SASS view): instruction and stall-sample totals per
address block, plus the top instructions.  usage: tools/ncu_hot.py file.ncu-rep [block=64]"""
import csv, subprocess, sys
rep = sys.argv[1]
blk = int(sys.argv[2]) if len(sys.argv) > 2 else 64
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
ia, isrc, isamp, iinst = hdr.index('Address'), hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
ithr = hdr.index('Thread Instructions Executed')
data = [r for r in rows[2:] if len(r) > iinst and r[ia].startswith('0x')]
tot_i = sum(int(r[iinst]) for r in data); tot_s = sum(int(r[isamp]) for r in data)
print('instructions %d  samples %d  sass lines %d' % (tot_i, tot_s, len(data)))
print('--- per %d-instruction block: [first..last line] inst%% samples%% avg-threads  first-op' % blk)
for b in range(0, len(data), blk):
    chunk = data[b:b + blk]
    ci = sum(int(r[iinst]) for r in chunk); cs = sum(int(r[isamp]) for r in chunk); ct = sum(int(r[ithr]) for r in chunk)
    if ci * 200 < tot_i and cs * 200 < tot_s:
        continue
    ops = {}
    for r in chunk:
        op = r[isrc].split()[0] if not r[isrc].strip().startswith('@') else r[isrc].split()[1]
        op = op.split('.')[0]
        ops[op] = ops.get(op, 0) + int(r[iinst])
    top = sorted(ops.items(), key=lambda kv: -kv[1])[:4]
    print('%5d..%5d  inst %5.1f%%  samples %5.1f%%  thr/inst %4.1f  %s' % (b, b + len(chunk) - 1, 100 * ci / tot_i, 100 * cs / max(tot_s, 1),
                                                                   ct / max(ci, 1), ' '.join('%s:%.0f%%' % (k, 100 * v / max(ci, 1)) for k, v in top)))
print('--- top sampled instructions')
for r in sorted(data, key=lambda r: -int(r[isamp]))[:14]:
    print('%6d samples %9d inst  line %5d  %s' % (int(r[isamp]), int(r[iinst]), data.index(r), r[isrc].strip()[:90]))
