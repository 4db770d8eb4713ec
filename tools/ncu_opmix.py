#!/usr/bin/env python3
"""Dynamic opcode mix of a kernel from an .ncu-rep (SASS source page). usage: tools/ncu_opmix.py file.ncu-rep [units]
units = number of work items (e.g. draws) to normalise by."""
import csv, subprocess, sys
rep = sys.argv[1]
units = float(sys.argv[2]) if len(sys.argv) > 2 else None
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
isrc, iinst = hdr.index('Source'), hdr.index('Instructions Executed')
ops = {}
tot = 0
for r in rows[2:]:
    if len(r) <= iinst or not r[0].startswith('0x'):
        continue
    toks = r[isrc].split()
    op = toks[1] if toks[0].startswith('@') else toks[0]
    op = op.rstrip(';')
    base = op.split('.')[0]
    if base == 'IMAD':
        base = 'IMAD.' + (op.split('.')[1] if '.' in op else '')
    n = int(r[iinst])
    ops[base] = ops.get(base, 0) + n
    tot += n
print('total warp instructions: %d%s' % (tot, '  (%.1f per unit)' % (tot / units) if units else ''))
for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:28]:
    print('  %-14s %12d  %5.1f%%%s' % (k, v, 100 * v / tot, '  %8.2f/unit' % (v / units) if units else ''))
