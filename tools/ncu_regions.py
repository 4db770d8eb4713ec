#!/usr/bin/env python3
"""
tools/ncu_regions.py <source.csv[.gz]> [instructions per region] -- where a kernel's time goes, from the SASS view of an ncu
report (ncu -i x.ncu-rep --page source --csv): warp-stall samples and executed instructions summed over consecutive runs of
SASS instructions, with the dominant opcodes and stall reasons of each run.  Device functions linked with -rdc (the
polynomials of poly_general.cu) appear as separate sections.
"""
import collections
import csv
import gzip
import sys


def main():
    path = sys.argv[1]
    seg = int(sys.argv[2]) if len(sys.argv) > 2 else 150
    rows = list(csv.reader(gzip.open(path, 'rt') if path.endswith('.gz') else open(path)))
    hdr = None
    sections, cur = [], None
    for r in rows:
        if r and r[0] == 'Kernel Name':
            cur = {'name': r[1], 'data': []}
            sections.append(cur)
        elif r and r[0] == 'Address':
            hdr = r
        elif cur is not None and hdr is not None and len(r) == len(hdr):
            cur['data'].append(r)
    ci = {h: i for i, h in enumerate(hdr)}
    S, I, SRC = ci['# Samples'], ci['Instructions Executed'], ci['Source']
    tot = sum(int(r[S]) for s in sections for r in s['data'])
    toti = sum(int(r[I]) for s in sections for r in s['data'])
    print('total samples %d, warp instructions %d' % (tot, toti))
    stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
    for s in sections:
        d = s['data']
        ss = sum(int(r[S]) for r in d)
        print('== %s: %.1f %% of samples, %.1f %% of instructions' % (s['name'][:110], 100.0 * ss / tot, 100.0 * sum(int(r[I]) for r in d) / toti))
        for k in range(0, len(d), seg):
            ch = d[k:k + seg]
            sm = sum(int(r[S]) for r in ch)
            if sm < tot * 0.01:
                continue
            ops = collections.Counter()
            for r in ch:
                t = r[SRC].split()
                ops[(t[1] if t[0].startswith('@') else t[0]).split('.')[0]] += int(r[I])
            st = collections.Counter({h: sum(int(r[ci[h]]) for r in ch) for h in stalls})
            print('  sass %5d..%5d  samples %5.1f %%  instructions %5.1f %%  %s | %s' % (
                k, k + len(ch), 100.0 * sm / tot, 100.0 * sum(int(r[I]) for r in ch) / toti,
                ' '.join('%s:%.0f%%' % (a, 100.0 * b / max(1, sum(ops.values()))) for a, b in ops.most_common(5)),
                ' '.join('%s:%.0f%%' % (a[6:], 100.0 * b / max(1, sm)) for a, b in st.most_common(4))))


if __name__ == '__main__':
    main()
