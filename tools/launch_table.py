#!/usr/bin/env python
""" Median duration per kernel (name, grid, block) of an ncu launch list (--metrics gpu__time_duration.sum --csv). """
import collections
import csv
import io
import sys


def table(path):
    txt = open(path).read()
    txt = txt[txt.index('"ID"'):]
    agg = collections.OrderedDict()
    for row in csv.DictReader(io.StringIO(txt)):
        v = float(row['Metric Value'].replace(',', ''))
        v *= {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 's': 1e6}.get(row['Metric Unit'], 1.0)
        agg.setdefault((row['Kernel Name'].split('(')[0][:60], row['Grid Size'], row['Block Size']), []).append(v)
    return agg


if __name__ == '__main__':
    for path in sys.argv[1:]:
        print(path)
        for k, v in table(path).items():
            print('  %-60s %-16s %-14s n=%-3d median %8.1f us' % (k + (len(v), sorted(v)[len(v) // 2])))
