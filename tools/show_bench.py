#!/usr/bin/env python3
"""One-line digest of a bench.py JSON line: tools/show_bench.py file.json"""
import json, sys
d = json.load(open(sys.argv[1]))
print('n_gpus', d['n_gpus'], 'value %.4g' % d['value'], 'ms/step %.4f' % d['ms_per_step'], 'e2e %.4g' % d['e2e']['value'],
      'e2e ms %.4f' % d['e2e'].get('ms_per_step', float('nan')), 'launches', d.get('gpu_launches'),
      'roofline %.3f' % d['roofline']['frac'] if d.get('roofline') else '', 'clocks', d.get('clocks', {}).get('reasons'))
