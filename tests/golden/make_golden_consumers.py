#!/usr/bin/env python3
"""
tests/golden/make_golden_consumers.py -- golden vectors for the consumers of the per-window LLRs
(SURVEY 8f rank 4), produced by EXECUTING THE UNMODIFIED REFERENCE: PLOT_PANEL.bin_genomic_windows / binning /
detect_crossovers / detect_crossovers_v2 and BALANCED_ROC_CURVE.binning from /root/reference.

    python tests/golden/make_golden_consumers.py        -> tests/golden/consumers.p.bz2

Inputs: the LLRs of three committed chromosome cases (chrom_*.p.bz2) and two synthetic LLR landscapes with
ragged draw counts, gaps between windows and BPH/SPH regime switches (seeded, stored with the outputs).
Nothing here is imported by the product; the reference is only read, never copied.
"""
import bz2, contextlib, io, os, pickle, random, sys
from operator import attrgetter

REF = '/root/reference'
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REF)
with contextlib.redirect_stdout(io.StringIO()):
    import PLOT_PANEL as PP
    import BALANCED_ROC_CURVE as BRC


def load(name):
    with bz2.open(os.path.join(HERE, name), 'rb') as f:
        return pickle.load(f)


def landscape(seed, chr_id, n_windows, draws, regimes):
    """ windows with irregular sizes and gaps over the whole chromosome; LLRs ~ N(regime mean, 0.4) with a few sentinels """
    rng = random.Random(seed)
    L = PP.chr_length(chr_id)
    step = L // (n_windows + 3)
    out, pos = {}, rng.randrange(1000, step)
    for w in range(n_windows):
        size = rng.choice((100000, 100000, 150000, 220000))
        a, b = pos, pos + size - 1
        pos += max(size, step) + (rng.randrange(0, 3 * step) if rng.random() < 0.04 else 0)
        if b >= L:
            break
        mu = regimes[min(len(regimes) - 1, int(len(regimes) * w / n_windows))]
        k = draws if rng.random() < 0.8 else rng.randrange(max(2, draws // 2), draws + 1)
        vals = [rng.gauss(mu, 0.4) for _ in range(k)]
        if rng.random() < 0.1:
            vals[rng.randrange(k)] = rng.choice((1.23456789, -1.23456789, 0))
        out[(a, b)] = tuple(vals)
    return out


cases = {}
for name in ('homog_m4', 'homog_m3o', 'recent_m4'):
    lik = {w: [PP.likelihoods_tuple(*l) for l in ls] for w, ls in load('chrom_%s.p.bz2' % name)['likelihoods'].items()}
    for pair in (('BPH', 'SPH'), ('BPH', 'disomy')):
        cases['%s/%s-%s' % ((name,) + pair)] = ('chr21', {w: tuple(PP.LLR(attrgetter(pair[0])(l), attrgetter(pair[1])(l)) for l in ls)
                                                          for w, ls in lik.items()})
cases['landscape21'] = ('chr21', landscape(11, 'chr21', 300, 12, (0.25, -0.3, 0.2, -0.25)))
cases['landscape7'] = ('chr7', landscape(12, 'chr7', 420, 9, (-0.2, 0.3, 0.35, -0.3, 0.25)))

golden = {}
for name, (chr_id, llrs) in cases.items():
    info = {'chr_id': chr_id}
    windows = [*llrs]
    rec = {'chr_id': chr_id, 'LLRs': llrs, 'bins': {}, 'binning_plot': {}, 'binning_roc': {}, 'crossovers': {}}
    for nb in (1, 5, 11, PP.chr_length(chr_id) // 4000000, 58, 200, 1000):
        rec['bins'][nb] = PP.bin_genomic_windows(windows, chr_id, nb)
        assert rec['bins'][nb] == BRC.bin_genomic_windows(windows, chr_id, nb)
        rec['binning_plot'][nb] = PP.binning(llrs, info, nb)
        rec['binning_roc'][nb] = BRC.binning(llrs, info, nb)
    mv = [PP.mean_and_var(v) for v in llrs.values()]
    means, variances = [m for m, v in mv], [v for m, v in mv]
    rec['mean_var'] = (means, variances)
    for z, look in ((1.96, 20), (1.0, 5), (2.5, 10), (0.5, 3)):
        with contextlib.redirect_stdout(io.StringIO()):
            v1 = PP.detect_crossovers(windows, means, variances, z_score=z, lookahead=look)
            try:
                v2 = PP.detect_crossovers_v2(windows, means, variances, z_score=z, lookahead=look)
            except Exception as e:          # the reference's recovery step can fail on an empty candidate range
                v2 = type(e).__name__
        rec['crossovers'][(z, look)] = (v1, v2)
    golden[name] = rec
    print('%-24s windows=%4d draws=%5d  crossovers(1.96,20)=%d/%s  (0.5,3)=%d/%s' % (
        name, len(windows), sum(map(len, llrs.values())), len(rec['crossovers'][(1.96, 20)][0]),
        len(rec['crossovers'][(1.96, 20)][1]) if isinstance(rec['crossovers'][(1.96, 20)][1], dict) else rec['crossovers'][(1.96, 20)][1],
        len(rec['crossovers'][(0.5, 3)][0]),
        len(rec['crossovers'][(0.5, 3)][1]) if isinstance(rec['crossovers'][(0.5, 3)][1], dict) else rec['crossovers'][(0.5, 3)][1]))

with bz2.open(os.path.join(HERE, 'consumers.p.bz2'), 'wb') as f:
    pickle.dump(golden, f, protocol=4)
print('consumers.p.bz2 %.1f KB' % (os.path.getsize(os.path.join(HERE, 'consumers.p.bz2')) / 1024))
