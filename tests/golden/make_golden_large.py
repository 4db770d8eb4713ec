#!/usr/bin/env python3
"""
tests/golden/make_golden_large.py -- golden vectors the small generator (make_golden.py) leaves out, again made by
EXECUTING THE UNMODIFIED REFERENCE from /root/reference:

  large_draws.p.bz2   draws of 10..NMAX reads (default 16; 17/18 with an argument, BUILD(18) takes ~40 min and ~25 GB)
                      through the general likelihoods() of all three model classes, with the model tables of
                      MAKE_STATISTICAL_MODEL.BUILD(NMAX) (:129-143).  Inputs are the committed LD-structured
                      chromosome of chrom_inputs.p.bz2 (331 + 200 haplotypes, deep intersections non-empty); recorded:
                      the reads of every draw, the likelihoods, and the joint-frequency table (in full up to 12 reads,
                      as a sha256 of the uint32 array beyond).
  chr21_5008.p.bz2    configs[0] of BASELINE.json: a chr21-sized synthetic disomy sample at 0.05x from
                      MIX_HAPLOIDS.build_obs_tab on a 5008-haplotype iid panel, through the unmodified
                      ANEUPLOIDY_TEST.bootstrap() (:250-287, -w 100000 -s 32 -m 6 -M 4) and statistics() with every
                      random.sample recorded.  The SNP density is reduced (100 000 SNPs instead of 1 M) so that the
                      fixture stays small; the panel is NOT stored -- chr21_panel() below regenerates it from its seed
                      (the tests import this file for that function only; nothing else here runs on import).

    PYTHONHASHSEED=0 python tests/golden/make_golden_large.py [NMAX]

The script appends to an existing large_draws.p.bz2 (sizes already present are kept), so 17 and 18 can be added later.
"""
import bz2
import hashlib
import os
import pickle
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = '/root/reference'
CHR21_LEN = 46709983                     # MIX_HAPLOIDS.chr_length('chr21'), MIX_HAPLOIDS.py:38
CHR21_SNPS, CHR21_HAP, CHR21_SEED = 100000, 5008, 12345
CHR21_HAPLOIDS = (17, 4242)              # panel haplotypes the synthetic embryo is made of


def chr21_panel():
    """ (leg_tab, hap_tab) of the configs[0] panel: CHR21_SNPS positions, one group 'ALL' of 5008 haplotypes, rows
    iid random bits.  Deterministic in CHR21_SEED (CPython's random is stable across versions for these calls). """
    rng = random.Random(CHR21_SEED)
    positions = sorted(rng.sample(range(1, CHR21_LEN + 1), CHR21_SNPS))
    refalt = [tuple(rng.sample('ACGT', 2)) for _ in positions]
    rows = tuple(rng.getrandbits(CHR21_HAP) for _ in positions)
    leg = tuple(('chr21', p, r, a) for p, (r, a) in zip(positions, refalt))
    return leg, rows


def counts_digest(counts, n):
    """ sha256 of the joint-frequency table as little-endian uint32[2^n] (entry 0 = 0) """
    import numpy as np
    arr = np.zeros(1 << n, dtype='<u4')
    for s, c in counts.items():
        arr[s] = c
    return hashlib.sha256(arr.tobytes()).hexdigest()


def plain(x):
    if isinstance(x, tuple) and hasattr(x, '_fields'):
        return tuple(plain(i) for i in x)
    if isinstance(x, dict):
        return {plain(k): plain(v) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return type(x)(plain(i) for i in x)
    return x


def dump(name, obj):
    with bz2.open(os.path.join(HERE, name), 'wb') as f:
        pickle.dump(obj, f, protocol=4)
    print('%-22s %8.1f KB' % (name, os.path.getsize(os.path.join(HERE, name)) / 1024))


def main():
    import contextlib
    import io
    import tempfile
    import time
    if os.environ.get('PYTHONHASHSEED') != '0':
        os.environ['PYTHONHASHSEED'] = '0'
        os.execv(sys.executable, [sys.executable] + sys.argv)
    nmax = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    only = sys.argv[2] if len(sys.argv) > 2 else 'all'
    os.chdir(tempfile.mkdtemp(prefix='ldpgta_golden_large_'))
    sys.path.insert(0, REF)
    with contextlib.redirect_stdout(io.StringIO()):
        import MAKE_STATISTICAL_MODEL as MSM
        import HOMOGENOUES_MODELS as HM
        import RECENT_ADMIXTURE_MODELS as RM
        import DISTANT_ADMIXTURE_MODELS as DM
        import ANEUPLOIDY_TEST as AT
        import MIX_HAPLOIDS as MIX

    if only in ('all', 'draws'):
        # ---- 1. draws of 10..nmax reads, three model classes ---------------------------------------------------
        path = os.path.join(HERE, 'large_draws.p.bz2')
        have = {}
        if os.path.exists(path):
            with bz2.open(path, 'rb') as f:
                have = pickle.load(f)
        with bz2.open(os.path.join(HERE, 'chrom_inputs.p.bz2'), 'rb') as f:
            inputs = pickle.load(f)
        t0 = time.time()
        with contextlib.redirect_stdout(io.StringIO()):
            models = MSM.BUILD(nmax)
        print('BUILD(%d): %.0f s' % (nmax, time.time() - t0))
        obs_tab, leg_tab = inputs['obs_tabs']['disomy'], inputs['leg_tab']
        hap_tab, nhap = inputs['hap_tab'], inputs['number_of_haplotypes']
        makeup = {'G0': 0.6, 'G1': 0.4}
        with contextlib.redirect_stdout(io.StringIO()):
            homog = HM.homogeneous(obs_tab, leg_tab, {'G0': hap_tab['G0']}, {'G0': nhap['G0']}, models)
            recent = RM.recent_admixture(obs_tab, leg_tab, hap_tab, nhap, models)
            distant = DM.distant_admixture(obs_tab, leg_tab, hap_tab, nhap, models, makeup)
        # reads of the sample that matched the panel, in observation order; a draw takes reads that lie close together
        reads = {}
        for pos, rid, base in obs_tab:
            if (pos, base) in homog.hap_dict:
                reads.setdefault(rid, []).append((pos, base))
        ids = list(reads)
        rng = random.Random(4321)
        records = dict(have.get('records', {}))
        for n in range(10, nmax + 1):
            if n in records:
                continue
            start = rng.randrange(len(ids) - 3 * n)
            chosen = rng.sample(ids[start:start + 3 * n], n)
            alleles = tuple(tuple(reads[r]) for r in chosen)
            t0 = time.time()
            counts = homog.joint_frequencies_combo(alleles, normalize=False)
            rec = {'n': n, 'alleles': alleles, 'counts_sha256_G0': counts_digest(counts, n),
                   'counts_G0': counts if n <= 12 else None, 'full_set_count_G0': counts[(1 << n) - 1],
                   'homog_G0': tuple(homog.likelihoods(alleles)), 'recent': tuple(recent.likelihoods(alleles)),
                   'distant': tuple(distant.likelihoods(alleles)), 'makeup': makeup}
            records[n] = rec
            print('n=%2d  %.1f s  homog=%s' % (n, time.time() - t0, rec['homog_G0']))
        dump('large_draws.p.bz2', plain({'records': records, 'groups': ('G0', 'G1'), 'obs': 'disomy',
                                         'note': 'inputs: chrom_inputs.p.bz2; models: MAKE_STATISTICAL_MODEL.BUILD'}))

    if only in ('all', 'chr21'):
        # ---- 2. configs[0]: chr21, 5008 haplotypes, through the unmodified bootstrap() ----------------------------
        leg, rows = chr21_panel()
        obs_dicts = [{l[1]: (l[3] if row >> h & 1 else l[2]) for l, row in zip(leg, rows)} for h in CHR21_HAPLOIDS]
        random.seed(777)
        with contextlib.redirect_stdout(io.StringIO()):
            obs_tab = MIX.build_obs_tab(obs_dicts, 'chr21', 36, 0.05, 'disomy', None)
            models = MSM.BUILD(4)
        leg_tab = tuple(HM.leg_tuple(*l) for l in leg)
        cfg = dict(window_size=100000, subsamples=32, offset=0, min_reads=6, max_reads=4, min_score=2, min_HF=0.05)
        draw_log = []
        _sample = random.sample

        def recording_sample(population, k, **kw):
            out = _sample(population, k, **kw)
            draw_log.append(tuple(out))
            return out
        random.seed(2021)
        random.sample = recording_sample
        try:
            with contextlib.redirect_stdout(io.StringIO()):
                lik, windows, matches = AT.bootstrap(obs_tab, leg_tab, {'ALL': rows}, {'ALL': CHR21_HAP}, {'ALL'}, models,
                                                     cfg['window_size'], cfg['subsamples'], cfg['offset'], cfg['min_reads'],
                                                     cfg['max_reads'], cfg['min_score'], cfg['min_HF'])
        finally:
            random.sample = _sample
        draws, it = {}, iter(draw_log)
        for w, rid in windows.items():
            effN = AT.effective_number_of_subsamples(len(rid), cfg['min_reads'], cfg['max_reads'], cfg['subsamples'])
            if effN:
                draws[w] = [next(it) for _ in range(effN)]
        assert next(it, None) is None and set(draws) == set(lik)
        stats = AT.statistics(lik, windows)
        print('chr21: obs rows=%d windows=%d evaluated=%d draws=%d' % (len(obs_tab), len(windows), len(lik), sum(map(len, lik.values()))))
        dump('chr21_5008.p.bz2', plain({'cfg': cfg, 'obs_tab': obs_tab, 'windows': windows, 'draws': draws, 'likelihoods': lik,
                                        'fraction_of_matches': matches, 'statistics': stats,
                                        'panel': {'snps': CHR21_SNPS, 'haplotypes': CHR21_HAP, 'seed': CHR21_SEED, 'haploids': CHR21_HAPLOIDS},
                                        'panel_sha256': hashlib.sha256(b''.join(r.to_bytes(626, 'little') for r in rows)).hexdigest()}))


if __name__ == '__main__':
    main()
