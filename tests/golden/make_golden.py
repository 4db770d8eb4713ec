#!/usr/bin/env python3
"""
tests/golden/make_golden.py -- regenerates the golden vectors by EXECUTING THE
UNMODIFIED REFERENCE from /root/reference (it is pure stdlib Python and imports
here as is).  The outputs are committed next to this script so that the tests
can run on the GPU box, where /root/reference does not exist.

    PYTHONHASHSEED=0 python tests/golden/make_golden.py

What is recorded
  models_ref.p.bz2      MAKE_STATISTICAL_MODEL.BUILD(7) verbatim + per-n digests up to 9
  testmodule.p.bz2      TEST_MODULE fixtures (seed 2021) and, for the three model
                        classes, joint frequencies and likelihoods (general and
                        closed form) on many allele/haplotype tuples, n = 2..9
  chrom_inputs.p.bz2    small LD-structured synthetic chromosome (two obs tables, two panel groups)
  chrom_<cfg>.p.bz2     per configuration: the reference's read scores, windows, every bootstrap draw
                        (random.sample is wrapped, pick_reads itself is untouched),
                        likelihoods, fraction_of_matches and statistics()

Nothing here is imported by the product; the reference is only read, never copied.
"""
import bz2, collections, hashlib, io, os, pickle, random, sys, tempfile, contextlib

REF = '/root/reference'
HERE = os.path.dirname(os.path.abspath(__file__))

if os.environ.get('PYTHONHASHSEED') != '0':
    os.environ['PYTHONHASHSEED'] = '0'
    os.execv(sys.executable, [sys.executable] + sys.argv)

work = tempfile.mkdtemp(prefix='ldpgta_golden_')
os.chdir(work)                       # TEST_MODULE and BUILD write into the CWD
sys.path.insert(0, REF)

with contextlib.redirect_stdout(io.StringIO()):
    import TEST_MODULE
    import MAKE_STATISTICAL_MODEL as MSM
    import HOMOGENOUES_MODELS as HM
    import RECENT_ADMIXTURE_MODELS as RM
    import DISTANT_ADMIXTURE_MODELS as DM
    import ANEUPLOIDY_TEST as AT
    import MIX_HAPLOIDS as MIX
    MODELS = MSM.BUILD(9)


def dump(name, obj):
    with bz2.open(os.path.join(HERE, name), 'wb') as f:
        pickle.dump(obj, f, protocol=4)
    print('%-22s %8.1f KB' % (name, os.path.getsize(os.path.join(HERE, name)) / 1024))


def plain(x):
    """ namedtuples/defaultdicts -> plain containers so the fixtures unpickle without the reference. """
    if isinstance(x, tuple) and hasattr(x, '_fields'):
        return tuple(plain(i) for i in x)
    if isinstance(x, dict):
        return {plain(k): plain(v) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return type(x)(plain(i) for i in x)
    return x


def canon(model):
    """ Order-insensitive canonical form of one scenario's model table. """
    out = {}
    for nblocks, classes in model.items():
        c = {}
        for w, members in classes.items():
            if isinstance(members, dict):
                c[w] = sorted((b0, tuple(sorted(v))) for b0, v in members.items())
            else:
                c[w] = sorted(members)
        out[nblocks] = sorted(c.items())
    return sorted(out.items())


def digest(models_n):
    return hashlib.sha256(repr({k: canon(v) for k, v in sorted(models_n.items())}).encode()).hexdigest()


# ---- 1. model tables ------------------------------------------------------
dump('models_ref.p.bz2', {'models': {n: MODELS[n] for n in range(2, 8)},
                          'digests': {n: digest(MODELS[n]) for n in range(2, 10)}})

# ---- 2. TEST_MODULE fixtures, the three model classes ----------------------
leg_tab, hap_tab, nhap, obs_tab = TEST_MODULE.leg_tab, TEST_MODULE.hap_tab, TEST_MODULE.number_of_haplotypes, TEST_MODULE.obs_tab
with contextlib.redirect_stdout(io.StringIO()):
    homog = {g: HM.homogeneous(obs_tab, leg_tab, {g: hap_tab[g]}, {g: nhap[g]}, MODELS) for g in hap_tab}
    recent = RM.recent_admixture(obs_tab, leg_tab, hap_tab, nhap, MODELS)
    makeups = [{'group0': 0.5, 'group1': 0.5}, {'group1': 0.3, 'group0': 0.7}]
    distant = [DM.distant_admixture(obs_tab, leg_tab, hap_tab, nhap, MODELS, mk) for mk in makeups]

keys = tuple(homog['group0'].hap_dict.keys())
rng = random.Random(2022)
cases = []
x = 544                               # the offset SURVEY section 4 quotes
for n in range(2, 10):
    cases.append(tuple(keys[x:x + n]))                       # bare alleles (the reference's self-check shape)
for n in range(2, 10):
    for rep in range(3):
        start = rng.randrange(len(keys) - 40)
        pool = keys[start:start + 40]
        reads = []
        for i in range(n):
            k = rng.choice((1, 1, 2, 3))
            reads.append(tuple(rng.sample(pool, k)))         # tuples of alleles = haplotypes/reads
        cases.append(tuple(reads))
cases.append((keys[10], (keys[11], keys[12]), keys[13], (keys[14],)))   # mixed bare / tuple forms

records = []
for alleles in cases:
    n = len(alleles)
    rec = {'alleles': alleles, 'n': n}
    for g, model in homog.items():
        rec['homog_counts_' + g] = model.joint_frequencies_combo(alleles, normalize=False)
        if all(type(a[0]) == int or len(a) > 1 for a in alleles):   # the redundant variant breaks on 1-allele tuples
            assert rec['homog_counts_' + g] == model.joint_frequencies_redundant(alleles, normalize=False)
        rec['homog_general_' + g] = tuple(model.likelihoods(alleles))
        rec['homog_get_' + g] = tuple(model.get_likelihoods(alleles))
    rec['recent_general'] = tuple(recent.likelihoods(alleles))
    rec['recent_get'] = tuple(recent.get_likelihoods(alleles))
    for i, d in enumerate(distant):
        rec['distant%d_freq' % i] = d.joint_frequencies_combo(alleles)
        rec['distant%d_general' % i] = tuple(d.likelihoods(alleles))
        if n != 5:
            rec['distant%d_get' % i] = tuple(d.get_likelihoods(alleles))
        else:
            try:
                d.get_likelihoods(alleles)
                rec['distant%d_get_error' % i] = None
            except AttributeError as e:
                rec['distant%d_get_error' % i] = 'AttributeError'
    records.append(rec)

dump('testmodule.p.bz2', plain({'leg_tab': leg_tab, 'hap_tab': hap_tab, 'number_of_haplotypes': nhap,
                                'obs_tab': obs_tab, 'makeups': makeups, 'records': records,
                                'fraction_of_matches': {g: m.fraction_of_matches for g, m in homog.items()}}))

# ---- 3. synthetic chromosome through the reference's bootstrap()/statistics() ----
def ld_panel(rng, positions, n_hap, founders=12, switch=0.02):
    """ Haplotypes as mosaics of a few founders with a skewed allele-frequency
    spectrum, so that deep intersections stay non-zero. Bit i of a
    row = haplotype i. """
    S = len(positions)
    found = []
    freqs = [min(0.95, max(0.02, rng.betavariate(0.5, 1.2))) for _ in range(S)]
    for _ in range(founders):
        found.append([1 if rng.random() < p else 0 for p in freqs])
    src = [rng.randrange(founders) for _ in range(n_hap)]
    rows = []
    for s in range(S):
        row = 0
        for h in range(n_hap):
            if rng.random() < switch:
                src[h] = rng.randrange(founders)
            bit = found[src[h]][s]
            if rng.random() < 0.01:
                bit ^= 1
            row |= bit << h
        rows.append(row)
    return rows


def haploid(rows, positions, refalt, h):
    return {pos: (ra[1] if row >> h & 1 else ra[0]) for pos, row, ra in zip(positions, rows, refalt)}


rng = random.Random(12345)
SPAN = 5_000_000
positions = sorted(rng.sample(range(1, SPAN), 6000))
refalt = [tuple(rng.sample('ACGT', 2)) for _ in positions]
sizes = {'G0': 331, 'G1': 200}
panel = {g: ld_panel(rng, positions, n) for g, n in sizes.items()}

random.seed(777)
obs_dicts = [haploid(panel['G0'], positions, refalt, 17), haploid(panel['G1'], positions, refalt, 42),
             haploid(panel['G0'], positions, refalt, 5)]
with contextlib.redirect_stdout(io.StringIO()):
    obs_full = {sc: MIX.build_obs_tab(obs_dicts, 'chr21', 36, depth, sc, None)
                for sc, depth in (('disomy', 0.45), ('BPH', 0.45))}

# a few observations that miss the panel (third base) / unknown position, to exercise mismatches
def with_noise(obs, rng):
    obs = list(obs)
    for _ in range(25):
        i = rng.randrange(len(obs))
        pos, rid, base = obs[i]
        other = [b for b in 'ACGT' if b not in refalt[positions.index(pos)]]
        obs[i] = MIX.obs_tuple(pos, rid, other[0])
    obs.append(MIX.obs_tuple(SPAN + 10, 'stray.read', 'A'))
    obs.sort(key=lambda r: r[0])
    return obs

observed = set()
obs_noisy = {}
for sc, obs in obs_full.items():
    obs_noisy[sc] = with_noise(obs, rng)
    observed.update(p for p, _, _ in obs)
keep = [i for i, p in enumerate(positions) if p in observed or rng.random() < 0.05]
leg = tuple(HM.leg_tuple('chr21', positions[i], *refalt[i]) for i in keep)
hap = {g: tuple(panel[g][i] for i in keep) for g in panel}

draw_log = []
_sample = random.sample
def recording_sample(population, k, **kw):
    out = _sample(population, k, **kw)
    draw_log.append(tuple(out))
    return out

CONFIGS = {
    'homog_m4':   dict(scenario='disomy', groups=('G0',), makeup={'G0'}, window_size=100000, subsamples=8, offset=0, min_reads=6, max_reads=4, min_score=2, min_HF=0.05),
    'homog_m8a':  dict(scenario='BPH', groups=('G0',), makeup={'G0'}, window_size=0, subsamples=6, offset=0, min_reads=12, max_reads=8, min_score=2, min_HF=0.05),
    'homog_m3o':  dict(scenario='BPH', groups=('G0',), makeup={'G0'}, window_size=60000, subsamples=5, offset=20000, min_reads=3, max_reads=3, min_score=1, min_HF=0.05),
    'recent_m4':  dict(scenario='disomy', groups=('G0', 'G1'), makeup={'G0', 'G1'}, window_size=100000, subsamples=8, offset=0, min_reads=6, max_reads=4, min_score=2, min_HF=0.05),
    'recent_m6':  dict(scenario='disomy', groups=('G0', 'G1'), makeup={'G0', 'G1'}, window_size=150000, subsamples=4, offset=0, min_reads=8, max_reads=6, min_score=2, min_HF=0.05),
    'distant_m4': dict(scenario='disomy', groups=('G0', 'G1'), makeup={'G0': 0.6, 'G1': 0.4}, window_size=100000, subsamples=8, offset=0, min_reads=6, max_reads=4, min_score=2, min_HF=0.05),
    'distant_m6': dict(scenario='BPH', groups=('G1', 'G0'), makeup={'G1': 0.25, 'G0': 0.75}, window_size=150000, subsamples=4, offset=0, min_reads=8, max_reads=6, min_score=2, min_HF=0.05),
}

for name, cfg in CONFIGS.items():
    obs_tab_c = obs_noisy[cfg['scenario']]
    hap_c = {g: hap[g] for g in cfg['groups']}
    nhap_c = {g: sizes[g] for g in cfg['groups']}
    args = (obs_tab_c, leg, hap_c, nhap_c, cfg['makeup'], MODELS, cfg['window_size'], cfg['subsamples'],
            cfg['offset'], cfg['min_reads'], cfg['max_reads'], cfg['min_score'], cfg['min_HF'])
    random.seed(2021)
    del draw_log[:]
    random.sample = recording_sample
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            likelihoods, windows, matches = AT.bootstrap(*args)
            obs_support = AT.supporting_dictionaries(obs_tab_c, leg, hap_c, nhap_c, cfg['makeup'], cfg['min_HF'])
    finally:
        random.sample = _sample
    # attribute the flat draw log to windows in evaluation order (ANEUPLOIDY_TEST.py:277-286)
    draws, it = {}, iter(draw_log)
    for w, ids in windows.items():
        effN = AT.effective_number_of_subsamples(len(ids), cfg['min_reads'], cfg['max_reads'], cfg['subsamples'])
        if effN:
            draws[w] = [next(it) for _ in range(effN)]
    assert next(it, None) is None and set(draws) == set(likelihoods)
    stats = AT.statistics(likelihoods, windows)
    n_eval = sum(len(v) for v in likelihoods.values())
    print('%-11s windows=%d evaluated=%d draws=%d' % (name, len(windows), len(likelihoods), n_eval))
    dump('chrom_%s.p.bz2' % name, plain({
        'cfg': cfg, 'score': dict(obs_support.score), 'windows': windows, 'draws': draws,
        'likelihoods': likelihoods, 'fraction_of_matches': matches, 'statistics': stats}))

# inputs shared by all chromosome configs (cfg['scenario'] picks the obs table, cfg['groups'] the panel groups)
dump('chrom_inputs.p.bz2', plain({'obs_tabs': obs_noisy, 'leg_tab': leg, 'hap_tab': hap, 'number_of_haplotypes': sizes}))
