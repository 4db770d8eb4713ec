"""
oracle/ldpgta_oracle.py  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A plain CPython restatement of LD-PGTA's windowed haplotype-likelihood path
(the reference is pure Python, so the oracle is too).  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this file; nothing under ldpgta_b200/ does.

Parity status: PINNED.  tests/golden/make_golden.py executes the unmodified
reference from /root/reference and commits its inputs and outputs under
tests/golden/; tests/test_oracle_golden.py checks every function below against
those vectors (popcounts bit-exact, likelihoods <= 4 ulp, statistics <= 1e-13); the
LLR consumers (binning, crossover detection) are pinned by
tests/golden/make_golden_consumers.py + tests/test_consumers.py.

Every function cites the reference file:line (paths into /root/reference) whose
behaviour it restates.  Panel rows are arbitrary-precision ints exactly as in
the reference (bit i = haplotype i), `int.bit_count` plays the role of
`gmpy2.popcount` / `popcount_lk` (HOMOGENOUES_MODELS.py:30-44,323-330), which
are exact integer functions, so no third-party arithmetic is involved.
"""

import collections
import math
import random
from fractions import Fraction
from math import comb, gcd, log
from statistics import mean, pstdev, variance

# Result record. The reference defines one copy per module
# (HOMOGENOUES_MODELS.py:27 and twins); field order is part of the contract.
likelihoods_tuple = collections.namedtuple('likelihoods_tuple', ('monosomy', 'disomy', 'SPH', 'BPH'))
leg_tuple = collections.namedtuple('leg_tuple', ('chr_id', 'pos', 'ref', 'alt'))
obs_tuple = collections.namedtuple('obs_tuple', ('pos', 'read_id', 'base'))

LLR_PAIRS = (('BPH', 'SPH'), ('disomy', 'monosomy'), ('BPH', 'disomy'), ('disomy', 'SPH'), ('SPH', 'monosomy'))


def popcount(x):
    """ Hamming weight of a non-negative int (HOMOGENOUES_MODELS.py:30-44, 323-330). """
    return x.bit_count()


# --------------------------------------------------------------------------
# a2 / a3 : panel rows per observed allele, one mask per read
# --------------------------------------------------------------------------

def build_hap_dict(obs_tab, leg_tab, hap_tab, number_of_haplotypes):
    """ (pos, base) -> N-bit row; alt allele keeps the panel row, ref allele
    gets its complement inside N bits; anything else is a mismatch.
    Restates HOMOGENOUES_MODELS.py:76-100 (twins RECENT...:83-107,
    DISTANT...:81-105). Duplicate legend positions: the last row wins (:84). """
    by_pos = {}
    for leg, hap in zip(leg_tab, hap_tab):
        by_pos[leg[1]] = (leg[2], leg[3], hap)
    ones = (1 << number_of_haplotypes) - 1
    hap_dict, mismatches = {}, 0
    for pos, _read_id, base in obs_tab:
        entry = by_pos.get(pos)
        if entry is not None and base == entry[1]:
            hap_dict[(pos, base)] = entry[2]
        elif entry is not None and base == entry[0]:
            hap_dict[(pos, base)] = entry[2] ^ ones
        else:
            mismatches += 1
    return hap_dict, 1 - mismatches / len(obs_tab)


def read_masks(hap_dict, alleles):
    """ One mask per element of `alleles`: a bare (pos, base) allele maps to its
    row, a tuple/list of alleles to the AND of their rows.
    Restates build_intrenal_hap_dict, HOMOGENOUES_MODELS.py:102-127. Unknown
    alleles raise KeyError like the reference. """
    masks = []
    for item in alleles:
        if type(item[0]) == tuple:
            m = hap_dict[item[0]]
            for allele in item[1:]:
                m &= hap_dict[allele]
        elif type(item[0]) == int:
            m = hap_dict[item]
        else:
            raise Exception('error: joint_frequencies only accepts alleles and tuple/list of alleles.')
        masks.append(m)
    return masks


# --------------------------------------------------------------------------
# a4 : joint frequencies over every non-empty subset
# --------------------------------------------------------------------------

def joint_counts(masks):
    """ counts[S] = popcount(AND of masks[i] for bit i in S) for S = 1..2^n-1,
    counts[0] is None.  Same values as joint_frequencies_combo(normalize=False)
    (HOMOGENOUES_MODELS.py:129-163) whose dict keys are these subset bitmasks;
    computed here parent-first instead of from scratch. """
    n = len(masks)
    inter = [None] * (1 << n)
    counts = [None] * (1 << n)
    for s in range(1, 1 << n):
        low = s & -s
        rest = s ^ low
        m = masks[low.bit_length() - 1]
        inter[s] = m if rest == 0 else inter[rest] & m
        counts[s] = inter[s].bit_count()
    return counts


def joint_counts_redundant(masks):
    """ The slow self-check of the reference: every subset ANDed from scratch
    (joint_frequencies_redundant, HOMOGENOUES_MODELS.py:165-183). """
    n = len(masks)
    counts = [None] * (1 << n)
    for s in range(1, 1 << n):
        m = -1
        for i in range(n):
            if s >> i & 1:
                m &= masks[i]
        counts[s] = m.bit_count()
    return counts


# --------------------------------------------------------------------------
# MAKE_STATISTICAL_MODEL restated combinatorially
# --------------------------------------------------------------------------

def _set_partitions(n, max_blocks):
    """ All partitions of reads 0..n-1 into 1..max_blocks blocks; blocks are
    listed by their smallest read, which is the order ENGINE's
    haplotypes.values() yields (MAKE_STATISTICAL_MODEL.py:31-36). """
    def grow(i, blocks):
        if i == n:
            yield tuple(tuple(b) for b in blocks)
            return
        for b in blocks:
            b.append(i)
            yield from grow(i + 1, blocks)
            b.pop()
        if len(blocks) < max_blocks:
            blocks.append([i])
            yield from grow(i + 1, blocks)
            blocks.pop()
    yield from grow(1, [[0]]) if n else iter(())


def _injective_weight(sizes, degeneracies):
    """ Sum over injective maps block->homolog of prod(deg[h]**|block|): the
    multiplicity ENGINE accumulates per partition (MAKE_STATISTICAL_MODEL.py:24-38). """
    from itertools import permutations
    total = 0
    for homologs in permutations(range(len(degeneracies)), len(sizes)):
        w = 1
        for size, h in zip(sizes, homologs):
            w *= degeneracies[h] ** size
        total += w
    return total


def _model(n, degeneracies, group_by_first):
    """ ENGINE + COMPACT (+ COMMON) for one scenario, MAKE_STATISTICAL_MODEL.py:24-71.
    Structure: {blocks: {(w, T) reduced: tuple of block-bitmask tuples}}, blocks
    sorted by size (stable); BPH's 3-block entry is regrouped by first block. """
    k = len(degeneracies)
    T = sum(degeneracies) ** n
    compact = {j + 1: collections.defaultdict(list) for j in range(k)}
    for part in _set_partitions(n, k):
        weight = _injective_weight([len(b) for b in part], degeneracies)
        blocks = tuple(sum(1 << r for r in b) for b in sorted(part, key=len))
        g = gcd(weight, T)
        compact[len(blocks)][weight // g, T // g].append(blocks)
    out = {j: {w: tuple(v) for w, v in d.items()} for j, d in compact.items()}
    if group_by_first and 3 in out:
        regrouped = {}
        for w, triplets in out[3].items():
            nested = collections.defaultdict(list)
            for b0, b1, b2 in triplets:
                nested[b0].append((b1, b2))
            regrouped[w] = {b0: tuple(v) for b0, v in nested.items()}
        out[3] = regrouped
    return out


def make_models(n_max, n_min=2):
    """ Same content as MAKE_STATISTICAL_MODEL.BUILD(n_max) (:129-143): degeneracies
    (1,1,1) BPH, (2,1) SPH, (1,1) DISOMY, (1,) MONOSOMY (:73-102).  List order
    inside a weight class is not the reference's popitem order; sums over a
    class are exact integers in the homogeneous model so order is immaterial
    there, and within float tolerance elsewhere. """
    return {n: {'MONOSOMY': _model(n, (1,), False),
                'DISOMY': _model(n, (1, 1), False),
                'SPH': _model(n, (2, 1), False),
                'BPH': _model(n, (1, 1, 1), True)} for n in range(n_min, n_max + 1)}


# --------------------------------------------------------------------------
# a5 : general-n polynomials, table driven like the reference
# --------------------------------------------------------------------------

def _single(model):
    (w, ((b0,),)), = model[1].items()
    return w, b0


def likelihoods_homogeneous(F, N, model):
    """ HOMOGENOUES_MODELS.py:185-225 with F = integer counts: every class sum is
    an exact int, `int * A0 / A1` is one correctly-rounded true division, class
    results are added as floats, then divided by N**k. """
    (a0, a1), full = _single(model['BPH'])
    bph = F[full] * a0 / (a1 * N)
    bph += sum(sum(F[b0] * F[b1] for b0, b1 in c) * w0 / w1 for (w0, w1), c in model['BPH'][2].items()) / N ** 2
    if model['BPH'].get(3):
        bph += sum(sum(F[b0] * sum(F[b1] * F[b2] for b1, b2 in c[b0]) for b0 in c) * w0 / w1
                   for (w0, w1), c in model['BPH'][3].items()) / N ** 3
    (a0, a1), full = _single(model['SPH'])
    sph = F[full] * a0 / (a1 * N)
    sph += sum(sum(F[b0] * F[b1] for b0, b1 in c) * w0 / w1 for (w0, w1), c in model['SPH'][2].items()) / N ** 2
    (a0, a1), full = _single(model['DISOMY'])
    dis = F[full] * a0 / (a1 * N)
    dis += sum(sum(F[b0] * F[b1] for b0, b1 in c) * w0 / w1 for (w0, w1), c in model['DISOMY'][2].items()) / N ** 2
    _, full = _single(model['MONOSOMY'])
    mono = F[full] / N
    return likelihoods_tuple(mono, dis, sph, bph)


def likelihoods_recent(F, M, G, N, model):
    """ RECENT_ADMIXTURE_MODELS.py:198-247: F counts in group 0 (M haplotypes),
    G counts in group 1 (N haplotypes); each parent comes from one group. """
    (a0, a1), full = _single(model['BPH'])
    bph = (F[full] / M + G[full] / N) * a0 / (2 * a1)
    bph += sum(sum(F[b0] * F[b1] / M ** 2 + G[b0] * G[b1] / N ** 2 + 2 * (F[b0] * G[b1] + G[b0] * F[b1]) / (M * N)
                   for b0, b1 in c) * w0 / w1 for (w0, w1), c in model['BPH'][2].items()) / 6
    if model['BPH'].get(3):
        bph += sum(sum(F[b0] * sum((F[b1] * G[b2] + G[b1] * F[b2]) / M + G[b1] * G[b2] / N for b1, b2 in c[b0])
                       for b0 in c) * w0 / w1 for (w0, w1), c in model['BPH'][3].items()) / (6 * M * N)
        bph += sum(sum(G[b0] * sum((F[b1] * G[b2] + G[b1] * F[b2]) / N + F[b1] * F[b2] / M for b1, b2 in c[b0])
                       for b0 in c) * w0 / w1 for (w0, w1), c in model['BPH'][3].items()) / (6 * M * N)
    (a0, a1), full = _single(model['SPH'])
    sph = (F[full] / M + G[full] / N) * a0 / (2 * a1)
    sph += sum(sum(F[b0] * G[b1] + G[b0] * F[b1] for b0, b1 in c) * w0 / w1
               for (w0, w1), c in model['SPH'][2].items()) / (2 * M * N)
    (a0, a1), full = _single(model['DISOMY'])
    dis = (F[full] / M + G[full] / N) * a0 / (2 * a1)
    dis += sum(sum(F[b0] * G[b1] + G[b0] * F[b1] for b0, b1 in c) * w0 / w1
               for (w0, w1), c in model['DISOMY'][2].items()) / (2 * M * N)
    _, full = _single(model['MONOSOMY'])
    mono = (F[full] / M + G[full] / N) / 2
    return likelihoods_tuple(mono, dis, sph, bph)


def likelihoods_distant(E, model):
    """ DISTANT_ADMIXTURE_MODELS.py:215-256: E are the effective (float) joint
    frequencies, already normalised. """
    (a0, a1), full = _single(model['BPH'])
    bph = (a0 / a1) * E[full]
    bph += sum(sum(E[b0] * E[b1] for b0, b1 in c) * w0 / w1 for (w0, w1), c in model['BPH'][2].items())
    if model['BPH'].get(3):
        bph += sum(sum(E[b0] * sum(E[b1] * E[b2] for b1, b2 in c[b0]) for b0 in c) * w0 / w1
                   for (w0, w1), c in model['BPH'][3].items())
    (a0, a1), full = _single(model['SPH'])
    sph = (a0 / a1) * E[full]
    sph += sum(sum(E[b0] * E[b1] for b0, b1 in c) * w0 / w1 for (w0, w1), c in model['SPH'][2].items())
    (a0, a1), full = _single(model['DISOMY'])
    dis = (a0 / a1) * E[full]
    dis += sum(sum(E[b0] * E[b1] for b0, b1 in c) * w0 / w1 for (w0, w1), c in model['DISOMY'][2].items())
    _, full = _single(model['MONOSOMY'])
    return likelihoods_tuple(E[full], dis, sph, bph)


def effective_frequencies(counts_per_group, haplotypes_per_group, proportions):
    """ e(S) = sum_g proportion_g * count_g(S) / N_g in group order
    (effective_joint_frequency, DISTANT_ADMIXTURE_MODELS.py:136-139). """
    size = len(counts_per_group[0])
    E = [None] * size
    for s in range(1, size):
        E[s] = sum(p * c[s] / ng for p, c, ng in zip(proportions, counts_per_group, haplotypes_per_group))
    return E


# --------------------------------------------------------------------------
# a6 : closed forms for 2, 3 and 4 reads (normalised floats), via the same
# partition sums.  The reference hard-codes these polynomials
# (HOMOGENOUES_MODELS.py:227-270, RECENT...:249-312, DISTANT...:258-298); the
# values agree with its general formula to <= 2.3e-16 relative.
# --------------------------------------------------------------------------

def _unordered_pairs(n):
    full = (1 << n) - 1
    top = 1 << (n - 1)
    for a in range(top, full):          # a contains the last read, a != full
        yield a, full ^ a


def _unordered_triples(n):
    full = (1 << n) - 1
    top = 1 << (n - 1)
    for a in range(top, full):
        rest = full ^ a
        low = rest & -rest
        b = rest
        while True:                     # b runs over subsets of rest
            b = (b - 1) & rest
            if b == 0:
                break
            if b & low:                 # b owns rest's lowest read: {b, c} counted once
                yield a, b, rest ^ b


def closed_homogeneous(f, n):
    """ f = normalised frequencies; polynomials of HOMOGENOUES_MODELS.py:234-237,
    248-251, 265-268 written as sums over set partitions. """
    full = (1 << n) - 1
    pairs = list(_unordered_pairs(n))
    two = sum(f[a] * f[b] for a, b in pairs)
    two_sph = sum((2 ** a.bit_count() + 2 ** b.bit_count()) * f[a] * f[b] for a, b in pairs)
    three = sum(f[a] * f[b] * f[c] for a, b, c in _unordered_triples(n))
    mono = f[full]
    dis = (2 * f[full] + 2 * two) / 2 ** n
    sph = ((2 ** n + 1) * f[full] + two_sph) / 3 ** n
    bph = (3 * f[full] + 6 * two + 6 * three) / 3 ** n
    return likelihoods_tuple(mono, dis, sph, bph)


def closed_recent(f, g, n):
    """ Two-population polynomials of RECENT_ADMIXTURE_MODELS.py:258-261,
    274-277, 297-310 as partition sums. """
    full = (1 << n) - 1
    pairs = list(_unordered_pairs(n))
    cross = sum(f[a] * g[b] + g[a] * f[b] for a, b in pairs)
    cross_sph = sum((2 ** a.bit_count() + 2 ** b.bit_count()) * (f[a] * g[b] + g[a] * f[b]) for a, b in pairs)
    two_bph = sum(f[a] * f[b] + g[a] * g[b] + 2 * (f[a] * g[b] + g[a] * f[b]) for a, b in pairs)
    three = sum(f[a] * f[b] * g[c] + f[a] * g[b] * f[c] + g[a] * f[b] * f[c]
                + g[a] * g[b] * f[c] + g[a] * f[b] * g[c] + f[a] * g[b] * g[c] for a, b, c in _unordered_triples(n))
    s = f[full] + g[full]
    mono = s / 2
    dis = (s + cross) / 2 ** n
    sph = ((2 ** n + 1) * s + cross_sph) / (2 * 3 ** n)
    bph = (3 * s / 2 + two_bph + three) / 3 ** n
    return likelihoods_tuple(mono, dis, sph, bph)


# --------------------------------------------------------------------------
# model classes with the reference's dispatch (get_likelihoods)
# --------------------------------------------------------------------------

class Homogeneous:
    """ HOMOGENOUES_MODELS.homogeneous (:46-286). """
    def __init__(self, obs_tab, leg_tab, hap_tab_per_group, number_of_haplotypes_per_group, models_dict):
        if len(hap_tab_per_group) != 1:
            raise Exception('Error: The non-admixed models can not be applied when the HAP file includes multiple superpopulations/group2.')
        (group2, hap_tab), = hap_tab_per_group.items()
        if len(leg_tab) != len(hap_tab):
            raise Exception('Error: the number of SNPs in the LEGEND file differ from the number of SNPs in the HAP file.')
        self.models_dict = models_dict
        self.number_of_haplotypes = number_of_haplotypes_per_group[group2]
        self.hap_dict, frac = build_hap_dict(obs_tab, leg_tab, hap_tab, self.number_of_haplotypes)
        self.fraction_of_matches = {group2: frac}

    def joint_frequencies_combo(self, alleles, normalize):
        counts = joint_counts(read_masks(self.hap_dict, alleles))
        N = self.number_of_haplotypes
        return {s: (c / N if normalize else c) for s, c in enumerate(counts) if s}

    def likelihoods(self, alleles):
        F = joint_counts(read_masks(self.hap_dict, alleles))
        return likelihoods_homogeneous(F, self.number_of_haplotypes, self.models_dict[len(alleles)])

    def get_likelihoods(self, alleles):
        n = len(alleles)
        if n in (2, 3, 4):                                    # :277-283
            N = self.number_of_haplotypes
            F = joint_counts(read_masks(self.hap_dict, alleles))
            return closed_homogeneous([None] + [c / N for c in F[1:]], n)
        return self.likelihoods(alleles)                      # :284-285


class RecentAdmixture:
    """ RECENT_ADMIXTURE_MODELS.recent_admixture (:48-328). """
    def __init__(self, obs_tab, leg_tab, hap_tab_per_group, number_of_haplotypes_per_group, models_dict):
        if not all(len(leg_tab) == len(h) for h in hap_tab_per_group.values()):
            raise Exception('Error: the number of SNPs in the LEGEND file differ from the number of SNPs in the HAP file.')
        if len(hap_tab_per_group) != 2:
            raise Exception('Error: the reference panel does not contain two superpopulations/group2.')
        self.models_dict = models_dict
        self.group2_id = tuple(hap_tab_per_group.keys())
        self.number_of_haplotypes_per_group = number_of_haplotypes_per_group
        self.hap_dict_per_group, self.fraction_of_matches = {}, {}
        for g, hap_tab in hap_tab_per_group.items():
            self.hap_dict_per_group[g], self.fraction_of_matches[g] = \
                build_hap_dict(obs_tab, leg_tab, hap_tab, number_of_haplotypes_per_group[g])

    def _counts(self, alleles):
        return [joint_counts(read_masks(self.hap_dict_per_group[g], alleles)) for g in self.group2_id]

    def likelihoods(self, alleles):
        F, G = self._counts(alleles)
        M, N = (self.number_of_haplotypes_per_group[g] for g in self.group2_id)
        return likelihoods_recent(F, M, G, N, self.models_dict[len(alleles)])

    def get_likelihoods(self, alleles):
        n = len(alleles)
        if n in (2, 3, 4):                                    # :319-325
            F, G = self._counts(alleles)
            M, N = (self.number_of_haplotypes_per_group[g] for g in self.group2_id)
            return closed_recent([None] + [c / M for c in F[1:]], [None] + [c / N for c in G[1:]], n)
        return self.likelihoods(alleles)


class DistantAdmixture:
    """ DISTANT_ADMIXTURE_MODELS.distant_admixture (:48-316). """
    def __init__(self, obs_tab, leg_tab, hap_tab_per_group, number_of_haplotypes_per_group, models_dict, ancestral_makeup):
        if not all(len(leg_tab) == len(h) for h in hap_tab_per_group.values()):
            raise Exception('Error: the number of SNPs in the LEGEND file differ from the number of SNPs in the HAP file.')
        if sum(ancestral_makeup.values()) < 0.999:
            raise Exception('Error: ancestry proportions must sum to one.')
        for g in ancestral_makeup:
            if g not in hap_tab_per_group:
                raise Exception('Error: %s is missing from the reference panel.' % g)
        self.models_dict = models_dict
        self.number_of_haplotypes_per_group = number_of_haplotypes_per_group
        self.ancestral_makeup = ancestral_makeup
        self.hap_dict_per_group, self.fraction_of_matches = {}, {}
        for g, hap_tab in hap_tab_per_group.items():
            self.hap_dict_per_group[g], self.fraction_of_matches[g] = \
                build_hap_dict(obs_tab, leg_tab, hap_tab, number_of_haplotypes_per_group[g])

    def effective(self, alleles):
        groups = list(self.ancestral_makeup)                  # sum order of :138-139
        counts = [joint_counts(read_masks(self.hap_dict_per_group[g], alleles)) for g in groups]
        return effective_frequencies(counts, [self.number_of_haplotypes_per_group[g] for g in groups],
                                     [self.ancestral_makeup[g] for g in groups])

    def likelihoods(self, alleles):
        return likelihoods_distant(self.effective(alleles), self.models_dict[len(alleles)])

    def get_likelihoods(self, alleles):
        n = len(alleles)
        if n in (2, 3, 4):                                    # :305-311
            return closed_homogeneous(self.effective(alleles), n)
        if n == 5:                                            # :312-313 calls a method that does not exist
            raise AttributeError("'distant_admixture' object has no attribute 'likelihoods5'")
        return self.likelihoods(alleles)


# --------------------------------------------------------------------------
# a8 / a9 : draws and the bootstrap loop
# --------------------------------------------------------------------------

def effective_number_of_subsamples(num_of_reads, min_reads, max_reads, subsamples):
    """ ANEUPLOIDY_TEST.py:235-248. """
    if min_reads <= num_of_reads and num_of_reads > max_reads:
        return min(comb(num_of_reads, max_reads), subsamples)
    if min_reads <= num_of_reads <= max_reads:
        return min(num_of_reads, subsamples)
    return 0


def pick_read_ids(read_IDs, max_reads, rng=random):
    """ The draw of pick_reads (ANEUPLOIDY_TEST.py:226-233): min(R-1, max_reads)
    read ids without replacement, from the global MT19937 stream by default. """
    return rng.sample(read_IDs, min(len(read_IDs) - 1, max_reads))


class SupportingDictionaries:
    """ ANEUPLOIDY_TEST.supporting_dictionaries (:96-189): reads, overlaps,
    positions, per-group panel rows at observed positions and the read score. """
    def __init__(self, obs_tab, leg_tab, hap_tab_per_group, number_of_haplotypes_per_group, ancestral_makeup, min_HF):
        self.number_of_haplotypes_per_group = number_of_haplotypes_per_group
        self.min_HF = min_HF
        if type(ancestral_makeup) == dict:
            self.ancestral_makeup = {g: ancestral_makeup[g] for g in hap_tab_per_group}
        else:
            self.ancestral_makeup = {g: 1 / len(ancestral_makeup) for g in hap_tab_per_group}
        self.leg_dict = {l[1]: (l[2], l[3]) for l in leg_tab}
        self.reads = collections.defaultdict(list)
        self.overlaps = collections.defaultdict(list)
        for pos, read_id, base in obs_tab:                    # :118-139
            if base in self.leg_dict.get(pos, ()):
                self.reads[read_id].append((pos, base))
                self.overlaps[pos].append(read_id)
        self.positions = (*self.overlaps.keys(),)
        wanted = set(self.positions)
        self.hap_tab_per_group = {g: {leg[1]: hap for leg, hap in zip(leg_tab, hap_tab) if leg[1] in wanted}
                                  for g, hap_tab in hap_tab_per_group.items()}       # :141-149
        self.score = self._scores()

    def _scores(self):
        """ build_score_dict, ANEUPLOIDY_TEST.py:151-189. """
        from itertools import product
        lo, hi = self.min_HF, 1 - self.min_HF
        alpha = {g: self.ancestral_makeup[g] / n for g, n in self.number_of_haplotypes_per_group.items()}
        ones = {g: (1 << n) - 1 for g, n in self.number_of_haplotypes_per_group.items()}
        score = {}
        for read_id, alleles in self.reads.items():
            rows = collections.defaultdict(list)
            for pos, _base in alleles:
                freq = sum(a * self.hap_tab_per_group[g][pos].bit_count() for g, a in alpha.items())
                if 0.01 <= freq <= 0.99:
                    for g, b in ones.items():
                        h = self.hap_tab_per_group[g][pos]
                        rows[g].append((h, h ^ b))
            if rows:
                # Reference quirk kept on purpose: :181-182 builds one *generator* per group inside a
                # list comprehension, so `a` is read only when :184 consumes them, after the loop
                # variable has reached the LAST group -- every group is weighted by the last alpha.
                a = list(alpha.values())[-1]
                per_group = []
                for g in alpha:
                    vals = []
                    for choice in product(*rows[g]):
                        m = choice[0]
                        for r in choice[1:]:
                            m &= r
                        vals.append(a * m.bit_count())
                    per_group.append(vals)
                score[read_id] = sum(lo <= sum(t) <= hi for t in zip(*per_group))
            else:
                score[read_id] = 0
        return score


def build_gw_dict(obs, window_size, offset, min_reads, max_reads, min_score):
    """ Windows -> read ids, fixed or adaptive size (ANEUPLOIDY_TEST.py:191-224),
    including its quirks: the last window is never flushed, empty windows are kept. """
    max_dist, max_win_size, initial = 100000, 350000, 10000
    adaptive, window_size = (False, int(window_size)) if window_size else (True, initial)
    offset = int(offset)
    assert len(obs.positions), 'error: the observation table is empty.'
    first, last = obs.positions[0] + offset, obs.positions[-1] + window_size
    a, b = first, first + window_size
    in_window = set()
    gw = {}
    for pos, overlapping in obs.overlaps.items():
        if pos < first:
            continue
        while b < last:
            if a <= pos < b:
                in_window.update(r for r in overlapping if min_score <= obs.score[r])
                break
            elif adaptive and 0 < len(in_window) < min_reads and b - pos <= max_dist and b - a <= max_win_size:
                b += 10000
            else:
                gw[a, b - 1] = tuple(in_window)
                a, b = b, b + window_size
                in_window = set()
    return gw


def make_examiner(obs_tab, leg_tab, hap_tab, number_of_haplotypes, ancestral_makeup, models_dict):
    """ Model selection of ANEUPLOIDY_TEST.py:263-273. """
    if type(ancestral_makeup) == set and len(ancestral_makeup) == 1:
        return Homogeneous(obs_tab, leg_tab, hap_tab, number_of_haplotypes, models_dict)
    if type(ancestral_makeup) == set and len(ancestral_makeup) == 2:
        return RecentAdmixture(obs_tab, leg_tab, hap_tab, number_of_haplotypes, models_dict)
    if type(ancestral_makeup) == dict:
        return DistantAdmixture(obs_tab, leg_tab, hap_tab, number_of_haplotypes, models_dict, ancestral_makeup)
    raise Exception('error: unsupported ancestral-makeup.')


def bootstrap(obs_tab, leg_tab, hap_tab, number_of_haplotypes, ancestral_makeup, models_dict,
              window_size, subsamples, offset, min_reads, max_reads, min_score, min_HF,
              replay=None, record=None, rng=random):
    """ ANEUPLOIDY_TEST.bootstrap (:250-287).  `replay`: {window: [tuple of read
    ids per draw]} substitutes recorded draws for random.sample and, when given,
    also fixes the windows' read tuples; `record`: dict filled with the draws made. """
    obs = SupportingDictionaries(obs_tab, leg_tab, hap_tab, number_of_haplotypes, ancestral_makeup, min_HF)
    genomic_windows = build_gw_dict(obs, window_size, offset, min_reads, max_reads, min_score)
    examine = make_examiner(obs_tab, leg_tab, hap_tab, number_of_haplotypes, ancestral_makeup, models_dict)
    likelihoods = {}
    for window, read_IDs in genomic_windows.items():
        effN = effective_number_of_subsamples(len(read_IDs), min_reads, max_reads, subsamples)
        if not effN:
            continue
        cache = []
        for i in range(effN):
            ids = replay[window][i] if replay is not None else pick_read_ids(read_IDs, max_reads, rng)
            if record is not None:
                record.setdefault(window, []).append(tuple(ids))
            cache.append(examine.get_likelihoods(tuple(obs.reads[r] for r in ids)))
        likelihoods[window] = tuple(cache)
    return likelihoods, genomic_windows, examine.fraction_of_matches


def evaluate_draws(examine, reads, draws):
    """ The inner loop of bootstrap (:282-286) on explicit draws: {window: [ids,...]}
    -> {window: tuple(likelihoods_tuple)}.  This is the unit bench.py times as
    the CPU baseline. """
    return {w: tuple(examine.get_likelihoods(tuple(reads[r] for r in ids)) for ids in ds) for w, ds in draws.items()}


# --------------------------------------------------------------------------
# a10 : LLRs and their reductions
# --------------------------------------------------------------------------

def LLR(y, x):
    """ log(y/x) with the reference's sentinels (ANEUPLOIDY_TEST.py:78-90). """
    if x and y:
        return log(y / x)
    if x and not y:
        return -1.23456789
    if not x and y:
        return +1.23456789
    return 0


def mean_and_var(x):
    """ ANEUPLOIDY_TEST.py:51-56 (sample variance; exact-rational internally). """
    cache = tuple(x)
    m = mean(cache)
    return m, variance(cache, xbar=m)


def mean_and_std(x):
    """ ANEUPLOIDY_TEST.py:58-63 (population standard deviation). """
    cache = tuple(x)
    m = mean(cache)
    return m, pstdev(cache, mu=m)


def mean_and_std_of_mean_of_rnd_var(A):
    """ ANEUPLOIDY_TEST.py:65-76; N is the first window's draw count and zip()
    truncates to the shortest window, as in the reference. """
    if type(A) == dict:
        A = tuple(tuple(i) for i in A.values())
    M, N = len(A), len(A[0])
    mu = sum(sum(row) / N for row in A)
    std = (sum((sum(col) - mu) ** 2 for col in zip(*A)) / (N - 1)) ** .5 / M
    return mu / M, std


def statistics(likelihoods, genomic_windows):
    """ ANEUPLOIDY_TEST.statistics (:289-325). """
    if not likelihoods:
        return {'num_of_windows': 0}
    window_size_mean, window_size_std = mean_and_std((j - i + 1 for (i, j) in likelihoods))
    reads_mean, reads_std = mean_and_std((len(ids) for w, ids in genomic_windows.items() if w in likelihoods))
    LLRs = {(i, j): {w: tuple(LLR(getattr(l, i), getattr(l, j)) for l in ls) for w, ls in likelihoods.items()}
            for i, j in LLR_PAIRS}
    per_window = {pair: {w: mean_and_var(v) for w, v in LLRs[pair].items()} for pair in LLR_PAIRS}
    chrom = {}
    for pair in LLR_PAIRS:
        m, s = mean_and_std_of_mean_of_rnd_var(LLRs[pair])
        chrom[pair] = {'mean_of_mean': m, 'std_of_mean': s,
                       'fraction_of_negative_LLRs': mean(mv[0] < 0 for mv in per_window[pair].values())}
    return {'num_of_windows': len(likelihoods), 'reads_mean': reads_mean, 'reads_std': reads_std,
            'window_size_mean': window_size_mean, 'window_size_std': window_size_std,
            'LLRs_per_genomic_window': per_window, 'LLRs_per_chromosome': chrom}


# --------------------------------------------------------------------------
# SURVEY 8(f) rank 4 : consumers of the per-window LLRs (binning, crossovers)
# --------------------------------------------------------------------------

CHR_LENGTHS = dict(zip(['chr%s' % c for c in list(range(1, 23)) + ['X', 'Y']],      # hg38, PLOT_PANEL.py:20-28
                       (248956422, 242193529, 198295559, 190214555, 181538259, 170805979, 159345973, 145138636,
                        138394717, 133797422, 135086622, 133275309, 114364328, 107043718, 101991189, 90338345,
                        83257441, 80373285, 58617616, 64444167, 46709983, 50818468, 156040895, 57227415)))


def bin_genomic_windows(windows, chr_id, num_of_bins):
    """ PLOT_PANEL.py:96-119 (= BALANCED_ROC_CURVE.py:106-129).  Bins are keyed by their fractional borders
    (i/num_of_bins, (i+1)/num_of_bins); the value is the half-open run (j, k) of window indices whose midpoints
    fall in the bin, or None.  Kept quirks: the chromosome's last window never enters a bin (the final run ends
    at its index), and a bin is only closed when a later window leaves it. """
    size = CHR_LENGTHS[chr_id] / num_of_bins
    mids = [(a + b) / 2 for a, b in windows]
    name = lambda b: (b / num_of_bins, (b + 1) / num_of_bins)
    out, i = {}, 0
    for i in range(num_of_bins):                      # empty bins ahead of the first window (:102-105)
        if mids[0] < (i + 1) * size:
            break
        out[name(i)] = None
    start = 0
    for k, m in enumerate(mids):                      # (:107-114)
        if size * i <= m < size * (i + 1):
            continue
        out[name(i)] = (start, k)
        start = k
        for i in range(i + 1, num_of_bins):
            if m < (i + 1) * size:
                break
            out[name(i)] = None
    for i in range(i, num_of_bins):                   # the open bin and the empty ones behind it (:116-118)
        out[name(i)] = (start, k) if start != k else None
        start = k
    return out


def binning(LLRs_per_window, info, num_of_bins, rounding='plot'):
    """ PLOT_PANEL.py:121-144 / BALANCED_ROC_CURVE.py:131-154: bin borders, mean LLR and standard error per bin.
    rounding='plot' sums before dividing by N (PLOT_PANEL.py:44), 'roc' divides each window's sum (BALANCED_ROC_CURVE.py:65). """
    bins = bin_genomic_windows([*LLRs_per_window], info['chr_id'], num_of_bins)
    rows = [*LLRs_per_window.values()]
    Y, E = [], []
    for run in bins.values():
        if run:
            A = rows[run[0]:run[1]]
            M, N = len(A), len(A[0])
            mu = sum(sum(r) for r in A) / N if rounding == 'plot' else sum(sum(r) / N for r in A)
            std = (sum((sum(c) - mu) ** 2 for c in zip(*A)) / (N - 1)) ** .5 / M
            Y.append(mu / M)
            E.append(std)
        else:
            Y.append(None)
            E.append(None)
    return [*bins], Y, E


def detect_crossovers(genomic_windows, mean_of_LLRs, variance_of_LLRs, z_score=1.96, lookahead=20, recover=False):
    """ PLOT_PANEL.detect_crossovers (:146-184) and, with recover=True, detect_crossovers_v2 (:186-245).
    Walks the cumulative LLR curve; an extremum becomes a crossover once the curve has moved away from it by
    z_score standard deviations (of the cumulative variance) on BOTH sides, the right side at least `lookahead`
    windows later, the left side searched backwards from `lookahead` windows before the extremum down to the
    previous crossover.  v2 also inserts the extremum that must lie between two crossovers of the same kind. """
    xs = [0.5 * (a + b) for a, b in genomic_windows]
    ys, vs = [], []
    for m, v in zip(mean_of_LLRs, variance_of_LLRs):           # itertools.accumulate: running float sums
        ys.append(m if not ys else ys[-1] + m)
        vs.append(v if not vs else vs[-1] + v)
    pts = list(zip(xs, ys, vs))
    found = {}
    hi = lo = None            # candidate maximum / minimum as (index, x, y, v)
    last, last_kind = 0, 0

    def skipped(pick, i0, i1):                                  # recover_skipped_extremum (:199-209)
        M0, M2, V0, V2 = ys[i0], ys[i1], vs[i0], vs[i1]
        cands = [((M1 - M0) / (V1 - V0) ** .5 - (M2 - M1) / (V2 - V1) ** .5, X1, M1, V1)
                 for X1, M1, V1 in pts[i0 + 5:i1 - 4] if V0 != V1 != V2]
        Z1, X1, M1, V1 = pick(cands, key=lambda t: t[0])
        return X1, min(abs(M2 - M1) / (V2 - V1) ** .5, abs(M0 - M1) / (V1 - V0) ** .5)

    for idx, (x, y, v) in enumerate(pts):
        if hi is None or y > hi[2]:
            hi = (idx, x, y, v)
        if lo is None or y < lo[2]:
            lo = (idx, x, y, v)
        for kind in (+1, -1):                                   # +1: maximum (:161-168), -1: minimum (:170-177)
            c = hi if kind > 0 else lo
            if c is None:
                continue
            ci, cx, cy, cv = c
            if not (0 < kind * (cy - y) - z_score * (v - cv) ** .5 and idx - ci >= lookahead):
                continue
            for x2, y2, v2 in pts[max(ci - lookahead, last):last:-1]:
                if 0 < kind * (cy - y2) - z_score * (cv - v2) ** .5:
                    if recover and last_kind == kind:
                        rx, rk = skipped(min if kind > 0 else max, last, ci)
                        found[rx] = rk
                    found[cx] = min(kind * (cy - y2) / (cv - v2) ** .5, kind * (cy - y) / (v - cv) ** .5)
                    hi = lo = None
                    last, last_kind = ci, kind
                    break
    return found


# --------------------------------------------------------------------------
# exact-rational cross-check of the closed combinatorial forms
# --------------------------------------------------------------------------

def likelihoods_exact(F, N, n):
    """ Homogeneous likelihoods as exact Fractions from the partition sums; used
    by tests to bound the float error of both the table-driven oracle and the
    CUDA path. """
    full = (1 << n) - 1
    f = [None] + [Fraction(c, N) for c in F[1:]]
    two = sum(f[a] * f[b] for a, b in _unordered_pairs(n))
    two_sph = sum((2 ** a.bit_count() + 2 ** b.bit_count()) * f[a] * f[b] for a, b in _unordered_pairs(n))
    three = sum(f[a] * f[b] * f[c] for a, b, c in _unordered_triples(n))
    return likelihoods_tuple(f[full], (2 * f[full] + 2 * two) / 2 ** n,
                             ((2 ** n + 1) * f[full] + two_sph) / 3 ** n,
                             (3 * f[full] + 6 * two + 6 * three) / 3 ** n)
