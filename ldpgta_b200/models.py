"""
The three likelihood-model classes of LD-PGTA with the reference's constructor
signatures, attributes, methods and error behaviour, backed by the CUDA engine:

    homogeneous        HOMOGENOUES_MODELS.py:46-286
    recent_admixture   RECENT_ADMIXTURE_MODELS.py:48-328
    distant_admixture  DISTANT_ADMIXTURE_MODELS.py:48-316

The arithmetic (AND/popcount over read subsets and the monosomy/disomy/SPH/BPH
polynomials) runs in libldpgta.so; these classes only translate (pos, base)
alleles into rows of the device-resident panel.  `models_dict` is accepted and
kept for interface compatibility -- the device evaluates the partition
polynomials directly and needs no model tables -- but, like the reference, a
draw size missing from it is a KeyError.
"""
import collections
from functools import reduce
from operator import and_

import numpy as np

from . import _native as N
from .engine import Engine, pack_ints


def _namedtuple(module):
    t = collections.namedtuple('likelihoods_tuple', ('monosomy', 'disomy', 'SPH', 'BPH'))
    t.__module__ = module
    return t


# One result-record class per reference module, defined ONCE: the reference pickles likelihoods_tuple instances under the
# bare module names (`HOMOGENOUES_MODELS.likelihoods_tuple` in *.LLR.p), and its readers (PLOT_PANEL, BALANCED_ROC_CURVE)
# unpickle them by that path.  The modules of those names in this package re-export these very objects, whichever way
# they were imported (flat or as ldpgta_b200.X), so what gets pickled never depends on the import order.
TUPLES = {name: _namedtuple(name) for name in ('HOMOGENOUES_MODELS', 'RECENT_ADMIXTURE_MODELS', 'DISTANT_ADMIXTURE_MODELS')}
likelihoods_tuple = TUPLES['HOMOGENOUES_MODELS']
leg_tuple = collections.namedtuple('leg_tuple', ('chr_id', 'pos', 'ref', 'alt'))
obs_tuple = collections.namedtuple('obs_tuple', ('pos', 'read_id', 'base'))


class ImplicitModels(dict):
    """ Stand-in for the MODELS*.p tables when none is supplied: every draw size
    the device supports is "present".  (The reference ships MODELS12/16/18.p;
    ANEUPLOIDY_TEST.py:379-380.) """
    def __init__(self, n_max=N.MAX_READS_PER_DRAW):
        super().__init__({n: None for n in range(2, n_max + 1)})


def popcount(x):
    """ Same contract as the reference's module-level popcount (HOMOGENOUES_MODELS.py:323-330). """
    return x.bit_count()


class _PanelModel:
    """ Shared machinery: allele table, device residency, allele -> row translation. """
    kind = None
    tuple_type = likelihoods_tuple

    def _load_panel(self, obs_tab, leg_tab, hap_tab_per_group, number_of_haplotypes_per_group, device, panel=None):
        groups = list(hap_tab_per_group)
        self._groups = groups
        self._nhap = {g: number_of_haplotypes_per_group[g] for g in groups}
        # build_hap_dict (HOMOGENOUES_MODELS.py:76-100): legend lookup, last duplicate position wins (:84)
        # (only observed positions are indexed: the panel has ~20x more SNPs than a low-coverage sample touches)
        if panel is not None:
            if panel.groups != groups or panel.n_hap != [self._nhap[g] for g in groups] or panel.n_snps != len(leg_tab):
                raise ValueError('the resident panel does not match the LEGEND/HAP tables of this model')
            by_pos = panel.by_pos
        else:
            wanted = {row[0] for row in obs_tab}
            by_pos = {leg[1]: (leg[2], leg[3], i) for i, leg in enumerate(leg_tab) if leg[1] in wanted}
        self._allele_index = {}
        rows, comp = [], []
        mismatches = 0
        for pos, _read_id, base in obs_tab:
            entry = by_pos.get(pos)
            if entry is not None and base == entry[1]:
                flag = 0
            elif entry is not None and base == entry[0]:
                flag = 1
            else:
                mismatches += 1
                continue
            key = (pos, base)
            if key not in self._allele_index:
                self._allele_index[key] = len(rows)
                rows.append(entry[2])
                comp.append(flag)
        frac = 1 - mismatches / len(obs_tab)
        self.fraction_of_matches = {g: frac for g in groups}
        self._legend_rows = rows
        self._complement = np.asarray(comp, dtype=np.uint8)
        self._hap_tabs = hap_tab_per_group
        self._hap_dicts = {}
        self._engine = Engine([self._nhap[g] for g in groups], device=device)
        if panel is not None:
            self._engine.gather_allele_rows(panel, rows, self._complement)          # device gather, no re-packing
        else:
            for gi, g in enumerate(groups):
                hap_tab = hap_tab_per_group[g]
                self._engine.upload_allele_rows(gi, pack_ints([hap_tab[i] for i in rows], self._nhap[g]), self._complement)

    def _hap_dict_of(self, group2):
        """ The reference's hap_dict as Python ints, materialised on first use. """
        if group2 not in self._hap_dicts:
            ones = (1 << self._nhap[group2]) - 1
            hap_tab = self._hap_tabs[group2]
            self._hap_dicts[group2] = {key: (hap_tab[self._legend_rows[i]] ^ ones if self._complement[i] else hap_tab[self._legend_rows[i]])
                                       for key, i in self._allele_index.items()}
        return self._hap_dicts[group2]

    def build_hap_dict(self, obs_tab, leg_tab, hap_tab, number_of_haplotypes):
        """ HOMOGENOUES_MODELS.py:76-100, kept as a host utility. """
        by_pos = {}
        for leg, hap in zip(leg_tab, hap_tab):
            by_pos[leg[1]] = (leg[2], leg[3], hap)
        ones = (1 << number_of_haplotypes) - 1
        hap_dict, mismatches = {}, 0
        for pos, _rid, base in obs_tab:
            entry = by_pos.get(pos)
            if entry is not None and base == entry[1]:
                hap_dict[(pos, base)] = entry[2]
            elif entry is not None and base == entry[0]:
                hap_dict[(pos, base)] = entry[2] ^ ones
            else:
                mismatches += 1
        return hap_dict, 1 - mismatches / len(obs_tab)

    def _rows_of(self, alleles):
        """ alleles -> one list of allele-row indices per haplotype (the translation
        build_intrenal_hap_dict does, HOMOGENOUES_MODELS.py:102-127; the ANDs run on device). """
        index = self._allele_index
        lists = []
        for haplotype in alleles:
            if type(haplotype[0]) == tuple:
                lists.append([index[a] for a in haplotype])
            elif type(haplotype[0]) == int:
                lists.append([index[haplotype]])
            else:
                raise Exception('error: joint_frequencies only accepts alleles and tuple/list of alleles.')
        return lists

    def _internal(self, alleles, group2):
        hap_dict = self._hap_dict_of(group2)
        internal = {}
        for i, haplotype in enumerate(alleles):
            if type(haplotype[0]) == tuple:
                internal[1 << i] = reduce(and_, (hap_dict[a] for a in haplotype))
            elif type(haplotype[0]) == int:
                internal[1 << i] = hap_dict[haplotype]
            else:
                raise Exception('error: joint_frequencies only accepts alleles and tuple/list of alleles.')
        return internal

    def _check_table(self, n):
        if self.models_dict is not None and n not in self.models_dict:
            raise KeyError(n)

    def _proportions(self):
        return None

    def _evaluate(self, alleles, general, want_counts=False):
        return self._engine.get_likelihoods(self.kind, self._rows_of(alleles), proportions=self._proportions(),
                                            general=general, want_counts=want_counts)

    def _counts(self, alleles):
        """ Integer joint frequencies of every non-empty subset, per group: uint32[G, 2^n]. """
        n = len(alleles)
        if n == 1:
            rows = self._rows_of(alleles)
            # a single haplotype has a single subset; evaluate it with a duplicated read
            _, c = self._engine.get_likelihoods(N.DISTANT_ADMIXTURE if self._engine.G > 1 else N.HOMOGENEOUS, rows + rows,
                                                proportions=[1.0 / self._engine.G] * self._engine.G, want_counts=True)
            return c[:, :2]
        _, c = self._engine.get_likelihoods(N.DISTANT_ADMIXTURE if self._engine.G > 1 else N.HOMOGENEOUS,
                                            self._rows_of(alleles), proportions=[1.0 / self._engine.G] * self._engine.G,
                                            want_counts=True)
        return c

    # -- the reference's public methods --------------------------------------
    def likelihoods(self, alleles):
        """ General partition polynomials for any number of reads (HOMOGENOUES_MODELS.py:185-225). """
        self._check_table(len(alleles))
        return self.tuple_type(*self._evaluate(alleles, general=True))

    def _closed(self, alleles, n):
        if len(alleles) != n:
            raise ValueError('likelihoods%d expects %d alleles/haplotypes' % (n, n))
        return self.tuple_type(*self._evaluate(alleles, general=False))

    def likelihoods2(self, alleles):
        return self._closed(alleles, 2)

    def likelihoods3(self, alleles):
        return self._closed(alleles, 3)

    def likelihoods4(self, alleles):
        return self._closed(alleles, 4)

    def get_likelihoods(self, alleles):
        """ Dispatch of HOMOGENOUES_MODELS.py:272-286: closed forms for 2..4 reads, the general
        polynomials otherwise. """
        n = len(alleles)
        if n in (2, 3, 4):
            return self.tuple_type(*self._evaluate(alleles, general=False))
        return self.likelihoods(alleles)

    def close(self):
        self._engine.close()


class homogeneous(_PanelModel):
    """ Single-population model; same constructor and checks as HOMOGENOUES_MODELS.py:52-73. """
    kind = N.HOMOGENEOUS
    tuple_type = TUPLES['HOMOGENOUES_MODELS']

    def __init__(self, obs_tab, leg_tab, hap_tab_per_group, number_of_haplotypes_per_group, models_dict, device=0, panel=None):
        print('Initializing the algorithm for non-admixtures:')
        if len(hap_tab_per_group) != 1:
            raise Exception('Error: The non-admixed models can not be applied when the HAP file includes multiple superpopulations/group2.')
        self.models_dict = models_dict
        group2 = [*hap_tab_per_group.keys()].pop()
        hap_tab = hap_tab_per_group[group2]
        self.number_of_haplotypes = number_of_haplotypes_per_group[group2]
        if len(leg_tab) != len(hap_tab):
            raise Exception('Error: the number of SNPs in the LEGEND file differ from the number of SNPs in the HAP file.')
        self._group2 = group2
        self._load_panel(obs_tab, leg_tab, {group2: hap_tab}, {group2: self.number_of_haplotypes}, device, panel)
        print('%.2f%% of the observed alleles matched the reference panel.' % (100 * self.fraction_of_matches[group2]))

    @property
    def hap_dict(self):
        return self._hap_dict_of(self._group2)

    def build_intrenal_hap_dict(self, alleles):
        return self._internal(alleles, self._group2)

    def joint_frequencies_combo(self, alleles, normalize):
        """ {subset bitmask: joint frequency} for every non-empty subset (HOMOGENOUES_MODELS.py:129-163). """
        counts = self._counts(alleles)[0]
        Nh = self.number_of_haplotypes
        return {s: (int(counts[s]) / Nh if normalize else int(counts[s])) for s in range(1, len(counts))}

    joint_frequencies_redundant = joint_frequencies_combo


class recent_admixture(_PanelModel):
    """ Two populations, each parent from one of them (RECENT_ADMIXTURE_MODELS.py:54-80). """
    kind = N.RECENT_ADMIXTURE
    tuple_type = TUPLES['RECENT_ADMIXTURE_MODELS']

    def __init__(self, obs_tab, leg_tab, hap_tab_per_group, number_of_haplotypes_per_group, models_dict, device=0, panel=None):
        print('Initializing the algorithm for recent admixtures:')
        if not all(len(leg_tab) == len(hap_tab) for hap_tab in hap_tab_per_group.values()):
            raise Exception('Error: the number of SNPs in the LEGEND file differ from the number of SNPs in the HAP file.')
        if len(hap_tab_per_group) != 2:
            raise Exception('Error: the reference panel does not contain two superpopulations/group2.')
        self.models_dict = models_dict
        self.group2_id = tuple(hap_tab_per_group.keys())
        self.number_of_haplotypes_per_group = number_of_haplotypes_per_group
        self._load_panel(obs_tab, leg_tab, hap_tab_per_group, number_of_haplotypes_per_group, device, panel)
        for group2 in self.group2_id:
            print('%.2f%% of the observed alleles matched the reference panel of %s.' % (100 * self.fraction_of_matches[group2], group2))

    @property
    def hap_dict_per_group(self):
        return {g: self._hap_dict_of(g) for g in self._groups}

    def build_intrenal_hap_dict(self, alleles, group2):
        return self._internal(alleles, group2)

    def joint_frequencies_combo(self, alleles, group2, normalize):
        counts = self._counts(alleles)[self._groups.index(group2)]
        Nh = self.number_of_haplotypes_per_group[group2]
        return {s: (int(counts[s]) / Nh if normalize else int(counts[s])) for s in range(1, len(counts))}

    joint_frequencies_redundant = joint_frequencies_combo


class distant_admixture(_PanelModel):
    """ Any number of populations with ancestral proportions (DISTANT_ADMIXTURE_MODELS.py:53-79). """
    kind = N.DISTANT_ADMIXTURE
    tuple_type = TUPLES['DISTANT_ADMIXTURE_MODELS']

    def __init__(self, obs_tab, leg_tab, hap_tab_per_group, number_of_haplotypes_per_group, models_dict, ancestral_makeup, device=0, panel=None):
        print('Initializing the algorithm for distant admixtures:')
        if not all(len(leg_tab) == len(hap_tab) for hap_tab in hap_tab_per_group.values()):
            raise Exception('Error: the number of SNPs in the LEGEND file differ from the number of SNPs in the HAP file.')
        if sum(ancestral_makeup.values()) < 0.999:
            raise Exception('Error: ancestry proportions must sum to one.')
        self.models_dict = models_dict
        self.number_of_haplotypes_per_group = number_of_haplotypes_per_group
        self.ancestral_makeup = ancestral_makeup
        for group2 in ancestral_makeup:
            if group2 not in hap_tab_per_group:
                raise Exception('Error: %s is missing from the reference panel.' % group2)
        # device groups follow the ancestral_makeup order: that is the summation order of
        # effective_joint_frequency (:138-139); panel groups outside the makeup carry no weight
        ordered = {g: hap_tab_per_group[g] for g in ancestral_makeup}
        # a ResidentPanel shared by many samples must hold its groups in that same order (engine.ResidentPanel(leg, {g: hap[g] for g in makeup}, ...))
        self._load_panel(obs_tab, leg_tab, ordered, number_of_haplotypes_per_group, device, panel)
        self.fraction_of_matches = {g: self.fraction_of_matches[next(iter(ordered))] for g in hap_tab_per_group}
        self._hap_tabs = hap_tab_per_group
        self._nhap = {g: number_of_haplotypes_per_group[g] for g in hap_tab_per_group}
        for group2 in hap_tab_per_group:
            print('%.2f%% of the observed alleles matched the reference panel of %s.' % (100 * self.fraction_of_matches[group2], group2))

    def _proportions(self):
        return [self.ancestral_makeup[g] for g in self._groups]

    @property
    def hap_dict_per_group(self):
        return {g: self._hap_dict_of(g) for g in self._hap_tabs}

    def build_intrenal_hap_dict(self, alleles, group2):
        return self._internal(alleles, group2)

    def effective_joint_frequency(self, hap):
        """ DISTANT_ADMIXTURE_MODELS.py:136-139 on Python ints. """
        return sum(proportion * popcount(hap[group2]) / self.number_of_haplotypes_per_group[group2]
                   for group2, proportion in self.ancestral_makeup.items())

    def joint_frequencies_combo(self, alleles):
        counts = self._counts(alleles)
        return {s: sum(p * int(counts[gi][s]) / self.number_of_haplotypes_per_group[g]
                       for gi, (g, p) in enumerate(self.ancestral_makeup.items())) for s in range(1, counts.shape[1])}

    joint_frequencies_redundant = joint_frequencies_combo

    def get_likelihoods(self, alleles):
        """ DISTANT_ADMIXTURE_MODELS.py:300-316, including its missing likelihoods5 (:312-313). """
        if len(alleles) == 5:
            raise AttributeError("'distant_admixture' object has no attribute 'likelihoods5'")
        return super().get_likelihoods(alleles)
