"""
ldpgta_b200 -- B200-native (sm_100a) implementation of LD-PGTA's windowed
haplotype-likelihood engine, behind the reference's Python API.

Modules named after the reference's scripts (HOMOGENOUES_MODELS, RECENT_ADMIXTURE_MODELS,
DISTANT_ADMIXTURE_MODELS, ANEUPLOIDY_TEST) live here too, so putting this
directory on sys.path makes `from ANEUPLOIDY_TEST import aneuploidy_test`
resolve to the GPU implementation.
"""
from ._native import (DISTANT_ADMIXTURE, HOMOGENEOUS, RECENT_ADMIXTURE, LdpgtaError, LIB_PATH, PinnedArray, load)
from .engine import LLR_PAIRS, Engine, ResidentPanel, finish_chromosome, pack_ints
from .models import ImplicitModels, distant_admixture, homogeneous, likelihoods_tuple, recent_admixture
from .driver import (aneuploidy_test, bootstrap, build_gw_dict, effective_number_of_subsamples, evaluate_draws,
                     pick_reads, plan_draw_indices, plan_draws, statistics, supporting_dictionaries)

# registers the reference's module names for pickling (see models.TUPLES)
from . import HOMOGENOUES_MODELS as _hm, RECENT_ADMIXTURE_MODELS as _rm, DISTANT_ADMIXTURE_MODELS as _dm  # noqa: E402,F401

__all__ = ['Engine', 'homogeneous', 'recent_admixture', 'distant_admixture', 'bootstrap', 'aneuploidy_test',
           'statistics', 'likelihoods_tuple', 'ImplicitModels', 'LdpgtaError', 'pack_ints', 'finish_chromosome',
           'ResidentPanel', 'HOMOGENEOUS', 'RECENT_ADMIXTURE', 'DISTANT_ADMIXTURE', 'LLR_PAIRS', 'load', 'LIB_PATH', 'PinnedArray']
