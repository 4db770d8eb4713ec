#!/usr/bin/env python3
"""
Drop-in for the reference driver ANEUPLOIDY_TEST.py (functions, CLI and pickle
formats; see driver.py).  Usable flat (`python ANEUPLOIDY_TEST.py OBS LEG HAP EUR`,
`from ANEUPLOIDY_TEST import aneuploidy_test`) or as `ldpgta_b200.ANEUPLOIDY_TEST`.
"""
import os as _os
import sys as _sys

try:
    from ldpgta_b200 import driver as _driver
except ImportError:
    _sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
    from ldpgta_b200 import driver as _driver

from ldpgta_b200.HOMOGENOUES_MODELS import homogeneous
from ldpgta_b200.RECENT_ADMIXTURE_MODELS import recent_admixture
from ldpgta_b200.DISTANT_ADMIXTURE_MODELS import distant_admixture

leg_tuple, obs_tuple, comb_tuple = _driver.leg_tuple, _driver.obs_tuple, _driver.comb_tuple
likelihoods_tuple = _driver.likelihoods_tuple
popcount = _driver.popcount
mean_and_var = _driver.mean_and_var
mean_and_std = _driver.mean_and_std
mean_and_std_of_mean_of_rnd_var = _driver.mean_and_std_of_mean_of_rnd_var
LLR = _driver.LLR
invert = _driver.invert
supporting_dictionaries = _driver.supporting_dictionaries
build_gw_dict = _driver.build_gw_dict
pick_reads = _driver.pick_reads
effective_number_of_subsamples = _driver.effective_number_of_subsamples
bootstrap = _driver.bootstrap
statistics = _driver.statistics
print_summary = _driver.print_summary
save_results = _driver.save_results
aneuploidy_test = _driver.aneuploidy_test

if __name__ == "__main__":
    _driver.main()
    _sys.exit(0)
