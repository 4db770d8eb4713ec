"""
Drop-in for the reference module RECENT_ADMIXTURE_MODELS: `recent_admixture` with the same
constructor, methods and result record, computed on a B200 (see models.py).
Importable flat (`import RECENT_ADMIXTURE_MODELS` with this directory on sys.path) or as
`ldpgta_b200.RECENT_ADMIXTURE_MODELS`.
"""
import os as _os
import sys as _sys

try:
    from ldpgta_b200 import models as _models
except ImportError:                      # flat import: make the repository root importable
    _sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
    from ldpgta_b200 import models as _models

leg_tuple = _models.leg_tuple
obs_tuple = _models.obs_tuple
popcount = _models.popcount

# The reference pickles likelihoods_tuple instances under the bare module name 'RECENT_ADMIXTURE_MODELS' (*.LLR.p files).  The class
# lives in models.py (one object however this file was imported); the bare name resolves to this module unless the
# caller has already put a module of that name (e.g. the reference's own) into sys.modules.
likelihoods_tuple = _models.TUPLES['RECENT_ADMIXTURE_MODELS']
_sys.modules.setdefault('RECENT_ADMIXTURE_MODELS', _sys.modules[__name__])
recent_admixture = _models.recent_admixture
