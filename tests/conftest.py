"""Shared pytest configuration: the `gpu` marker, import paths and golden-fixture loaders."""
import bz2, os, pickle, sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, 'tests', 'golden')
for p in (ROOT, os.path.join(ROOT, 'oracle')):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')
    # a fresh checkout has no libldpgta.so (build artefact): build it once when nvcc is there, otherwise the tests that
    # need it skip (tests/test_host_cpu.py) and the rest of the CPU suite still runs
    import shutil
    lib = os.path.join(ROOT, 'ldpgta_b200', 'libldpgta.so')
    if not os.path.exists(lib) and shutil.which('nvcc'):
        import __graft_entry__
        __graft_entry__.build()


def load_golden(name):
    with bz2.open(os.path.join(GOLDEN, name), 'rb') as f:
        return pickle.load(f)


@pytest.fixture(scope='session')
def golden_testmodule():
    return load_golden('testmodule.p.bz2')


@pytest.fixture(scope='session')
def golden_inputs():
    return load_golden('chrom_inputs.p.bz2')


CHROM_CONFIGS = ('homog_m4', 'homog_m8a', 'homog_m3o', 'recent_m4', 'recent_m6', 'distant_m4', 'distant_m6')


def chrom_case(name, inputs):
    """ (bootstrap args without models_dict, golden outputs) for one chromosome config. """
    out = load_golden('chrom_%s.p.bz2' % name)
    cfg = out['cfg']
    hap = {g: inputs['hap_tab'][g] for g in cfg['groups']}
    nhap = {g: inputs['number_of_haplotypes'][g] for g in cfg['groups']}
    return dict(obs_tab=inputs['obs_tabs'][cfg['scenario']], leg_tab=inputs['leg_tab'], hap_tab=hap,
                number_of_haplotypes=nhap, ancestral_makeup=cfg['makeup'], cfg=cfg), out


def rel_err(a, b):
    if a == b:
        return 0.0
    return abs(a - b) / max(abs(a), abs(b))
