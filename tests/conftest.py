import json
import os
import re
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')

SORT_KEY = lambda item: tuple((int(num) if num else alpha) for (num, alpha) in
                              re.findall(r'(\d+)|(\D+)', item))


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


@pytest.fixture(scope='session')
def golden_scalars():
    with open(os.path.join(GOLDEN, 'scalars.json')) as fh:
        return json.load(fh)


@pytest.fixture(scope='session')
def fixture_trees():
    """name -> Newick of the 15 ML trees under tests/golden/trees (reference tests/data/trees)."""
    out = {}
    for f in os.listdir(os.path.join(GOLDEN, 'trees')):
        if f.endswith('.nwk'):
            with open(os.path.join(GOLDEN, 'trees', f)) as fh:
                out[f[:-4]] = fh.read().strip()
    return out


@pytest.fixture(scope='session')
def cache_trees():
    """(names in SORT_KEY order, ml_tree strings) from the reference's tests/data/cache/*.json.gz."""
    with open(os.path.join(GOLDEN, 'cache_trees.json')) as fh:
        d = json.load(fh)
    names = sorted(d, key=SORT_KEY)
    return names, [d[n]['ml_tree'] for n in names]


@pytest.fixture(scope='session')
def oracle_mod():
    import oracle
    oracle.build()
    return oracle


@pytest.fixture(scope='session')
def lib_built():
    """Make sure libtreedist_cuda.so exists (cross-compiles without a GPU)."""
    import subprocess
    so = os.path.join(ROOT, 'treecl_b200', 'libtreedist_cuda.so')
    if not os.path.exists(so):
        subprocess.check_call(['make', '-s', '-C', os.path.join(ROOT, 'treecl_b200', 'csrc')])
    return so
