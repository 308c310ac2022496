#!/usr/bin/env python
"""Extract the reference's own golden vectors for the tree-distance path into tests/golden/.

Run ONCE in the build container (where /root/reference is mounted); the output is committed so that
the GPU box (which has no /root/reference) can run the parity tests.

Sources (all DATA fixtures and hard-coded expected values, no reference source code):
  * /root/reference/tests/data/trees/*.nwk          15 ML trees, 10 taxa (tests/test.py:235-241)
  * /root/reference/tests/data/cache/*.json.gz      'ml_tree' strings used by tests/test.py:320-326,437-467
  * /root/reference/tests/data/cache/geo_dm.csv     expected 15x15 geodesic matrix (tests/test.py:298-300)
  * the scalar known answers typed into tests/test.py:243-294,320-326
"""
import glob
import gzip
import json
import os
import re
import shutil

REF = '/root/reference/tests/data'
HERE = os.path.dirname(os.path.abspath(__file__))

SORT_KEY = lambda item: tuple((int(num) if num else alpha) for (num, alpha) in
                              re.findall(r'(\d+)|(\D+)', item))


def main():
    os.makedirs(os.path.join(HERE, 'trees'), exist_ok=True)
    for f in sorted(glob.glob(os.path.join(REF, 'trees', '*.nwk'))):
        shutil.copy(f, os.path.join(HERE, 'trees', os.path.basename(f)))
    shutil.copy(os.path.join(REF, 'cache', 'geo_dm.csv'), os.path.join(HERE, 'geo_dm.csv'))

    cache = {}
    for f in sorted(glob.glob(os.path.join(REF, 'cache', '*.json.gz')), key=SORT_KEY):
        name = os.path.basename(f)[:-len('.json.gz')]
        d = json.loads(gzip.open(f).read().decode('utf-8'))
        cache[name] = {'ml_tree': d.get('ml_tree'), 'nj_tree': d.get('nj_tree')}
    with open(os.path.join(HERE, 'cache_trees.json'), 'w') as out:
        json.dump(cache, out, indent=1, sort_keys=True)

    scalars = {
        '_source': 'reference tests/test.py, hard-coded expected values; assertAlmostEqual = 7 d.p.',
        'pair': ['class1_1', 'class1_2'],
        'geo': 0.7051762453884542,            # tests/test.py:245
        'geo_norm': 0.11511918913951155,      # :249
        'euc': 0.6463384798834909,            # :253
        'euc_norm': 0.10551399341715575,      # :257
        'wrf': 2.1666480728499717,            # :261
        'wrf_norm': 0.1063259354618597,       # :265
        'rf': 2.0,                            # :269
        'rf_norm': 0.14285714285714285,       # :273
        'partial_overlap': {                  # :280-294
            'keep_0': ['Sp1', 'Sp2', 'Sp3', 'Sp4', 'Sp5', 'Sp6', 'Sp7'],
            'keep_1': ['Sp4', 'Sp5', 'Sp6', 'Sp7', 'Sp8', 'Sp9', 'Sp10'],
            'geo': 0.35314280333144532,
            'min_overlap_5_fail_value': -1,
        },
        'geo_matrix_checksum': 412.70677069540181,   # :300,326,446-467
    }
    with open(os.path.join(HERE, 'scalars.json'), 'w') as out:
        json.dump(scalars, out, indent=1)
    print('wrote', len(cache), 'cache trees,', len(glob.glob(os.path.join(HERE, 'trees', '*.nwk'))), 'nwk files')


if __name__ == '__main__':
    main()
