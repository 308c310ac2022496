"""Debug helper: compare the rf/euc output of two kernel variants at bench size and describe the mismatches."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from treecl_b200 import _lib, synth

n, taxa = int(sys.argv[1]) if len(sys.argv) > 1 else 5000, 50
trees = synth.uniform_trees(n, taxa, seed=0xCA55E77E + 1)
ts = _lib.TreeSet(trees)
res = {}
for k in ('split', 'join_tile', 'join', 'merge'):
    os.environ['TREEDIST_PAIR_KERNEL'] = k
    res[k] = ts.pairwise(('rf', 'euc'))
    res[k + '2'] = ts.pairwise(('rf', 'euc'))
ref = res['merge']
for k in ('split', 'split2', 'join_tile', 'join', 'join2'):
    for m in ('rf', 'euc'):
        bad = np.nonzero(res[k][m] != ref[m])[0] if m == 'rf' else np.nonzero(np.abs(res[k][m] - ref[m]) > 1e-9 * ref[m])[0]
        print(k, m, 'mismatches', bad.size)
        if bad.size:
            # decode (i, j)
            i = (n - 2 - np.floor(np.sqrt(-8.0 * bad + 4.0 * n * (n - 1) - 7) / 2.0 - 0.5)).astype(np.int64)
            j = bad + i + 1 - n * (n - 1) // 2 + (n - i) * ((n - i) - 1) // 2
            print('  rows', np.unique(i // 64)[:20], 'coltiles', np.unique(j // 32)[:20])
            print('  sample', list(zip(i[:8], j[:8], res[k][m][bad[:8]], ref[m][bad[:8]])))
print('split vs split2 identical:', all(np.array_equal(res['split'][m], res['split2'][m]) for m in ('rf', 'euc')))
print('join vs join2 identical:', all(np.array_equal(res['join'][m], res['join2'][m]) for m in ('rf', 'euc')))
