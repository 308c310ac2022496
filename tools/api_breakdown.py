"""Where the time of the drop-in call goes: Collection.get_inter_tree_distances(metric) on the bench set, phase by phase."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import treecl_b200
from treecl_b200 import _lib, synth

trees = synth.uniform_trees(5000, 50, seed=0xCA55E77E + 1)
coll = treecl_b200.Collection(records=[treecl_b200.Record('g%05d' % k, ml_tree=t) for k, t in enumerate(trees)])


def lap(label, fn, reps=3):
    fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        r = fn()
    print('{:46s} {:8.2f} ms'.format(label, (time.perf_counter() - t0) / reps * 1e3), flush=True)
    return r


for m in ('rf', 'euc', 'wrf', 'geo'):
    lap('get_inter_tree_distances(%s)' % m, lambda: coll.get_inter_tree_distances(m, show_progress=False), reps=2 if m == 'geo' else 3)
ts = lap('TreeSet(trees) [parse + encode]', lambda: _lib.TreeSet(trees))
lap('upload', lambda: _lib.TreeSet(trees).upload(0))
ts.upload(0)
for m in ('rf', 'euc'):
    os.environ['TREEDIST_DEBUG'] = '1'
    ts.pairwise_square(m)
    os.environ.pop('TREEDIST_DEBUG')
    sq = lap('pairwise_square(%s) [kernels + D2H N x N]' % m, lambda: ts.pairwise_square(m))
    lap('pairwise(%s) condensed' % m, lambda: ts.pairwise((m,)))
    lap('DistanceMatrix.from_array', lambda: treecl_b200.DistanceMatrix.from_array(sq, coll.names))
