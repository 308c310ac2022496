"""Host encoder: time against thread count, with the phase laps (TREEDIST_TIMING) of the last run."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from treecl_b200 import _lib, synth
trees = synth.uniform_trees(5000, 50, seed=0xCA55E77E + 1)
print('cores', os.cpu_count())
for nt in (1, 2, 4, 8, 16, 32, 0):
    _lib.TreeSet(trees, n_threads=nt).close()
    t0 = time.perf_counter()
    for _ in range(5): _lib.TreeSet(trees, n_threads=nt).close()
    print('threads %2d: %.2f ms' % (nt, (time.perf_counter() - t0) / 5 * 1e3), flush=True)
os.environ['TREEDIST_TIMING'] = '1'
_lib.TreeSet(trees).close()
