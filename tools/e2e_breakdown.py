"""Where the end-to-end time of one matrix goes (Newick strings -> condensed matrices in pinned host memory)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from treecl_b200 import _lib, synth
n, taxa = 5000, 50
trees = synth.uniform_trees(n, taxa, seed=0xCA55E77E + 1)
pairs = n * (n - 1) // 2
host = {m: torch.empty(pairs, dtype=torch.float64).pin_memory().numpy() for m in ('rf', 'euc')}
for it in range(3):
    t0 = time.perf_counter(); ts = _lib.TreeSet(trees); t1 = time.perf_counter()
    ts.upload(0); t2 = time.perf_counter()
    ts.pairwise(('rf', 'euc'), device=0, out=host); t3 = time.perf_counter()
    ts.close(); t4 = time.perf_counter()
    print('encode %.1f ms | upload %.1f ms | kernels + D2H %.1f ms | free %.1f ms | total %.1f ms' % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, (t4 - t3) * 1e3, (t4 - t0) * 1e3))
os.environ['TREEDIST_TIMING'] = '1'
ts = _lib.TreeSet(trees)
