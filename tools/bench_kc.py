"""Kendall-Colijn matrix on one B200: kernels only (CUDA events) and end to end, beside the oracle on the host cores."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from treecl_b200 import _lib, synth

n, taxa = (int(sys.argv[1]) if len(sys.argv) > 1 else 5000), (int(sys.argv[2]) if len(sys.argv) > 2 else 50)
trees = synth.uniform_trees(n, taxa, seed=0xCA55E77E + 7)
pairs = n * (n - 1) // 2
t0 = time.perf_counter(); ts = _lib.TreeSet.kendall_colijn(trees, 0.5); t1 = time.perf_counter()
ts.upload(0); t2 = time.perf_counter()
out = torch.empty(pairs, dtype=torch.float64, device='cuda')
for _ in range(3):
    ts.pairwise_device(('euc',), {'euc': out.data_ptr()})
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    ts.pairwise_device(('euc',), {'euc': out.data_ptr()})
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
D = ts.info.n_taxa
print('%d trees x %d taxa: vectors of %d entries; encode %.1f ms, upload %.1f ms, kernels %.3f ms = %.1f G pairs/s, %.1f TFLOP/s fp64 (2 D flops per pair)'
      % (n, taxa, D, (t1 - t0) * 1e3, (t2 - t1) * 1e3, ms, pairs / ms / 1e6, 2.0 * D * pairs / ms / 1e9))
host = torch.empty(pairs, dtype=torch.float64).pin_memory().numpy()
t0 = time.perf_counter()
t = _lib.TreeSet.kendall_colijn(trees, 0.5); t.pairwise(('euc',), device=0, out={'euc': host}); t.close()
t1 = time.perf_counter()
print('end to end (Newick -> matrix in pinned host memory): %.1f ms = %.1f M pairs/s' % ((t1 - t0) * 1e3, pairs / (t1 - t0) / 1e6))
# parity on a sample + CPU oracle timing on a bounded sample
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle
m = 160
t0 = time.perf_counter(); want = oracle.kc_matrix(trees[:m], 0.5); t1 = time.perf_counter()
idx = [_lib.condensed_offset(n, i) + (j - i - 1) for i in range(m) for j in range(i + 1, m)]
got = host[idx]
print('parity on the first %d trees: max rel err %.2e; oracle (1 core, trees parsed once): %.0f pairs/s' % (m, np.max(np.abs(got - want) / np.maximum(want, 1e-300)), len(idx) / (t1 - t0)))
