"""A non-uniform set for profiling general_kernel: 600 x 40 taxa, every tree pruned to its own random leaf subset (no groups)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, json
import treecl_b200 as td
from treecl_b200 import _lib, synth
n, taxa = 600, 40
rng = np.random.default_rng(3)
labels = synth.taxon_labels(taxa)
trees = [td.Tree(t).prune_to_subset({labels[i] for i in rng.choice(taxa, size=int(rng.integers(28, taxa)), replace=False)}).newick
         for t in synth.uniform_trees(n, taxa, seed=5)]
ts = _lib.TreeSet(trees); st = torch.cuda.current_stream(); ts.upload(0, st.cuda_stream)
pairs = n * (n - 1) // 2
outs = {m: torch.empty(pairs, dtype=torch.float64, device='cuda') for m in ('rf', 'wrf', 'euc', 'geo')}
f = lambda: ts.pairwise_device(tuple(outs), {m: outs[m].data_ptr() for m in outs}, stream=st.cuda_stream)
f(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(st); f(); e1.record(st); torch.cuda.synchronize()
print(json.dumps({'n': n, 'taxa': taxa, 'ms': e0.elapsed_time(e1), 'pairs_per_s': pairs / e0.elapsed_time(e1) * 1e3, 'groups': int(ts.info.n_groups)}))
