"""a 5,000-tree set in which 1 % of the trees miss taxa - uniform set vs grouped vs ungrouped (TREEDIST_NO_GROUPS=1)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from treecl_b200 import _lib, synth
import treecl_b200 as td

n, taxa = 5000, 50
full = synth.uniform_trees(n, taxa, seed=0xCA55E77E + 1)
labels = synth.taxon_labels(taxa)
rng = np.random.default_rng(1)
trees = list(full)
for k in rng.choice(n, size=n // 100, replace=False):
    keep = rng.choice(taxa, size=int(rng.integers(taxa - 6, taxa)), replace=False)
    trees[k] = td.Tree(full[k]).prune_to_subset({labels[i] for i in keep}).newick
st = torch.cuda.current_stream()
pairs = n * (n - 1) // 2
out = {m: torch.empty(pairs, dtype=torch.float64, device='cuda') for m in ('rf', 'euc')}


def timed(ts):
    ts.upload(0, st.cuda_stream)
    f = lambda: ts.pairwise_device(('rf', 'euc'), {m: out[m].data_ptr() for m in out}, stream=st.cuda_stream)
    f(); torch.cuda.synchronize()
    ms = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st); f(); e1.record(st); torch.cuda.synchronize(); ms.append(e0.elapsed_time(e1))
    return float(np.median(ms))


res = {'uniform_ms': timed(_lib.TreeSet(full))}
g = _lib.TreeSet(trees); res['grouped_ms'] = timed(g); res['n_groups'] = int(g.info.n_groups)
grouped = {m: out[m].clone() for m in out}
os.environ['TREEDIST_NO_CROSS_LISTS'] = '1'
res['grouped_chunk_skipping_ms'] = timed(_lib.TreeSet(trees))
os.environ.pop('TREEDIST_NO_CROSS_LISTS')
os.environ['TREEDIST_NO_GROUPS'] = '1'
res['ungrouped_ms'] = timed(_lib.TreeSet(trees))
res['same_rf'] = bool(torch.equal(grouped['rf'], out['rf'])); res['max_rel_euc'] = float(((grouped['euc'] - out['euc']).abs() / out['euc'].abs().clamp_min(1e-300)).max())
print(res)
