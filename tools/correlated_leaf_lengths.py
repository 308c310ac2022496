import os, re, sys
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from treecl_b200 import _lib, synth
n, taxa = 5000, 50
trees = synth.uniform_trees(n, taxa, seed=99)
rng = np.random.default_rng(5)
labels = sorted(set(re.findall(r'([A-Za-z_]\w*):', trees[0])))
base = {l: rng.exponential(1.0) for l in labels}
def corr(t, noise):
    return re.sub(r'([A-Za-z_]\w*):([0-9.eE+-]+)', lambda m: '%s:%.17g' % (m.group(1), base[m.group(1)] * (1 + noise * rng.standard_normal())), t)
pairs = n * (n - 1) // 2
outs = {m: torch.empty(pairs, dtype=torch.float64, device='cuda') for m in ('rf', 'euc')}
ptrs = {m: outs[m].data_ptr() for m in outs}
for name, ts_trees in (('independent', trees), ('leaf lengths correlated, 10% noise', [corr(t, 0.1) for t in trees]), ('leaf lengths correlated, 1% noise', [corr(t, 0.01) for t in trees])):
    ts = _lib.TreeSet(ts_trees); ts.upload(0)
    for k in ('split', 'merge'):
        os.environ['TREEDIST_PAIR_KERNEL'] = k
        for _ in range(3): ts.pairwise_device(('rf', 'euc'), ptrs)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): ts.pairwise_device(('rf', 'euc'), ptrs)
        e1.record(); torch.cuda.synchronize()
        if k == 'split': a = outs['euc'].clone()
        print('%-40s %-6s %.3f ms' % (name, k, e0.elapsed_time(e1) / 5))
    print('   max rel diff split vs merge: %.2e' % float(((a - outs['euc']).abs() / outs['euc'].clamp_min(1e-300)).max()))
    ts.close()
