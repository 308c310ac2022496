"""Time the rf+euc matrix with the postings pipeline ('split') and the merge kernel on sets of increasing split sharing."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from treecl_b200 import _lib, synth

n, taxa = 5000, 50
pairs = n * (n - 1) // 2
outs = {m: torch.empty(pairs, dtype=torch.float64, device='cuda') for m in ('rf', 'euc')}
ptrs = {m: outs[m].data_ptr() for m in outs}
sets = [('uniform', synth.uniform_trees(n, taxa, seed=7))]
for classes, nni, lam in ((256, 60, 30.0), (64, 60, 30.0), (64, 40, 20.0), (64, 30, 10.0), (64, 20, 6.0), (64, 10, 3.0), (8, 10, 1.0)):
    sets.append(('structured c%d nni%d lam%g' % (classes, nni, lam), synth.structured_trees(n, taxa, n_classes=classes, class_nni=nni, lam=lam, seed=7)))
for name, trees in sets:
    ts = _lib.TreeSet(trees); ts.upload(0)
    res = {}
    for k in ('split', 'merge', 'extended', 'auto'):
        os.environ['TREEDIST_PAIR_KERNEL'] = k
        try:
            for _ in range(3):
                ts.pairwise_device(('rf', 'euc'), ptrs)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                ts.pairwise_device(('rf', 'euc'), ptrs)
            e1.record(); torch.cuda.synchronize()
            res[k] = e0.elapsed_time(e1) / 5
        except Exception as ex:
            res[k] = float('nan')
    rf = outs['rf']
    mean_common = float((47 * 2 - rf.double().mean()) / 2) if True else 0
    print('%-32s mean common ~%6.2f dict_shared %6d | split %.3f ms | merge %.3f ms | extended %.3f ms | auto %.3f ms' % (name, mean_common, ts.info.dict_shared, res['split'], res['merge'], res['extended'], res['auto']))
    ts.close()
