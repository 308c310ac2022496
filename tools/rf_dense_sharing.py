import os, sys
sys.path.insert(0, '/root/repo')
import torch
from treecl_b200 import _lib, synth
n, taxa = (int(sys.argv[1]) if len(sys.argv) > 1 else 5000), (int(sys.argv[2]) if len(sys.argv) > 2 else 50)
trees = synth.structured_trees(n, taxa, n_classes=64, class_nni=10, lam=3.0, seed=7)
ts = _lib.TreeSet(trees); ts.upload(0)
pairs = n * (n - 1) // 2
out = torch.empty(pairs, dtype=torch.float64, device='cuda')
res = {}
for k in ('merge', 'auto'):
    if k == 'auto': os.environ.pop('TREEDIST_PAIR_KERNEL', None)
    else: os.environ['TREEDIST_PAIR_KERNEL'] = k
    for _ in range(3): ts.pairwise_device(('rf',), {'rf': out.data_ptr()})
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): ts.pairwise_device(('rf',), {'rf': out.data_ptr()})
    e1.record(); torch.cuda.synchronize()
    res[k] = out.clone()
    print(k, '%.3f ms' % (e0.elapsed_time(e1) / 10), 'dict_shared', ts.info.dict_shared)
print('identical:', torch.equal(res['merge'], res['auto']))
# euc and rf+euc: merge kernel vs the split-extended tile kernel (+ incidence contraction for rf)
outs = {m: torch.empty(pairs, dtype=torch.float64, device='cuda') for m in ('rf', 'wrf', 'euc')}
for metrics in (('euc',), ('rf', 'euc'), ('wrf',), ('rf', 'wrf', 'euc')):
    keep = {}
    for k in ('merge', 'auto'):
        if k == 'auto': os.environ.pop('TREEDIST_PAIR_KERNEL', None)
        else: os.environ['TREEDIST_PAIR_KERNEL'] = k
        ptrs = {m: outs[m].data_ptr() for m in metrics}
        for _ in range(3): ts.pairwise_device(metrics, ptrs)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): ts.pairwise_device(metrics, ptrs)
        e1.record(); torch.cuda.synchronize()
        keep[k] = {m: outs[m].clone() for m in metrics}
        print('+'.join(metrics), k, '%.3f ms' % (e0.elapsed_time(e1) / 10))
    for m in metrics:
        if m == 'rf':
            print('   rf identical: %s' % torch.equal(keep['merge'][m], keep['auto'][m]))
        else:
            d = (keep['merge'][m] - keep['auto'][m]).abs() / keep['merge'][m].clamp_min(1e-300)
            print('   %s max rel diff %.2e, exact zeros agree: %s' % (m, float(d[keep['merge'][m] > 0].max()), bool(((keep['merge'][m] == 0) == (keep['auto'][m] == 0)).all())))
