"""Does the library's kernel choice (cost model in api.cu, calibrated at 5,000 x 50 and 12,000 x 80) still pick a good kernel at other
sizes?  Times every forced kernel against 'auto' on uniform and clustered sets of 100 and 200 taxa.  Prints one line per (set, metrics)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from treecl_b200 import _lib, synth

st = torch.cuda.current_stream()
cases = [('uniform 20000 x 100', synth.NativeTrees(20000, 100, 'uniform', seed=11)),
         ('clustered 20000 x 100 (32 classes, lam 4)', synth.NativeTrees(20000, 100, 'structured', seed=11, n_classes=32, class_nni=20, lam=4.0)),
         ('clustered 20000 x 100 (256 classes, lam 20)', synth.NativeTrees(20000, 100, 'structured', seed=12, n_classes=256, class_nni=40, lam=20.0)),
         ('uniform 8000 x 200', synth.NativeTrees(8000, 200, 'uniform', seed=13)),
         ('clustered 8000 x 200 (16 classes, lam 6)', synth.NativeTrees(8000, 200, 'structured', seed=13, n_classes=16, class_nni=40, lam=6.0))]
worst = 1.0
for name, g in cases:
    ts = g.treeset(); g.close()
    n = ts.n_trees
    ts.upload(0, st.cuda_stream)
    pairs = n * (n - 1) // 2
    info = ts.info
    for metrics in (('rf',), ('euc',), ('rf', 'euc'), ('wrf', 'euc')):
        outs = {m: torch.empty(pairs, dtype=torch.float64, device='cuda') for m in metrics}
        ptrs = {m: outs[m].data_ptr() for m in metrics}
        res = {}
        for k in ('auto', 'split', 'merge', 'extended', 'incidence'):
            os.environ['TREEDIST_PAIR_KERNEL'] = k
            try:
                for _ in range(2):
                    ts.pairwise_device(metrics, ptrs, stream=st.cuda_stream)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(st)
                for _ in range(3):
                    ts.pairwise_device(metrics, ptrs, stream=st.cuda_stream)
                e1.record(st); torch.cuda.synchronize()
                res[k] = e0.elapsed_time(e1) / 3
            except Exception as ex:          # noqa
                res[k] = float('nan')
        os.environ.pop('TREEDIST_PAIR_KERNEL')
        best = np.nanmin([v for k, v in res.items() if k != 'auto'])
        worst = max(worst, res['auto'] / best)
        print('%-44s %-8s mean common %6.2f | auto %8.3f ms | ' % (name, '+'.join(metrics), info.mean_common, res['auto'])
              + ' '.join('%s %.3f' % (k, res[k]) for k in ('split', 'merge', 'extended', 'incidence')) + ' | auto / best %.2f' % (res['auto'] / best), flush=True)
        del outs
    ts.close(); _lib.trim_cache(0); torch.cuda.empty_cache()
print('worst auto / best forced: %.2f' % worst)
