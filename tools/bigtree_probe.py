"""Big-tree kernel (more than 128 internal edges per tree): geodesic and rf+euc throughput on uniform random trees."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from treecl_b200 import _lib, synth
n, taxa = int(sys.argv[1]), int(sys.argv[2])
kind = sys.argv[3] if len(sys.argv) > 3 else 'uniform'
trees = synth.uniform_trees(n, taxa, seed=5) if kind == 'uniform' else synth.structured_trees(n, taxa, n_classes=8, class_nni=30, lam=8.0, seed=5)
ts = _lib.TreeSet(trees); st = torch.cuda.current_stream(); ts.upload(0, st.cuda_stream)
pairs = n * (n - 1) // 2
out = torch.zeros(pairs, dtype=torch.float64, device='cuda:0')
f = lambda: ts.pairwise_device(('geo',), {'geo': out.data_ptr()}, stream=st.cuda_stream)
f(); torch.cuda.synchronize(); ts.geo_stats()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(st); f(); b.record(st); torch.cuda.synchronize()
ms = a.elapsed_time(b); s = ts.geo_stats()
print(json.dumps({'n': n, 'taxa': taxa, 'kind': kind, 'ms': round(ms, 2), 'pairs_per_s': round(pairs / ms * 1e3), 'maxflows_per_pair': round(s['maxflow_solves'] / pairs, 2),
                  'augmentations_per_maxflow': round(s['augmentations'] / max(s['maxflow_solves'], 1), 2), 'checksum': float(out.sum())}))
