"""Geodesic kernel: throughput and GTP statistics on a uniform set (config 5 = 2000 x 30)."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from treecl_b200 import _lib, synth
n, taxa = int(sys.argv[1]), int(sys.argv[2])
trees = synth.uniform_trees(n, taxa, seed=0xCA55E77E + 4)
ts = _lib.TreeSet(trees); st = torch.cuda.current_stream(); ts.upload(0, st.cuda_stream)
pairs = n * (n - 1) // 2
out = torch.zeros(pairs, dtype=torch.float64, device='cuda:0')
ts.pairwise_device(('geo',), {'geo': out.data_ptr()}, stream=st.cuda_stream); torch.cuda.synchronize(); ts.geo_stats()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(st); ts.pairwise_device(('geo',), {'geo': out.data_ptr()}, stream=st.cuda_stream); b.record(st); torch.cuda.synchronize()
ms = a.elapsed_time(b); s = ts.geo_stats()
print(json.dumps({'n': n, 'taxa': taxa, 'ms': ms, 'pairs_per_s': pairs / ms * 1e3, 'maxflows_per_pair': s['maxflow_solves'] / pairs,
                  'ratios_per_pair': s['ratios'] / pairs, 'augmentations_per_maxflow': s['augmentations'] / max(s['maxflow_solves'], 1)}))
