"""BASELINE config 3 (20,000 x 100 uniform random trees), wrf+euc on one GPU, device-resident outputs: one timed call (ncu target)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from treecl_b200 import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
metrics = tuple(sys.argv[2].split(',')) if len(sys.argv) > 2 else ('wrf', 'euc')
g = synth.baseline_config(2, n_trees=n); ts = g.treeset(); g.close()
st = torch.cuda.current_stream(); ts.upload(0, st.cuda_stream)
pairs = n * (n - 1) // 2
outs = {m: torch.empty(pairs, dtype=torch.float64, device='cuda') for m in metrics}
f = lambda: ts.pairwise_device(metrics, {m: outs[m].data_ptr() for m in metrics}, stream=st.cuda_stream)
f(); torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(st); f(); b.record(st); torch.cuda.synchronize()
print(json.dumps({'n': n, 'metrics': metrics, 'ms': a.elapsed_time(b), 'pairs_per_s': pairs / a.elapsed_time(b) * 1e3}))
