"""Experiment: RF through the int8 tcgen05 incidence contraction vs the merge / join kernels on a structured (clustered) set.

usage: python tools/bench_incidence.py N_TREES N_TAXA [N_CLASSES] [LAMBDA]
Prints one JSON line with per-kernel times (CUDA events on torch's current stream), dictionary size and tensor throughput.
"""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from treecl_b200 import _lib, synth

n, taxa = int(sys.argv[1]), int(sys.argv[2])
n_classes = int(sys.argv[3]) if len(sys.argv) > 3 else 32
lam = float(sys.argv[4]) if len(sys.argv) > 4 else 4.0
skip_slow = len(sys.argv) > 5 and sys.argv[5] == 'only'
t0 = time.time()
# BASELINE config 4's generator (synth.BASELINE_CONFIGS[3]) at the requested size
g = synth.NativeTrees(n, taxa, 'structured', seed=0xCA55E77E + 3, n_classes=n_classes, class_nni=taxa // 5, lam=lam)
t_gen = time.time() - t0
ts = g.treeset()
g.close()
info = ts.info
st = torch.cuda.current_stream()
ts.upload(0, st.cuda_stream)
pairs = n * (n - 1) // 2
out16 = torch.zeros(pairs, dtype=torch.int16, device='cuda:0')
out64 = torch.zeros(pairs if pairs * 8 < 60e9 else 1, dtype=torch.float64, device='cuda:0')


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in ev:
        a.record(st); fn(); b.record(st)
    torch.cuda.synchronize()
    return float(np.median([a.elapsed_time(b) for a, b in ev]))


res = {'n_trees': n, 'n_taxa': taxa, 'pairs': pairs, 'dict_size': int(info.dict_size), 'dict_shared': int(info.dict_shared),
       'max_shared': int(info.max_shared), 'gen_s': round(t_gen, 1)}
res['incidence_u16_ms'] = timed(lambda: ts.rf_incidence_device(out16.data_ptr(), out_is_u16=True, stream=st.cuda_stream))
if out64.numel() > 1:
    res['incidence_f64_ms'] = timed(lambda: ts.rf_incidence_device(out64.data_ptr(), stream=st.cuda_stream))
ref = out64.clone() if not skip_slow else None
for k in (() if skip_slow or out64.numel() == 1 else ('merge', 'join')):
    os.environ['TREEDIST_PAIR_KERNEL'] = k
    try:
        res[k + '_f64_ms'] = timed(lambda: ts.pairwise_device(('rf',), {'rf': out64.data_ptr()}, stream=st.cuda_stream), reps=3)
        res[k + '_equal'] = bool(torch.equal(out64, ref))
    except Exception as e:          # noqa
        res[k + '_f64_ms'] = None; res[k + '_error'] = str(e)[:80]
d_pad = (info.dict_shared + 127) // 128 * 128
tiles = (n + 127) // 128
ops = 2.0 * 128 * 128 * d_pad * (tiles * (tiles + 1) // 2)
res['d_pad'] = d_pad
res['incidence_tops_u16'] = ops / (res['incidence_u16_ms'] * 1e-3) / 1e12
res['incidence_pairs_per_s'] = pairs / (res['incidence_u16_ms'] * 1e-3)
print(json.dumps(res))
