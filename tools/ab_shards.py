"""Per-rank cost of a strong-scaled matrix on ONE GPU: times the row blocks rank r of an N-way fold would run (BASELINE config 3 by
default), one rank after the other - the max over ranks is what an N-GPU step would take.  usage: ab_shards.py [parts] [config]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from treecl_b200 import _lib, engine, synth

parts = int(sys.argv[1]) if len(sys.argv) > 1 else 8
index = int(sys.argv[2]) if len(sys.argv) > 2 else 2
cfg = synth.BASELINE_CONFIGS[index]
g = synth.baseline_config(index) if index in (2, 3) else None
ts = g.treeset() if g else _lib.TreeSet(synth.uniform_trees(cfg['n_trees'], cfg['n_taxa'], seed=0xCA55E77E + index))
n, metrics = ts.n_trees, cfg['metrics']
st = torch.cuda.current_stream()
ts.upload(0, st.cuda_stream)
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
times = []
for r, blocks in enumerate(engine.folded_row_ranges(n, parts)):
    outs = [(a, b, {m: torch.empty(_lib.condensed_offset(n, b) - _lib.condensed_offset(n, a), dtype=torch.float64, device='cuda') for m in metrics}) for a, b in blocks]
    blocks_dev = [(a, b, {m: o[m].data_ptr() for m in metrics}) for a, b, o in outs]
    def step():
        if os.environ.get('AB_SEPARATE_CALLS'):
            for a, b, o in outs:
                ts.pairwise_device(metrics, {m: o[m].data_ptr() for m in metrics}, row_begin=a, row_end=b, stream=st.cuda_stream)
        else:
            ts.pairwise_device_blocks(metrics, blocks_dev, stream=st.cuda_stream)
    for _ in range(3): step()
    ms = []
    for _ in range(8):
        flush.zero_(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st); step(); e1.record(st); torch.cuda.synchronize(); ms.append(e0.elapsed_time(e1))
    times.append(float(np.median(ms)))
    del outs
full = {m: torch.empty(n * (n - 1) // 2, dtype=torch.float64, device='cuda') for m in metrics}
ms = []
for _ in range(6):
    flush.zero_(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st); ts.pairwise_device(metrics, {m: full[m].data_ptr() for m in metrics}, stream=st.cuda_stream); e1.record(st); torch.cuda.synchronize(); ms.append(e0.elapsed_time(e1))
one = float(np.median(ms[1:]))
print('parts %d: per-rank ms %s  max %.4f  whole matrix on one GPU %.4f  => strong-scaling efficiency %.3f (EVENT_PARTS=%s, NO_PDL=%s)' % (
    parts, ' '.join('%.4f' % t for t in times), max(times), one, one / (parts * max(times)), os.environ.get('TREEDIST_EVENT_PARTS', 'auto'), os.environ.get('TREEDIST_NO_PDL', '0')))
