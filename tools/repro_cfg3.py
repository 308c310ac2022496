"""Repro target: BASELINE config 3 set (or its first N trees), row-sharded launches.
usage: python tools/repro_cfg3.py [n_trees] [metrics] [parts] [py|native]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from treecl_b200 import _lib, engine, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
metrics = tuple(sys.argv[2].split(',')) if len(sys.argv) > 2 else ('wrf',)
parts = int(sys.argv[3]) if len(sys.argv) > 3 else 8
taxa = int(sys.argv[5]) if len(sys.argv) > 5 else 100
if taxa != 100:
    g = synth.NativeTrees(n, taxa, 'uniform', seed=0xCA55E77E + 2)
    ts = g.treeset()
elif len(sys.argv) > 4 and sys.argv[4] == 'py':
    ts = _lib.TreeSet(synth.uniform_trees(n, 100, seed=0xCA55E77E + 2))
else:
    g = synth.baseline_config(2, n_trees=n)
    ts = g.treeset()
i = ts.info
print('max_shared', i.max_shared, 'dict_shared', i.dict_shared, 'heavy_rows', i.heavy_rows, 'mean_common', i.mean_common, flush=True)
for a, b in engine.balanced_row_ranges(n, parts):
    out = ts.pairwise(metrics, row_begin=a, row_end=b)
    print('rows', a, b, {m: float(out[m].sum()) for m in metrics}, flush=True)
print('ok')
