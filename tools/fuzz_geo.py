"""Randomised cross-check of the warp-per-pair kernels against the oracle on one GPU: geodesic_kernel (all register forms, flow matrix in
shared or global memory), general_kernel (with and without the geodesic, leaf-set groups, pairs across groups from lists or by skipping),
bigtree_kernel - random sizes, taxa, tree shapes, pruning patterns, metric subsets, normalisation, min_overlap and row ranges.
usage: python tools/fuzz_geo.py [seconds] [seed]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import treecl_b200 as td
from treecl_b200 import _lib, synth
from oracle import oracle

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 2024
rng = np.random.default_rng(seed0)
ALL = ('rf', 'wrf', 'euc', 'geo')
t_end, cases, worst = time.time() + budget, 0, 0.0
while time.time() < t_end:
    taxa = int(rng.choice([5, 8, 13, 20, 33, 35, 40, 50, 66, 67, 90, 131, 140, 210]))
    n = int(rng.choice([2, 3, 9, 40, 70, 130])) if taxa <= 67 else int(rng.choice([2, 5, 12, 24]))
    kind = rng.choice(['uniform', 'structured'])
    seed = int(rng.integers(1 << 30))
    trees = synth.uniform_trees(n, taxa, seed=seed) if kind == 'uniform' else \
        synth.structured_trees(n, taxa, n_classes=int(rng.integers(1, 5)), class_nni=int(rng.integers(1, 12)), lam=float(rng.uniform(0.5, 6)), seed=seed)
    labels = synth.taxon_labels(taxa)
    mode = rng.choice(['same', 'odd_few', 'odd_all', 'two_groups'])
    trees = list(trees)
    if mode != 'same' and taxa >= 8:
        lo = max(4, taxa // 2)
        if mode == 'two_groups':
            miss = {labels[i] for i in rng.choice(taxa, size=taxa - int(rng.integers(1, 4)), replace=False)}
            for k in rng.choice(n, size=n // 2, replace=False):
                trees[k] = td.Tree(trees[k]).prune_to_subset(miss).newick
        for k in range(n):
            if mode == 'odd_all' or (mode == 'odd_few' and rng.random() < 0.1) or (mode == 'two_groups' and rng.random() < 0.03):
                keep = rng.choice(taxa, size=int(rng.integers(lo, taxa + 1)), replace=False)
                t = td.Tree(trees[k])
                sub = {labels[i] for i in keep} & set(t.labels)          # (prune_to_subset refuses a set that is not within the tree's leaves)
                if len(sub) >= 4:
                    trees[k] = t.prune_to_subset(sub).newick
    metrics = tuple(m for m in ALL if rng.random() < 0.6) or ('geo',)
    norm = bool(rng.random() < 0.4)
    mo = int(rng.choice([4, max(4, taxa // 2), max(4, (3 * taxa) // 4)]))
    env = {}
    if rng.random() < 0.3: env['TREEDIST_GEO_FLOW'] = str(rng.choice(['smem', 'global']))
    if rng.random() < 0.2: env['TREEDIST_NO_CROSS_LISTS'] = '1'
    for k in ('TREEDIST_GEO_FLOW', 'TREEDIST_NO_CROSS_LISTS'):
        os.environ.pop(k, None)
    os.environ.update(env)
    ts = _lib.TreeSet(trees)
    r0 = int(rng.integers(0, max(n - 1, 1))) if rng.random() < 0.3 else 0
    r1 = int(rng.integers(r0 + 1, n + 1)) if r0 or rng.random() < 0.3 else n
    got = ts.pairwise(metrics, normalise=norm, min_overlap=mo, overlap_fail_value=-3.25, row_begin=r0, row_end=r1)
    p0, p1 = _lib.condensed_offset(n, r0), _lib.condensed_offset(n, min(r1, n))
    for m in metrics:
        want = oracle.matrix(trees, m, normalise=norm, min_overlap=mo, overlap_fail_value=-3.25, nthreads=8)[p0:p1]
        ok = np.array_equal(got[m], want) if m == 'rf' else np.allclose(got[m], want, rtol=1e-9, atol=1e-12)
        if not ok:
            bad = np.flatnonzero(~np.isclose(got[m], want, rtol=1e-9, atol=1e-12))
            print('MISMATCH', dict(n=n, taxa=taxa, kind=str(kind), seed=seed, mode=str(mode), metrics=metrics, norm=norm, mo=mo, rows=(r0, r1), env=env, metric=m,
                                   n_bad=int(bad.size), first=(int(bad[0]), float(got[m][bad[0]]), float(want[bad[0]])) if bad.size else None), flush=True)
            sys.exit(1)
        if m != 'rf' and want.size:
            worst = max(worst, float(np.max(np.abs(got[m] - want) / np.maximum(np.abs(want), 1e-300) * (np.abs(want) > 1e-9))))
    ts.close(); cases += 1
print('fuzz_geo: %d cases in %.0f s, all match the oracle; worst relative difference %.2e (seed %d)' % (cases, budget, worst, seed0))
