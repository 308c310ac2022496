"""Randomised cross-check of the rf/wrf/euc kernels on one GPU: postings pipeline ('split') against the merge kernel and, for small
sets, against the oracle - random sizes, taxa, sharing levels, normalisation, row ranges and rectangular splits.
usage: python tools/fuzz_parity.py [seconds] [seed]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from treecl_b200 import _lib, synth
from oracle import oracle

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 12345)
METRICS = ('rf', 'wrf', 'euc')


def run(ts, kernel, n, norm, r0, r1):
    os.environ['TREEDIST_PAIR_KERNEL'] = kernel
    return ts.pairwise(METRICS, normalise=norm, row_begin=r0, row_end=r1)


def run_rect(ts, kernel, n, split, norm):
    os.environ['TREEDIST_PAIR_KERNEL'] = kernel
    outs = {m: torch.full((split * (n - split),), -7.0, dtype=torch.float64, device='cuda') for m in METRICS}
    ts.upload(0)
    ts.pairwise_rect_device(split, METRICS, {m: outs[m].data_ptr() for m in METRICS}, normalise=norm)
    torch.cuda.synchronize()
    return {m: outs[m].cpu().numpy() for m in METRICS}


def close(a, b):
    return np.allclose(a, b, rtol=1e-9, atol=1e-12)


t_end, cases, worst = time.time() + budget, 0, 0.0
while time.time() < t_end:
    n = int(rng.choice([2, 3, 17, 63, 64, 65, 100, 129, 257, 400, 700]))
    taxa = int(rng.choice([4, 5, 9, 16, 33, 50, 64, 65, 100, 131]))
    kind = rng.choice(['uniform', 'structured', 'clustered'])
    seed = int(rng.integers(1 << 30))
    if kind == 'uniform':
        trees = synth.uniform_trees(n, taxa, seed=seed)
    elif kind == 'structured':
        trees = synth.structured_trees(n, taxa, n_classes=int(rng.integers(1, 9)), class_nni=int(rng.integers(0, 12)), lam=float(rng.uniform(0, 6)), seed=seed)
    else:       # a few exact duplicates and near-duplicates inside a random set: exercises the exact fix-up
        base = synth.uniform_trees(max(n // 2, 1), taxa, seed=seed)
        trees = (base + base)[:n] if n > 1 else base
        trees = list(rng.permutation(np.array(trees, dtype=object)))
        n = len(trees)
    norm = bool(rng.integers(2))
    ts = _lib.TreeSet(trees)
    r0 = int(rng.integers(0, n - 1)) if rng.integers(2) else 0
    r1 = int(rng.integers(r0 + 1, n + 1)) if rng.integers(2) else n
    a, b = run(ts, 'split', n, norm, r0, r1), run(ts, 'merge', n, norm, r0, r1)
    for m in METRICS:
        ok = np.array_equal(a[m], b[m]) if m == 'rf' else close(a[m], b[m])
        assert ok, ('triangular', m, n, taxa, kind, seed, norm, r0, r1)
        if m != 'rf' and a[m].size:
            worst = max(worst, float(np.max(np.abs(a[m] - b[m]) / np.maximum(np.abs(b[m]), 1e-300) * (np.abs(b[m]) > 1e-9))))
    if n <= 129 and taxa <= 65 and r0 == 0 and r1 == n:
        for m in METRICS:
            want = oracle.matrix(trees, m, normalise=norm, nthreads=4)
            assert (np.array_equal(a[m], want) if m == 'rf' else close(a[m], want)), ('oracle', m, n, taxa, kind, seed, norm)
    if n >= 3:
        split = int(rng.integers(1, n))
        ra, rb = run_rect(ts, 'split', n, split, norm), run_rect(ts, 'merge', n, split, norm)
        for m in METRICS:
            assert (np.array_equal(ra[m], rb[m]) if m == 'rf' else close(ra[m], rb[m])), ('rect', m, n, taxa, kind, seed, norm, split)
    ts.close()
    cases += 1
print('fuzz ok: %d cases, worst relative difference split vs merge %.2e' % (cases, worst))
