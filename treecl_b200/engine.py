"""Host-side driver of libtreedist_cuda for whole matrices: encode once, shard rows over GPUs, gather condensed.

This is the block that replaces treeCl/collection.py:447-455 (task table, argument generator, JobHandler map).
"""
import concurrent.futures
import os

import numpy as np

from . import _lib

METRICS = ('rf', 'wrf', 'euc', 'geo')


def default_device():
    for key in ('TREECL_B200_DEVICE', 'LOCAL_RANK'):
        if os.environ.get(key, '') != '':
            return int(os.environ[key])
    return 0


def balanced_row_ranges(n, parts):
    """Split rows 0..n-1 of the upper triangle into `parts` contiguous ranges with (almost) equal pair counts."""
    total = n * (n - 1) // 2
    bounds, row = [0], 0
    for k in range(1, parts):
        target = total * k // parts
        # smallest row whose condensed offset reaches the target (offset is monotone in row)
        lo, hi = row, n
        while lo < hi:
            mid = (lo + hi) // 2
            if _lib.condensed_offset(n, mid) < target:
                lo = mid + 1
            else:
                hi = mid
        row = lo
        bounds.append(row)
    bounds.append(n)
    return [(bounds[k], bounds[k + 1]) for k in range(parts)]


def pairwise_condensed(treeset, metrics, normalise=False, min_overlap=4, overlap_fail_value=0.0, devices=None,
                       out=None):
    """Condensed vectors (scipy order = itertools.combinations order, treeCl/tasks.py:651-653) for `metrics`.

    devices: list of CUDA device ordinals; rows are sharded over them with balanced pair counts, the encoded set is
    replicated, and each device streams its slice into the shared host buffers.  No collective is involved.
    """
    n = treeset.n_trees
    total = n * (n - 1) // 2
    if devices is None:
        devices = [default_device()]
    if out is None:
        out = {m: np.empty(total, dtype=np.float64) for m in metrics}
    if total == 0:
        return out
    ranges = balanced_row_ranges(n, len(devices))

    def run(k):
        r0, r1 = ranges[k]
        if r1 <= r0:
            return
        p0 = _lib.condensed_offset(n, r0)
        p1 = _lib.condensed_offset(n, r1)
        if p1 <= p0:
            return
        treeset.pairwise(metrics, normalise, min_overlap, overlap_fail_value, r0, r1, devices[k],
                         out={m: out[m][p0:p1] for m in metrics})

    if len(devices) == 1:
        run(0)
    else:
        with concurrent.futures.ThreadPoolExecutor(max_workers=len(devices)) as ex:
            list(ex.map(run, range(len(devices))))
    return out


def distance_condensed(newicks, metric, normalise=False, min_overlap=4, overlap_fail_value=0.0, devices=None):
    ts = _lib.TreeSet(list(newicks))
    try:
        return pairwise_condensed(ts, (metric,), normalise, min_overlap, overlap_fail_value, devices)[metric]
    finally:
        ts.close()
