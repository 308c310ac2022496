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
        out = {m: _lib.result_empty(total) for m in metrics}
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


def folded_row_ranges(n, parts, align=64):
    """Rows 0..n-1 of the upper triangle for `parts` ranks, every rank a block of LONG rows from the top and a block of SHORT rows from
    the bottom (cuts at multiples of `align` rows = whole tile rows).  The top blocks have equal heights; the bottom block of a rank is
    sized so that its pair count reaches total / parts - short rows make that a fine adjustment - and the last rank takes the middle.
    Contiguous equal-pair ranges would give rank 0 a few very long rows (a grid too small for its GPU) and the last rank many short
    ones.  Returns, per rank, a list of one or two (row_begin, row_end) ranges that together tile [0, n)."""
    def pairs(a, b):                                        # pairs of rows [a, b)
        return (b - a) * (2 * n - a - b - 1) // 2
    if parts <= 1 or n < 2 * parts * align:
        return [[r] for r in balanced_row_ranges(n, parts)]
    units = -(-n // align)
    row = lambda u: min(n, u * align)
    total, out = pairs(0, n), []
    top, bottom = 0, units                                  # tile rows [top, bottom) are not assigned yet
    for r in range(parts - 1):
        h = (units * (r + 1)) // (2 * parts) - (units * r) // (2 * parts)
        t0, t1 = top, min(top + h, bottom)
        want = total * (r + 1) // parts - sum(pairs(a, b) for rr in out for a, b in rr)       # keeps the running total on target
        b0 = bottom
        while b0 - 1 > t1 and pairs(row(t0), row(t1)) + pairs(row(b0 - 1), row(bottom)) <= want:
            b0 -= 1
        # one more tile row if that lands closer to the target
        if b0 - 1 > t1:
            under = want - pairs(row(t0), row(t1)) - pairs(row(b0), row(bottom))
            over = pairs(row(t0), row(t1)) + pairs(row(b0 - 1), row(bottom)) - want
            if over < under:
                b0 -= 1
        out.append([(a, b) for a, b in ((row(t0), row(t1)), (row(b0), row(bottom))) if b > a])
        top, bottom = t1, b0
    out.append([(row(top), row(bottom))] if bottom > top else [])
    return out


def distance_square(newicks, metric, normalise=False, min_overlap=4, overlap_fail_value=0.0, device=None):
    """N x N matrix of one metric on one GPU: squareform() happens on the device (td_pairwise_square)."""
    ts = _lib.TreeSet(list(newicks))
    try:
        return ts.pairwise_square(metric, normalise, min_overlap, overlap_fail_value, default_device() if device is None else device)
    finally:
        ts.close()


def distance_condensed(newicks, metric, normalise=False, min_overlap=4, overlap_fail_value=0.0, devices=None):
    ts = _lib.TreeSet(list(newicks))
    try:
        return pairwise_condensed(ts, (metric,), normalise, min_overlap, overlap_fail_value, devices)[metric]
    finally:
        ts.close()
