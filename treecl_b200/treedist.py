"""Drop-in for treeCl/treedist.py: same names, arguments and return types, computed by libtreedist_cuda.

    eucdist / geodist / rfdist / wrfdist (t1, t2, normalise, min_overlap=4, overlap_fail_value=0) -> float
    *_matrix(trees, normalise, min_overlap=4, overlap_fail_value=0, show_progress=True)           -> N x N ndarray

Leaf-set semantics follow _generic_distance_calc (treeCl/treedist.py:30-70): if the two label sets differ and fewer
than min_overlap labels are shared the result is overlap_fail_value, otherwise both trees are restricted to the
shared labels first.  `t1`, `t2`, `trees` may be treecl_b200.Tree objects or Newick strings.
"""
import functools

import scipy.spatial.distance

from . import engine

__all__ = ["eucdist", "eucdist_matrix", "geodist", "geodist_matrix", "rfdist", "rfdist_matrix", "wrfdist",
           "wrfdist_matrix"]


def _newick(t):
    return t if isinstance(t, (str, bytes)) else t.newick


def _generic_distance_calc(metric, t1, t2, normalise, min_overlap=4, overlap_fail_value=0):
    """One pair.  A two-tree set is encoded and a one-tile grid is launched; there is no CPU path."""
    res = engine.distance_condensed([_newick(t1), _newick(t2)], metric, normalise, min_overlap,
                                    float(overlap_fail_value))
    return float(res[0])


def _generic_matrix_calc(metric, trees, normalise, min_overlap=4, overlap_fail_value=0, show_progress=True):
    """All pairs -> squareform ndarray (treeCl/treedist.py:73-109).  show_progress is accepted and ignored."""
    trees = list(trees)
    res = engine.distance_condensed([_newick(t) for t in trees], metric, normalise, min_overlap,
                                    float(overlap_fail_value))
    return scipy.spatial.distance.squareform(res)


eucdist = functools.partial(_generic_distance_calc, 'euc')
geodist = functools.partial(_generic_distance_calc, 'geo')
rfdist = functools.partial(_generic_distance_calc, 'rf')
wrfdist = functools.partial(_generic_distance_calc, 'wrf')
eucdist_matrix = functools.partial(_generic_matrix_calc, 'euc')
geodist_matrix = functools.partial(_generic_matrix_calc, 'geo')
rfdist_matrix = functools.partial(_generic_matrix_calc, 'rf')
wrfdist_matrix = functools.partial(_generic_matrix_calc, 'wrf')
