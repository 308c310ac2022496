"""Kendall-Colijn tree metric: mirror of treeCl/utils/kendallcolijn.py on libtreedist_cuda.

The reference computes, per tree, the vectors m (edges from the root to the MRCA of every leaf pair, then 1 per leaf) and M (path
length to the MRCA, then the leaf edge lengths) with dendropy, and the distance of two trees as the Euclidean distance of
v = (1 - lbda) m + lbda M.  Here the vectors are built by the host encoder (td_encode_kendall_colijn) and all pairs go through the
same fp64 tensor-core Euclidean kernel as the 'euc' metric: a set of Kendall-Colijn vectors is uploaded as trees without internal
edges whose leaf-edge lengths are the vector entries.
"""
import numpy as np
from scipy.spatial.distance import squareform

from . import _lib, engine
from .tree import Tree


def _tree(t):
    return t if isinstance(t, Tree) else Tree(t)


class KendallColijn(object):
    """KendallColijn(tree): same constructor and methods as the reference class (kendallcolijn.py:5-95)."""

    def __init__(self, tree):
        self.tree = _tree(tree)
        self.little_m = self._vector(0.0)          # kendallcolijn.py:17-19
        self.big_m = self._vector(1.0)

    def _vector(self, lbda):
        ts = _lib.TreeSet.kendall_colijn([self.tree.newick], lbda)
        try:
            return ts.get(_lib.TD_GET_LEAFLEN, np.float64).copy()
        finally:
            ts.close()

    def get_vector(self, lbda=0.5):
        """(1 - lbda) m + lbda M (kendallcolijn.py:74-79)."""
        return (1 - lbda) * self.little_m + lbda * self.big_m

    def get_distance(self, other, lbda=0.5, min_overlap=4):
        """Euclidean distance of the two vectors; differing leaf sets are pruned to the intersection first, and fewer than
        min_overlap common leaves give 0 (kendallcolijn.py:81-95)."""
        t1, t2 = self.tree, other.tree
        if t1 ^ t2:
            common = t1 & t2
            if len(common) < min_overlap:
                return 0
            t1 = t1.prune_to_subset(common) if t1.labels != common else t1
            t2 = t2.prune_to_subset(common) if t2.labels != common else t2
        return float(kendall_colijn_condensed([t1.newick, t2.newick], lbda)[0])


def kendall_colijn_condensed(newicks, lbda=0.5, devices=None):
    """Condensed Kendall-Colijn distance vector (scipy order) of trees on ONE leaf set, computed on the GPU(s)."""
    ts = _lib.TreeSet.kendall_colijn(list(newicks), lbda)
    try:
        return engine.pairwise_condensed(ts, ('euc',), devices=devices)['euc']
    finally:
        ts.close()


def kendall_colijn_matrix(trees, lbda=0.5, devices=None):
    """Square ndarray of all pairwise Kendall-Colijn distances (trees: Tree objects or Newick strings, one leaf set)."""
    return squareform(kendall_colijn_condensed([_tree(t).newick for t in trees], lbda, devices))
