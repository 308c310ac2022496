"""Query x reference tree distances for the bootstrap-placement routines of treeCl/bootstrap.py:174-218.

run_optimise_bootstrap_coords, run_out_of_sample_mds and run_analytical_fit each fill, bootstrap tree by bootstrap tree,
    ref_dists = np.array([task(tree, ref_tree, False) for ref_tree in ref_trees])          # task = tasks._fast_geo by default
i.e. one row of the B x N block of distances between the bootstrap (query) trees and the reference trees, with a doubly nested Python
loop over tree_distance calls.  query_reference_distances() returns the whole block from one rectangular kernel launch
(td_pairwise_rect); the fitting that follows (OptimiseDistanceFit / OutOfSampleMDS / AnalyticalFit, NumPy / SciPy on B rows) stays in
treeCl and takes the rows unchanged.

The reference passes an explicit `rooted` flag (default False) to PhyloTree here; this encoder infers rootedness from the root's degree,
as Tree.phylotree does on the matrix path (treeCl/tree.py:858-862) - identical for the trifurcating-root ML trees treeCl produces.
"""
import numpy as np

from . import _lib, engine

__all__ = ['query_reference_distances']


def _newicks(trees):
    if hasattr(trees, 'trees'):                 # a Collection
        trees = trees.trees
    return [t if isinstance(t, (str, bytes)) else t.newick for t in trees]


def query_reference_distances(boot_collection, ref_collection, metric='geo', normalise=False, device=None):
    """(len(boot), len(ref)) ndarray of `metric` distances; row i is the `ref_dists` vector the reference computes for bootstrap tree i."""
    q, r = _newicks(boot_collection), _newicks(ref_collection)
    device = engine.default_device() if device is None else device
    tq, tr = _lib.TreeSet(q), _lib.TreeSet(r)
    try:
        same_taxa = tq.get(_lib.TD_GET_LABELS, None) == tr.get(_lib.TD_GET_LABELS, None)
        if same_taxa and tq.info.uniform and tr.info.uniform:
            # two separately encoded sets over one taxon table: the library joins the dictionaries and runs the rectangular kernels
            return _lib.pairwise_rect(tq, tr, (metric,), normalise, device)[metric]
    finally:
        tq.close(); tr.close()
    # differing leaf sets somewhere: encode all trees together and take the block of the square matrix (restrict-then-join / groups)
    ts = _lib.TreeSet(q + r)
    try:
        sq = ts.pairwise_square(metric, normalise, device=device)
        return np.ascontiguousarray(sq[:len(q), len(q):])
    finally:
        ts.close()
