"""treecl_b200: B200-native engine for treeCl's all-pairs inter-tree distance matrix.

Drop-in surface (same names and arguments as kgori/treeCl for this path):
    treecl_b200.Collection(...).get_inter_tree_distances('rf'|'wrf'|'euc'|'geo', ...) -> DistanceMatrix
    treecl_b200.treedist.{rf,wrf,euc,geo}dist / *_matrix
    treecl_b200.Tree, treecl_b200.DistanceMatrix, treecl_b200.parutils.*JobHandler
    treecl_b200.kendallcolijn.KendallColijn / kendall_colijn_matrix (treeCl/utils/kendallcolijn.py)
    treecl_b200.bootstrap.query_reference_distances (the B x N block behind treeCl/bootstrap.py:174-218)
    DistanceMatrix.embedding(dims, 'cmds') (classical MDS on the GPU, treeCl/distance_matrix.py:301-315)
    treecl_b200.spectral.Spectral (.affinity, .spectral_embedding_: treeCl/clustering.py:125-303 for the non-ESTIMATE options)
Everything is computed by libtreedist_cuda.so (hand-written sm_100a kernels); there is no CPU fallback.
"""
from . import bootstrap, kendallcolijn, parutils, spectral, tasks, treedist  # noqa: F401
from .collection import Collection, Record  # noqa: F401
from .distance_matrix import CoordinateMatrix, DistanceMatrix  # noqa: F401
from .errors import OptionError  # noqa: F401
from .tree import Tree  # noqa: F401

__version__ = '0.1.0'
