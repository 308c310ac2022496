"""DistanceMatrix: the container get_inter_tree_distances returns (treeCl/distance_matrix.py:454-499,557-566).

Only the parts the distance path and its direct consumers touch are mirrored: construction from a square array with
names on index and columns, csv io, names, values - and, as the first consumer moved to the GPU, the classical
MDS embedding (`embedding(dimensions, 'cmds')`, treeCl/distance_matrix.py:301-315,517-555).  The other embeddings / affinities stay in
treeCl proper: they take `self.to_array()` unchanged.
"""
import sys

import numpy as np
import pandas as pd


class CoordinateMatrix(object):
    """treeCl/distance_matrix.py:443-451: coordinates with the record names on the index."""

    def __init__(self, array, names=None):
        nrow, ncol = array.shape
        self.df = pd.DataFrame(array, index=names, columns=['PC{}'.format(i) for i in range(1, ncol + 1)])

    def __repr__(self):
        return self.df.__repr__()

    @property
    def values(self):
        return self.df.values

    @property
    def shape(self):
        return self.df.shape

    def to_array(self):
        return self.df.values


class DistanceMatrix(object):
    def __init__(self):
        self.df = pd.DataFrame()

    def embedding(self, dimensions, method='cmds', device=0, **kwargs):
        """Embed the matrix in a coordinate space (treeCl/distance_matrix.py:517-555).  'cmds' (classical multidimensional scaling) runs on
        the GPU: double-centring and the `dimensions` leading eigenpairs by blocked subspace iteration (libtreedist_cuda td_cmds_host);
        eigenvectors are defined up to sign, as with numpy.linalg.eigh.  The other methods of the reference are sklearn wrappers that
        take to_array() unchanged and are not rebuilt here."""
        from . import _lib
        from .errors import optioncheck
        optioncheck(method, ['cmds', 'kpca', 'mmds', 'nmmds', 'spectral', 'tsne'])
        if method != 'cmds':
            raise NotImplementedError("embedding method '{}' is an sklearn wrapper in treeCl (treeCl/distance_matrix.py:318-395) and is "
                                      "not part of the GPU path; only 'cmds' is".format(method))
        if kwargs.get('additive_correct'):
            raise NotImplementedError('additive_correct=True (treeCl/distance_matrix.py:76-108) is not part of the GPU path')
        coords, _, _ = _lib.cmds_host(self.to_array(), dimensions, device=device)
        return CoordinateMatrix(coords, names=self.df.index)

    def __repr__(self):
        return self.df.__repr__()

    def __eq__(self, other):
        return bool((np.abs(self.sort().values - other.sort().values) < 1e-10).all())

    @property
    def shape(self):
        return self.df.shape

    @property
    def values(self):
        return self.df.values

    def to_array(self):
        return self.df.values

    @classmethod
    def from_csv(cls, filename, **kwargs):
        new_instance = cls()
        if 'index_col' not in kwargs:
            kwargs['index_col'] = 0
        new_instance.df = pd.read_csv(filename, **kwargs)
        return new_instance

    def to_csv(self, filename, **kwargs):
        self.df.to_csv(filename)

    @classmethod
    def from_array(cls, array, names=None):
        nrow, ncol = array.shape
        if nrow != ncol:
            raise ValueError('Not a square array')
        new_instance = cls()
        new_instance.df = pd.DataFrame(array, copy=False)          # (pandas 3 would copy the N x N array otherwise)
        try:
            new_instance.set_names(names)
        except ValueError as err:     # distance_matrix.py:486-488: reported, not raised
            sys.stderr.write(str(err))
        return new_instance

    def get_names(self):
        return [str(x) for x in self.df.index]

    def set_names(self, names):
        if names is None:
            return
        if len(names) != self.df.index.size:
            raise ValueError('Expected {} names, got {}'.format(self.df.index.size, len(names)))
        self.df.index = self.df.columns = names

    def reorder(self, new_order):
        newobj = self.__class__()
        newobj.df = self.df.reindex(columns=new_order, index=new_order)
        return newobj

    def sort(self):
        order = self.df.index[self.df.index.argsort()]
        return self.reorder(order)
