"""DistanceMatrix: the container get_inter_tree_distances returns (treeCl/distance_matrix.py:454-499,557-566).

Only the parts the distance path and its direct consumers touch are mirrored: construction from a square array with
names on index and columns, csv io, names, values.  Embeddings / affinities stay in treeCl proper.
"""
import sys

import numpy as np
import pandas as pd


class DistanceMatrix(object):
    def __init__(self):
        self.df = pd.DataFrame()

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
