"""Front end of treeCl's spectral clustering on the GPU: the slice of `treeCl.clustering.Spectral` (treeCl/clustering.py:125-303) that turns a
distance matrix into an affinity matrix and an embedding.

    Spectral(dm, pruning_option, scale_option, manual_pruning, manual_scale)      constructor and options as in the reference
        .affinity                      Spectral.decompose: kdists / kmask / kscale / affinity (treeCl/distance_matrix.py:32-55,153-186)
        .spectral_embedding_(n)        algo ZELNIKMANOR: laplace + eigen + normalise_rows (treeCl/clustering.py:288-303)

Pruning NONE / MANUAL and scale MEDIAN / MANUAL run in libtreedist_cuda (td_affinity_host, td_spectral_embedding_host).  The ESTIMATE
options (a host-side search for the smallest connected neighbour graph, binsearch_mask) and the sklearn-backed embeddings
(spectral_embedding, kpca) are not part of the GPU path; the k-means / GMM step that follows an embedding takes the coordinates unchanged.
"""
import ctypes

import numpy as np

from . import _lib
from .distance_matrix import CoordinateMatrix, DistanceMatrix
from .errors import OptionError


class options(object):
    """treeCl/clustering.py:32-38"""
    PRUNING_NONE, PRUNING_ESTIMATE, PRUNING_MANUAL = 0, 1, 2
    LOCAL_SCALE_MEDIAN, LOCAL_SCALE_ESTIMATE, LOCAL_SCALE_MANUAL = 3, 4, 5
    reverse = {0: 'PRUNING_NONE', 1: 'PRUNING_ESTIMATE', 2: 'PRUNING_MANUAL', 3: 'LOCAL_SCALE_MEDIAN', 4: 'LOCAL_SCALE_ESTIMATE',
               5: 'LOCAL_SCALE_MANUAL'}


class Spectral(object):
    def __init__(self, dm, pruning_option=options.PRUNING_NONE, scale_option=options.LOCAL_SCALE_MEDIAN, manual_pruning=None,
                 manual_scale=None, logic='or', device=0):
        if isinstance(dm, np.ndarray):
            dm = DistanceMatrix.from_array(dm)
        if not isinstance(dm, DistanceMatrix):
            raise ValueError('Distance matrix should be a numpy array or treeCl.DistanceMatrix')
        for opt in (pruning_option, scale_option):
            if opt not in options.reverse:
                raise OptionError(opt, list(options.reverse.values()))
        if pruning_option == options.PRUNING_ESTIMATE or scale_option == options.LOCAL_SCALE_ESTIMATE:
            raise NotImplementedError('the ESTIMATE options (binsearch_mask, treeCl/distance_matrix.py:258-282) are not part of the GPU path')
        n = dm.df.shape[0]
        self.dm, self._device = dm, device
        self._prune_k = int(manual_pruning) if pruning_option == options.PRUNING_MANUAL else 0
        self._scale_k = int(manual_scale) if scale_option == options.LOCAL_SCALE_MANUAL else 0
        for k in (self._prune_k, self._scale_k):
            if k and not 2 <= k <= n:
                raise ValueError('neighbour counts must lie in [2, {}]'.format(n))
        self._and = logic in ('and', '&')
        self._affinity = None

    @property
    def affinity(self):
        """Spectral.decompose(): the affinity matrix with a unit diagonal."""
        if self._affinity is None:
            sq = np.ascontiguousarray(self.dm.to_array(), dtype=np.float64)
            out = np.empty_like(sq)
            _lib.check(_lib.load().td_affinity_host(sq.ctypes.data, len(sq), self._prune_k, int(self._and), self._scale_k, out.ctypes.data, self._device))
            self._affinity = out
        return self._affinity

    def spectral_embedding_(self, n):
        """Coordinates with unit-length rows from the leading eigenvectors of D^-1/2 A D^-1/2 (treeCl/clustering.py:288-303)."""
        sq = np.ascontiguousarray(self.dm.to_array(), dtype=np.float64)
        coords = np.empty((len(sq), int(n)))
        vals = np.empty(int(n))
        it, res = ctypes.c_int(), ctypes.c_double()
        _lib.check(_lib.load().td_spectral_embedding_host(sq.ctypes.data, len(sq), int(n), self._prune_k, int(self._and), self._scale_k,
                                                          coords.ctypes.data, vals.ctypes.data, 0, 0.0, ctypes.byref(it), ctypes.byref(res), self._device))
        self.eigenvalues = vals
        return CoordinateMatrix(coords, names=self.dm.df.index)
