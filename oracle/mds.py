"""CPU restatement (NumPy) of the reference's classical-MDS chain - test infrastructure, like the rest of oracle/.

    double_centre             treeCl/distance_matrix.py:58-74
    eigen                     treeCl/distance_matrix.py:285-298   (np.linalg.eigh, eigenvalues descending)
    classical_mds             treeCl/distance_matrix.py:301-315   (_embedding_classical_mds, additive_correct=False)

Parity status: PINNED - tests/golden/mds_golden.json holds the outputs of the reference's own functions (the module was loaded from
/root/reference by tests/golden/make_mds_golden.py) and tests/test_oracle_golden.py checks this restatement against them.
"""
import numpy as np

__all__ = ['double_centre', 'eigen', 'classical_mds', 'kdists', 'kmask', 'kscale', 'affinity', 'laplace', 'normalise_rows',
           'spectral_affinity', 'spectral_embedding']


def double_centre(matrix, square_input=True):
    m = np.array(matrix, dtype=np.float64, copy=True)
    if square_input:
        m **= 2
    rows = m.shape[0]
    cm = np.mean(m, axis=0)
    rm = np.mean(m, axis=1).reshape((rows, 1))
    gm = np.mean(cm)
    m -= rm + cm - gm
    m /= -2
    return m


def eigen(matrix):
    vals, vecs = np.linalg.eigh(matrix)
    ind = vals.argsort()[::-1]
    return vals[ind], vecs[:, ind]


def classical_mds(matrix, dimensions=3):
    vals, vecs = eigen(double_centre(matrix))
    return vecs[:, :dimensions].dot(np.diag(np.sqrt(np.abs(vals[:dimensions])))), vals


# ---- Spectral manager front end (pruning NONE / MANUAL, scale MEDIAN / MANUAL) ---------------------------------------------
#     kdists / kindex / kmask / kscale     treeCl/distance_matrix.py:153-186
#     affinity                             treeCl/distance_matrix.py:32-55   (without the isconnected assertion)
#     laplace                              treeCl/distance_matrix.py:189-207 (default form)
#     normalise_rows                       treeCl/distance_matrix.py:141-150
#     spectral_affinity / spectral_embedding   the call order of treeCl/clustering.py:171-228 (decompose) and :288-303 (spectral_embedding_)
def kdists(matrix, k=7):
    matrix = np.asarray(matrix, dtype=np.float64)
    ix = (np.arange(len(matrix)), matrix.argsort(axis=0)[k])
    return matrix[ix][np.newaxis].T


def kmask(matrix, k=7, logic='or'):
    mask = (matrix <= kdists(matrix, k))
    return (mask | mask.T) if logic in ('or', '|') else (mask & mask.T)


def kscale(matrix, k=7):
    d = kdists(matrix, k)
    return d.dot(d.T)


def affinity(matrix, mask, scale):
    ix = np.where(np.logical_not(mask))
    with np.errstate(divide='ignore', invalid='ignore'):
        scaled = -np.asarray(matrix, dtype=np.float64) ** 2 / scale
    scaled[np.where(np.isnan(scaled))] = -1.0
    aff = np.exp(scaled)
    aff[ix] = 0.
    aff.flat[::len(aff) + 1] = 0.
    return aff


def laplace(affinity_matrix):
    diagonal = affinity_matrix.sum(axis=1) - affinity_matrix.diagonal()
    diagonal[diagonal <= 1e-10] = 1
    diagonal = np.sqrt(diagonal)
    return affinity_matrix / diagonal / diagonal[:, np.newaxis]


def normalise_rows(matrix):
    lengths = np.apply_along_axis(np.linalg.norm, 1, matrix)
    lengths[lengths == 0] = 1
    return matrix / lengths[:, np.newaxis]


def spectral_affinity(matrix, prune_k=0, logic='or', scale_k=0):
    matrix = np.asarray(matrix, dtype=np.float64)
    mask = np.ones(matrix.shape, dtype=bool) if prune_k == 0 else kmask(matrix, prune_k, logic)
    if scale_k == 0:
        dist = np.median(matrix, axis=1)
        scale = np.outer(dist, dist)
    else:
        scale = kscale(matrix, scale_k)
    aff = affinity(matrix, mask, scale)
    aff.flat[::len(aff) + 1] = 1.0
    return aff


def spectral_embedding(matrix, dimensions, prune_k=0, logic='or', scale_k=0):
    aff = spectral_affinity(matrix, prune_k, logic, scale_k)
    aff.flat[::len(aff) + 1] = 0
    vals, vecs = eigen(laplace(aff))
    return normalise_rows(vecs[:, :dimensions]), vals
