"""CPU restatement (NumPy) of the reference's classical-MDS chain - test infrastructure, like the rest of oracle/.

    double_centre             treeCl/distance_matrix.py:58-74
    eigen                     treeCl/distance_matrix.py:285-298   (np.linalg.eigh, eigenvalues descending)
    classical_mds             treeCl/distance_matrix.py:301-315   (_embedding_classical_mds, additive_correct=False)

Parity status: PINNED - tests/golden/mds_golden.json holds the outputs of the reference's own functions (the module was loaded from
/root/reference by tests/golden/make_mds_golden.py) and tests/test_oracle_golden.py checks this restatement against them.
"""
import numpy as np

__all__ = ['double_centre', 'eigen', 'classical_mds']


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
