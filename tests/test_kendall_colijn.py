"""Kendall-Colijn metric (treeCl/utils/kendallcolijn.py): oracle against hand-derived values of the published
definition (the reference has no test for it: parity UNPINNED beyond that), host encoder against the oracle, GPU matrix against the
oracle."""
import numpy as np
import pytest

@pytest.fixture(scope='module')
def td(lib_built):
    import treecl_b200
    from treecl_b200 import _lib
    assert _lib.device_count() > 0, 'these tests need a GPU; the product has no CPU fallback'
    return treecl_b200


T1 = '((a:1,b:2):3,(c:4,d:5):6);'
T2 = '((a:1,c:2):3,(b:4,d:5):6);'


def test_oracle_vectors_by_hand(oracle_mod):
    # pairs in combinations order of the sorted labels: ab ac ad bc bd cd, then the leaves a b c d
    assert np.array_equal(oracle_mod.kc_vector(T1, 0.0), [1, 0, 0, 0, 0, 1, 1, 1, 1, 1])          # m: edges root -> MRCA; 1 per leaf
    assert np.array_equal(oracle_mod.kc_vector(T1, 1.0), [3, 0, 0, 0, 0, 6, 1, 2, 4, 5])          # M: path length; leaf edge lengths
    assert np.array_equal(oracle_mod.kc_vector(T2, 1.0), [0, 3, 0, 0, 6, 0, 1, 4, 2, 5])
    # lambda = 0.5: v1 - v2 = [2, -2, 0, 0, -3.5, 3.5, 0, -1, 1, 0]  ->  sqrt(34.5)
    assert abs(oracle_mod.kc_pair(T1, T2, 0.5) - np.sqrt(34.5)) < 1e-15
    assert oracle_mod.kc_pair(T1, T1, 0.5) == 0.0
    # topology only (lambda = 0): sqrt(1 + 1 + 1 + 1)
    assert oracle_mod.kc_pair(T1, T2, 0.0) == 2.0


def test_oracle_partial_overlap(oracle_mod):
    a = '(((a:1,b:1):1,c:1):1,(d:1,e:1):1,f:1);'
    b = '(((a:1,c:1):1,b:1):1,(d:1,g:1):1,f:1);'
    # common leaves a b c d f: both trees are pruned to them first (kendallcolijn.py:86-91)
    keep = ['a', 'b', 'c', 'd', 'f']
    want = oracle_mod.kc_pair(oracle_mod.prune_newick(a, keep), oracle_mod.prune_newick(b, keep), 0.5)
    assert oracle_mod.kc_pair(a, b, 0.5) == want and want > 0
    assert oracle_mod.kc_pair(a, b, 0.5, min_overlap=6) == 0.0                                     # kendallcolijn.py:84-85


@pytest.mark.parametrize('kind,n_trees,n_taxa', [('uniform', 12, 5), ('uniform', 20, 17), ('structured', 16, 30), ('uniform', 6, 66)])
@pytest.mark.parametrize('lbda', [0.0, 0.3, 1.0])
def test_encoder_vectors_match_oracle(lib_built, oracle_mod, kind, n_trees, n_taxa, lbda):
    from treecl_b200 import _lib, synth
    trees = synth.uniform_trees(n_trees, n_taxa, seed=21) if kind == 'uniform' else \
        synth.structured_trees(n_trees, n_taxa, n_classes=3, class_nni=3, lam=1.0, seed=22)
    ts = _lib.TreeSet.kendall_colijn(trees, lbda)
    inf = ts.info
    assert inf.n_taxa == n_taxa * (n_taxa - 1) // 2 + n_taxa and inf.max_splits == 0 and inf.uniform == 1
    got = ts.get(_lib.TD_GET_LEAFLEN, np.float64).reshape(n_trees, -1)
    want = np.stack([oracle_mod.kc_vector(t, lbda) for t in trees])
    assert np.allclose(got, want, rtol=1e-14, atol=1e-14)


def test_encoder_handles_rooted_and_multifurcating_trees(lib_built, oracle_mod):
    from treecl_b200 import _lib
    trees = [T1, T2, '(a:1,b:2,(c:4,d:5):6);', '((a:1,b:2,c:3):1,d:2);', '(a:0.5,(b:1,(c:2,d:3):4):5);']
    for lbda in (0.0, 0.5, 1.0):
        got = _lib.TreeSet.kendall_colijn(trees, lbda).get(_lib.TD_GET_LEAFLEN, np.float64).reshape(len(trees), -1)
        want = np.stack([oracle_mod.kc_vector(t, lbda) for t in trees])
        assert np.allclose(got, want, rtol=1e-15, atol=1e-15)


def test_encoder_refuses_mixed_leaf_sets_and_bad_lambda(lib_built):
    from treecl_b200 import _lib
    with pytest.raises(NotImplementedError):
        _lib.TreeSet.kendall_colijn([T1, '((a:1,b:2):3,(c:4,e:5):6);'])
    with pytest.raises(NotImplementedError):
        _lib.TreeSet.kendall_colijn([T1, '((a:1,b:2):3,c:4);'])
    with pytest.raises(ValueError):
        _lib.TreeSet.kendall_colijn([T1, T2], lbda=1.5)
    with pytest.raises(ValueError):
        _lib.TreeSet.kendall_colijn([T1, '((a:1,b:2):3,(c:4,d:5)'])


@pytest.mark.gpu
@pytest.mark.parametrize('kind,n_trees,n_taxa', [('uniform', 2, 4), ('uniform', 70, 9), ('uniform', 150, 20), ('structured', 130, 33),
                                                 ('uniform', 40, 50)])
@pytest.mark.parametrize('lbda', [0.0, 0.5, 1.0])
def test_gpu_matrix_matches_oracle(td, oracle_mod, kind, n_trees, n_taxa, lbda):
    from treecl_b200 import kendallcolijn, synth
    trees = synth.uniform_trees(n_trees, n_taxa, seed=31) if kind == 'uniform' else \
        synth.structured_trees(n_trees, n_taxa, n_classes=3, class_nni=3, lam=1.0, seed=32)
    got = kendallcolijn.kendall_colijn_condensed(trees, lbda)
    want = oracle_mod.kc_matrix(trees, lbda)
    assert np.allclose(got, want, rtol=1e-9, atol=1e-12)
    # identical trees are exactly 0 apart (the exact fix-up path of the dot-product kernel)
    both = kendallcolijn.kendall_colijn_condensed([trees[0], trees[0], trees[1]], lbda)
    assert both[0] == 0.0 and abs(both[1] - want[0]) <= 1e-9 * max(want[0], 1e-300) + 1e-12


@pytest.mark.gpu
def test_gpu_class_mirror(td, oracle_mod):
    from treecl_b200 import kendallcolijn
    a = '(((a:1,b:1):1,c:1):1,(d:1,e:1):1,f:1);'
    b = '(((a:1,c:1):1,b:1):1,(d:1,g:1):1,f:1);'
    ka, kb = kendallcolijn.KendallColijn(a), kendallcolijn.KendallColijn(b)
    assert np.allclose(ka.get_vector(0.25), oracle_mod.kc_vector(a, 0.25), rtol=1e-14)
    assert abs(ka.get_distance(kb, 0.5) - oracle_mod.kc_pair(a, b, 0.5)) < 1e-12                  # pruned to the 5 common leaves
    assert ka.get_distance(kb, 0.5, min_overlap=6) == 0
    k1, k2 = kendallcolijn.KendallColijn(T1), kendallcolijn.KendallColijn(T2)
    assert abs(k1.get_distance(k2) - np.sqrt(34.5)) < 1e-12
    sq = kendallcolijn.kendall_colijn_matrix([T1, T2, T1])
    assert sq.shape == (3, 3) and sq[0, 2] == 0.0 and abs(sq[0, 1] - np.sqrt(34.5)) < 1e-12
