"""Pin the CPU oracle against every golden vector the reference's own tests hold for the distance path
(reference tests/test.py:234-294, :320-326, :437-467 and tests/data/cache/geo_dm.csv)."""
import os

import numpy as np
import pandas as pd
import pytest
from scipy.spatial.distance import squareform

from conftest import GOLDEN


@pytest.mark.parametrize('metric', ['geo', 'euc', 'wrf', 'rf'])
@pytest.mark.parametrize('norm', [False, True])
def test_scalar_goldens(oracle_mod, fixture_trees, golden_scalars, metric, norm):
    a, b = (fixture_trees[n] for n in golden_scalars['pair'])
    got = oracle_mod.pair(a, b, metric, norm)
    want = golden_scalars[metric + ('_norm' if norm else '')]
    assert got == pytest.approx(want, rel=1e-12, abs=0)          # reference asserts 7 decimal places


def test_rf_is_exact(oracle_mod, fixture_trees, golden_scalars):
    a, b = (fixture_trees[n] for n in golden_scalars['pair'])
    assert oracle_mod.pair(a, b, 'rf') == 2.0
    assert oracle_mod.pair(a, b, 'rf', True) == 2.0 / 14.0


def test_partial_overlap(oracle_mod, fixture_trees, golden_scalars):
    po = golden_scalars['partial_overlap']
    a, b = (fixture_trees[n] for n in golden_scalars['pair'])
    pa = oracle_mod.prune_newick(a, po['keep_0'])
    pb = oracle_mod.prune_newick(b, po['keep_1'])
    assert oracle_mod.pair(pa, pb, 'geo') == pytest.approx(po['geo'], rel=1e-12)
    assert oracle_mod.pair(pa, pb, 'geo', min_overlap=5, overlap_fail_value=-1) == po['min_overlap_5_fail_value']
    # pruning first or letting the pair function prune gives the same tree pair
    assert oracle_mod.pair(pa, pb, 'geo') == pytest.approx(
        oracle_mod.pair(oracle_mod.prune_newick(a, ['Sp4', 'Sp5', 'Sp6', 'Sp7']),
                        oracle_mod.prune_newick(b, ['Sp4', 'Sp5', 'Sp6', 'Sp7']), 'geo'), rel=1e-13)


def test_geodesic_matrix(oracle_mod, cache_trees, golden_scalars):
    names, trees = cache_trees
    cond = oracle_mod.matrix(trees, 'geo')
    sq = squareform(cond)
    assert sq.sum() == pytest.approx(golden_scalars['geo_matrix_checksum'], abs=5e-8)      # reference: 7 d.p.
    df = pd.read_csv(os.path.join(GOLDEN, 'geo_dm.csv'), index_col=0)
    assert list(df.index) == names
    iu = np.triu_indices(len(names), 1)
    rel = np.abs(sq[iu] - df.values[iu]) / df.values[iu]
    assert rel.max() < 1e-9           # csv was written by an older build: agrees to ~1e-11
    # every jobhandler of the reference must give the same matrix: threads / per-pair re-parse change nothing
    assert np.array_equal(cond, oracle_mod.matrix(trees, 'geo', nthreads=3))
    assert np.array_equal(cond, oracle_mod.matrix(trees, 'geo', nthreads=2, reparse_per_pair=True))


def test_matrix_equals_scalar(oracle_mod, cache_trees):
    names, trees = cache_trees
    for metric in ('rf', 'wrf', 'euc'):
        cond = oracle_mod.matrix(trees, metric)
        assert cond[0] == oracle_mod.pair(trees[0], trees[1], metric)       # tests/test.py:275-278
        assert cond[len(trees) - 2] == oracle_mod.pair(trees[0], trees[-1], metric)


def test_parse_errors(oracle_mod):
    for bad in ['', 'A;', '((A,B),C', '(A,B))', '(A,,B);', "('A,B);", '(A:x,B:1);']:
        with pytest.raises(ValueError):
            oracle_mod.pair(bad, '(A,B,C);', 'rf')


def test_metric_axioms_small(oracle_mod):
    from treecl_b200 import synth
    trees = synth.uniform_trees(12, 9, seed=7)
    for metric in ('rf', 'wrf', 'euc', 'geo'):
        sq = squareform(oracle_mod.matrix(trees, metric))
        assert (sq >= 0).all()
        assert oracle_mod.pair(trees[3], trees[3], metric) == 0.0
        # triangle inequality
        for i in range(12):
            for j in range(12):
                assert (sq[i, j] <= sq[i, :] + sq[:, j] + 1e-9).all()
    # geodesic is squeezed between the Euclidean lower bound and the cone path
    geo = oracle_mod.matrix(trees, 'geo'); euc = oracle_mod.matrix(trees, 'euc')
    assert (geo >= euc - 1e-12).all()


def test_mds_restatement_matches_the_reference_functions():
    """oracle/mds.py against tests/golden/mds_golden.json = outputs of the reference's own double_centre / eigen /
    _embedding_classical_mds (treeCl/distance_matrix.py:58-74,285-298,301-315), produced by tests/golden/make_mds_golden.py."""
    import json
    import oracle
    with open(os.path.join(GOLDEN, 'mds_golden.json')) as fh:
        cases = json.load(fh)
    assert set(cases) == {'geo_dm_15', 'clustered_64'}
    for name, c in cases.items():
        m, dims = np.array(c['matrix']), c['dimensions']
        dbc = oracle.mds.double_centre(m)
        assert np.allclose(dbc, np.array(c['double_centre']), rtol=1e-13, atol=1e-13), name
        vals, vecs = oracle.mds.eigen(dbc)
        want_vals = np.array(c['eigenvalues'])
        assert np.allclose(vals, want_vals, rtol=1e-10, atol=1e-10 * abs(want_vals).max()), name
        coords, _ = oracle.mds.classical_mds(m, dims)
        want = np.array(c['coords'])
        # eigenvectors are defined up to sign: compare column by column after aligning the signs
        for j in range(dims):
            s = np.sign(np.dot(coords[:, j], want[:, j]))
            assert np.allclose(s * coords[:, j], want[:, j], rtol=1e-8, atol=1e-9 * abs(want).max()), (name, j)


def test_spectral_restatement_matches_the_reference_functions():
    """oracle/mds.py's Spectral chain against the outputs of the reference's own kmask / kscale / affinity / laplace / eigen /
    normalise_rows, called in the order of treeCl/clustering.py:171-228,288-303 (tests/golden/make_mds_golden.py)."""
    import json
    import oracle
    with open(os.path.join(GOLDEN, 'mds_golden.json')) as fh:
        cases = json.load(fh)
    n_checked = 0
    for name, c in cases.items():
        m, dims = np.array(c['matrix']), c['dimensions']
        for sp in c['spectral']:
            assert np.allclose(oracle.mds.kdists(m, sp['kdists_k']).ravel(), sp['kdists'], rtol=0, atol=0)
            aff = oracle.mds.spectral_affinity(m, sp['prune_k'], sp['logic'], sp['scale_k'])
            assert np.allclose(aff, np.array(sp['affinity']), rtol=1e-14, atol=1e-300), (name, sp['prune_k'])
            coords, vals = oracle.mds.spectral_embedding(m, dims, sp['prune_k'], sp['logic'], sp['scale_k'])
            assert np.allclose(vals, sp['eigenvalues'], rtol=0, atol=1e-10)
            want = np.array(sp['coords'])
            # (nearly) equal leading eigenvalues leave a rotation free: compare the inner products of the unit rows
            assert np.allclose(coords.dot(coords.T), want.dot(want.T), rtol=0, atol=1e-7), (name, sp['prune_k'], sp['scale_k'])
            n_checked += 1
    assert n_checked >= 8
