"""CPU tests of the product's host side: C-ABI surface, split encoder, Tree / Collection mirrors, sharding plan.
No compute entry point is called here (there is no CPU fallback to call)."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import GOLDEN, ROOT


def test_every_declared_symbol_is_exported(lib_built):
    declared = set()
    inc = os.path.join(ROOT, 'include')
    for name in sorted(os.listdir(inc)):
        if name.endswith('.h'):
            header = re.sub(r'/\*.*?\*/', '', open(os.path.join(inc, name)).read(), flags=re.S)
            declared |= set(re.findall(r'\b(td_[a-z0-9_]+)\s*\(', header))
    assert {'td_synth_trees', 'td_check_errors', 'td_trim_cache'} <= declared
    assert {'td_encode_newick', 'td_upload', 'td_pairwise', 'td_pairwise_device', 'td_last_error'} <= declared
    lib = ctypes.CDLL(lib_built)
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, missing


def test_compute_fails_loudly_without_gpu(lib_built, cache_trees):
    from treecl_b200 import _lib
    if _lib.device_count() > 0:
        pytest.skip('a GPU is present')
    ts = _lib.TreeSet(cache_trees[1])
    with pytest.raises(_lib.NoDeviceError):
        ts.upload(0)
    with pytest.raises(_lib.TreedistError):
        ts.pairwise(('rf',))


def test_encoder_matches_oracle_encoding(lib_built, oracle_mod, cache_trees):
    from treecl_b200 import _lib
    names, trees = cache_trees
    ts = _lib.TreeSet(trees)
    inf = ts.info
    assert (inf.n_trees, inf.n_taxa, inf.words, inf.uniform, inf.rooted) == (15, 10, 1, 1, 0)
    assert ts.get(_lib.TD_GET_LABELS, None) == sorted('Sp%d' % i for i in range(1, 11))
    ns = ts.get(_lib.TD_GET_NSPLITS, np.int32)
    lens = ts.get(_lib.TD_GET_LENGTHS, np.float64).reshape(15, -1)
    leaf = ts.get(_lib.TD_GET_LEAFLEN, np.float64).reshape(15, -1)
    for i, t in enumerate(trees):
        s = oracle_mod.encode_summary(t)
        assert (s['n_leaves'], s['n_splits'], s['rooted']) == (10, ns[i], False)
        assert lens[i, :ns[i]].sum() == pytest.approx(s['sum_internal'], rel=1e-14)
        assert leaf[i].sum() == pytest.approx(s['sum_leaf'], rel=1e-14)


def _host_distances(ts, i, j):
    """rf / wrf / euc recomputed in numpy from the product's encoding (ids, lengths, leaf lengths)."""
    from treecl_b200 import _lib
    n = ts.n_trees
    ns = ts.get(_lib.TD_GET_NSPLITS, np.int32)
    ids = ts.get(_lib.TD_GET_IDS, np.uint32).reshape(n, -1)
    lens = ts.get(_lib.TD_GET_LENGTHS, np.float64).reshape(n, -1)
    leaf = ts.get(_lib.TD_GET_LEAFLEN, np.float64).reshape(n, -1)
    di = dict(zip(ids[i][:ns[i]].tolist(), lens[i])); dj = dict(zip(ids[j][:ns[j]].tolist(), lens[j]))
    keys = set(di) | set(dj)
    rf = sum(1 for k in keys if (k in di) != (k in dj))
    wrf = sum(abs(di.get(k, 0) - dj.get(k, 0)) for k in keys) + np.abs(leaf[i] - leaf[j]).sum()
    euc = np.sqrt(sum((di.get(k, 0) - dj.get(k, 0)) ** 2 for k in keys) + ((leaf[i] - leaf[j]) ** 2).sum())
    return rf, wrf, euc


@pytest.mark.parametrize('n_taxa', [5, 20, 64, 65, 130])
def test_dictionary_ids_reproduce_oracle_distances(lib_built, oracle_mod, n_taxa):
    from treecl_b200 import _lib, synth
    trees = synth.uniform_trees(10, n_taxa, seed=11) + synth.structured_trees(10, n_taxa, n_classes=2, seed=5)
    ts = _lib.TreeSet(trees, n_threads=3)
    assert ts.info.words == (n_taxa + 63) // 64 and ts.info.uniform == 1
    ids = ts.get(_lib.TD_GET_IDS, np.uint32).reshape(20, -1)
    ns = ts.get(_lib.TD_GET_NSPLITS, np.int32)
    for i in range(20):
        assert (np.diff(ids[i][:ns[i]].astype(np.int64)) > 0).all()          # sorted, unique
    for (i, j) in [(0, 1), (3, 17), (10, 11), (12, 19), (5, 5)]:
        rf, wrf, euc = _host_distances(ts, i, j)
        assert rf == oracle_mod.pair(trees[i], trees[j], 'rf')
        assert wrf == pytest.approx(oracle_mod.pair(trees[i], trees[j], 'wrf'), rel=1e-12, abs=1e-15)
        assert euc == pytest.approx(oracle_mod.pair(trees[i], trees[j], 'euc'), rel=1e-12, abs=1e-15)


def test_rooted_and_multifurcating_encoding(lib_built, oracle_mod):
    from treecl_b200 import _lib
    rooted = ['((A:1,B:2):0.5,((C:1,D:1):0.25,E:3):0.75);', '(((A:1,C:2):0.5,B:1):0.3,(D:1,E:1):0.4);']
    ts = _lib.TreeSet(rooted)
    assert ts.info.rooted == 1 and list(ts.get(_lib.TD_GET_NSPLITS, np.int32)) == [3, 3]      # n-2 clades
    assert _host_distances(ts, 0, 1)[0] == oracle_mod.pair(rooted[0], rooted[1], 'rf')
    poly = ['(A:1,B:1,(C:1,D:1,E:1):2,F:1);', '((A:1,B:1):1,(C:1,D:1):1,E:1,F:1);']
    ts = _lib.TreeSet(poly)
    assert list(ts.get(_lib.TD_GET_NSPLITS, np.int32)) == [1, 2]
    for k, m in enumerate(('rf', 'wrf', 'euc')):
        assert _host_distances(ts, 0, 1)[k] == pytest.approx(oracle_mod.pair(poly[0], poly[1], m), rel=1e-13)
    # quoted labels, comments, support values, scientific notation
    fancy = "(('t 1':1e-1,t_2:2E0)0.99:1[&c],[x]'it''s':1,(t4:1,t5:1)55:+3.5);"
    ts = _lib.TreeSet([fancy, "((t_2:1,t4:1):1,'t 1':1,('it''s':1,t5:1):1);"])
    assert ts.get(_lib.TD_GET_LABELS, None) == sorted(["t 1", "t_2", "it's", "t4", "t5"])
    assert ts.get(_lib.TD_GET_LENGTHS, np.float64).reshape(2, -1)[0].sum() == pytest.approx(4.5)


def test_non_uniform_sets_are_flagged(lib_built):
    from treecl_b200 import _lib
    ts = _lib.TreeSet(['((A:1,B:1):1,C:1,(D:1,E:1):1);', '((A:1,B:1):1,C:1,(D:1,F:1):1);'])
    assert ts.info.uniform == 0 and ts.info.n_taxa == 6
    lm = ts.get(_lib.TD_GET_LEAFMASK, np.uint64)
    assert [bin(int(x)).count('1') for x in lm] == [5, 5]


@pytest.mark.parametrize('bad', ['', 'A;', '((A,B),C', '(A,B))', '(A,,B);', "('A,B);", '(A:x,B:1);', '(A,A,B);',
                                 '(A:nan,B:1,C:1);', '(A:inf,B:1,C:1);', '(A:-inf,B:1,C:1);', '((C:1,D:1):1:2,A:1,B:1);', '(A:1:2,B:1,C:1);'])
def test_parse_errors_are_value_errors(lib_built, bad):
    from treecl_b200 import _lib
    with pytest.raises(ValueError):
        _lib.TreeSet(['(A,B,C);', bad])


def test_tree_mirror(lib_built, oracle_mod, fixture_trees):
    import treecl_b200
    t1 = treecl_b200.Tree(fixture_trees['class1_1'])
    t2 = treecl_b200.Tree(fixture_trees['class1_2'])
    assert len(t1) == 10 and t1.rooted is False and not (t1 ^ t2) and (t1 & t2) == t1.labels
    keep = {'Sp1', 'Sp2', 'Sp3', 'Sp4', 'Sp5', 'Sp6', 'Sp7'}
    p1 = t1.prune_to_subset(keep)
    assert p1.labels == keep and t1.labels != keep
    # the product's pruning (C++) and the oracle's (C, tree level) describe the same tree
    a = oracle_mod.encode_summary(p1.newick); b = oracle_mod.encode_summary(oracle_mod.prune_newick(t1.newick, keep))
    assert a['n_splits'] == b['n_splits'] == 4
    assert a['sum_internal'] == pytest.approx(b['sum_internal'], rel=1e-15)
    assert a['sum_leaf'] == pytest.approx(b['sum_leaf'], rel=1e-15)
    assert oracle_mod.pair(p1.newick, oracle_mod.prune_newick(t1.newick, keep), 'euc') == pytest.approx(0.0, abs=1e-15)
    assert treecl_b200.Tree('((A:1,B:1):1,(C:1,D:1):1);').rooted is True
    with pytest.raises(ValueError):
        treecl_b200.Tree('((A,B)').labels


def test_collection_loading_and_option_check(lib_built, cache_trees):
    import json
    import tempfile
    import treecl_b200
    names, trees = cache_trees
    c = treecl_b200.Collection(trees_dir=os.path.join(GOLDEN, 'trees'), show_progress=False)
    assert c.names == names and len(c) == 15
    assert c[0].tree.strip() == open(os.path.join(GOLDEN, 'trees', 'class1_1.nwk')).read().strip()
    with tempfile.TemporaryDirectory() as d:
        for n, t in zip(names, trees):
            open(os.path.join(d, n + '.phy'), 'w').write(' 0 0\n')
            json.dump({'ml_tree': t, 'nj_tree': None}, open(os.path.join(d, n + '.json'), 'w'))
        c2 = treecl_b200.Collection(input_dir=d, param_dir=d, file_format='phylip', show_progress=False)
        assert c2.names == names and c2.trees == trees
    with pytest.raises(treecl_b200.OptionError):
        c.get_inter_tree_distances('nope', show_progress=False)
    with pytest.raises(AttributeError):
        treecl_b200.Collection(records=[treecl_b200.Record('x')]).trees


def test_balanced_row_ranges_cover_the_triangle():
    from treecl_b200 import _lib, engine
    for n in (2, 3, 17, 200, 5000):
        for parts in (1, 2, 3, 8):
            r = engine.balanced_row_ranges(n, parts)
            assert r[0][0] == 0 and r[-1][1] == n and all(a[1] == b[0] for a, b in zip(r, r[1:]))
            counts = [_lib.condensed_offset(n, b) - _lib.condensed_offset(n, a) for a, b in r]
            assert sum(counts) == n * (n - 1) // 2
            if n >= 200:
                assert max(counts) - min(counts) <= 2 * n


def test_folded_row_ranges_tile_the_triangle_and_balance():
    from treecl_b200 import _lib, engine
    for n in (2, 17, 333, 2000, 5000, 20000, 100000):
        for parts in (1, 2, 3, 4, 8):
            fr = engine.folded_row_ranges(n, parts)
            assert len(fr) == parts and all(len(r) <= 2 for r in fr)
            rows = sorted(x for r in fr for x in r)
            assert rows[0][0] == 0 and rows[-1][1] == n and all(a[1] == b[0] for a, b in zip(rows, rows[1:]))
            counts = [sum(_lib.condensed_offset(n, b) - _lib.condensed_offset(n, a) for a, b in r) for r in fr]
            assert sum(counts) == n * (n - 1) // 2
            if n >= 2000:
                assert all(a % 64 == 0 for r in fr for a, _ in r)                 # whole tile rows
                assert max(counts) <= (1.03 if n >= 2500 * parts else 1.12) * (sum(counts) / parts), (n, parts, counts)
                if parts > 1:
                    assert all(len(r) == 2 for r in fr[:-1])                      # a long-row block and a short-row block each


def test_serialised_set_round_trip(lib_built):
    from treecl_b200 import _lib, synth
    for trees in (synth.structured_trees(120, 30, n_classes=3, seed=2), synth.uniform_trees(50, 70, seed=3),
                  ['((A:1,B:1):1,C:1,(D:1,E:1):1);', '((A:1,B:1):1,C:1,(D:1,F:1):1);']):
        ts = _lib.TreeSet(trees)
        blob = ts.serialize()
        back = _lib.TreeSet.deserialize(blob)
        assert back.n_trees == len(trees) and np.array_equal(back.serialize(), blob)
        for what, dt in ((_lib.TD_GET_MASKS, np.uint64), (_lib.TD_GET_LENGTHS, np.float64), (_lib.TD_GET_IDS, np.uint32),
                         (_lib.TD_GET_LEAFLEN, np.float64), (_lib.TD_GET_LEAFMASK, np.uint64), (_lib.TD_GET_ROOTED, np.int32)):
            assert np.array_equal(ts.get(what, dt), back.get(what, dt))
        assert ts.get(_lib.TD_GET_LABELS, None) == back.get(_lib.TD_GET_LABELS, None)
        a, b = ts.info, back.info
        assert (a.uniform, a.dict_size, a.dict_shared, a.max_shared, a.heavy_rows, a.n_postings) == (b.uniform, b.dict_size, b.dict_shared, b.max_shared, b.heavy_rows, b.n_postings)
        for bad in (blob[:-8], blob[:64], np.concatenate([blob, blob[:8]]), np.zeros(256, dtype=np.uint8)):
            with pytest.raises(ValueError):
                _lib.TreeSet.deserialize(bad)


def test_synthetic_sets_are_deterministic(lib_built):
    from treecl_b200 import _lib, synth
    a = synth.uniform_trees(6, 20, seed=3); b = synth.uniform_trees(6, 20, seed=3)
    assert a == b and a != synth.uniform_trees(6, 20, seed=4)
    ts = _lib.TreeSet(a)
    assert list(ts.get(_lib.TD_GET_NSPLITS, np.int32)) == [17] * 6      # binary, trifurcating root: n-3
    s = _lib.TreeSet(synth.structured_trees(64, 30, n_classes=4, lam=0.5, seed=1))
    assert s.info.dict_size < 64 * 27 // 4            # shared structure -> small dictionary


def test_native_generator(lib_built, oracle_mod):
    """csrc/synth.cpp: deterministic, independent of the set size and of the thread count, binary trees of the requested shape."""
    from treecl_b200 import _lib, synth
    a = synth.native_trees(200, 30, 'uniform', seed=7, n_threads=1)
    b = synth.native_trees(300, 30, 'uniform', seed=7, n_threads=4)
    assert a == b[:200] and a != synth.native_trees(200, 30, 'uniform', seed=8)
    ts = _lib.TreeSet(a)
    assert ts.info.uniform == 1 and list(ts.get(_lib.TD_GET_NSPLITS, np.int32)) == [27] * 200 and ts.info.n_taxa == 30
    assert ts.get(_lib.TD_GET_LABELS, None) == synth.taxon_labels(30)
    g = synth.NativeTrees(256, 40, 'structured', seed=2, n_classes=4, class_nni=6, lam=1.0)
    s = g.treeset()
    assert s.n_trees == 256 and s.info.uniform == 1 and s.info.dict_size < 256 * 37 // 4      # shared structure -> small dictionary
    assert g.strings() == synth.native_trees(256, 40, 'structured', seed=2, n_classes=4, class_nni=6, lam=1.0)
    # mean RF of independent uniform trees is close to its maximum; the oracle reads the generator's Newick
    rf = oracle_mod.matrix(a[:40], 'rf')
    assert rf.mean() > 2 * 27 * 0.9
    g.close()


def test_oracle_matrix_rows_agree_with_the_condensed_matrix(oracle_mod):
    from scipy.spatial.distance import squareform
    from treecl_b200 import synth
    trees = synth.structured_trees(60, 16, n_classes=3, class_nni=3, lam=1.0, seed=4)
    rows = [0, 7, 59, 31]
    for m in ('rf', 'wrf', 'euc', 'geo'):
        sq = squareform(oracle_mod.matrix(trees, m, nthreads=2))
        assert np.array_equal(oracle_mod.matrix_rows(trees, m, rows, nthreads=3), sq[rows]), m


def test_marshalling_paths_build_the_same_set(lib_built, monkeypatch):
    """The optional CPython helper (zero-copy pointer table) and the joined-buffer path must encode identically."""
    from treecl_b200 import _lib, synth
    trees = synth.uniform_trees(40, 17, seed=5) + synth.structured_trees(40, 17, n_classes=3, class_nni=3, lam=1.0, seed=6)

    def snapshot():
        ts = _lib.TreeSet(trees)
        inf = ts.info
        return ((inf.n_trees, inf.n_taxa, inf.max_splits, inf.max_shared, inf.dict_size, inf.dict_shared, inf.total_splits, inf.n_postings),
                ts.get(_lib.TD_GET_IDS, np.uint32).copy(), ts.get(_lib.TD_GET_LENGTHS, np.float64).copy(),
                ts.get(_lib.TD_GET_MASKS, np.uint64).copy(), ts.get(_lib.TD_GET_LEAFLEN, np.float64).copy())

    monkeypatch.setattr(_lib, '_glue', False)
    monkeypatch.delenv('TREEDIST_NO_PYGLUE', raising=False)
    a = snapshot()
    used_glue = _lib._pyglue() is not None
    monkeypatch.setattr(_lib, '_glue', False)
    monkeypatch.setenv('TREEDIST_NO_PYGLUE', '1')
    b = snapshot()
    assert _lib._pyglue() is None
    assert a[0] == b[0]
    for x, y in zip(a[1:], b[1:]):
        assert np.array_equal(x, y)
    # bytes input goes through either path as well
    monkeypatch.setattr(_lib, '_glue', False)
    monkeypatch.delenv('TREEDIST_NO_PYGLUE', raising=False)
    ts = _lib.TreeSet([t.encode() for t in trees])
    assert np.array_equal(ts.get(_lib.TD_GET_IDS, np.uint32), a[1])
    assert used_glue or True          # the helper is optional; when it was built it has just been exercised


@pytest.mark.parametrize('no_glue', ['', '1'])
def test_embedded_nul_is_refused(lib_built, monkeypatch, no_glue):
    from treecl_b200 import _lib
    monkeypatch.setattr(_lib, '_glue', False)
    if no_glue:
        monkeypatch.setenv('TREEDIST_NO_PYGLUE', no_glue)
    else:
        monkeypatch.delenv('TREEDIST_NO_PYGLUE', raising=False)
    with pytest.raises(ValueError):
        _lib.TreeSet(['((a:1,b:1):1,(c:1,d:1):1);', '((a:1,c:1):1,\0(b:1,d:1):1);'])


def test_postings_cover_exactly_the_shared_entries(lib_built):
    """n_postings = number of (tree, split) entries whose split occurs in at least two trees (what the events kernel enumerates)."""
    from treecl_b200 import _lib, synth
    trees = synth.structured_trees(60, 24, n_classes=4, class_nni=3, lam=2.0, seed=11)
    ts = _lib.TreeSet(trees)
    inf = ts.info
    ids = ts.get(_lib.TD_GET_IDS, np.uint32)
    ids = ids[ids != 0xFFFFFFFF]
    uniq, counts = np.unique(ids, return_counts=True)
    assert inf.dict_size == uniq.size
    assert inf.dict_shared == int((counts >= 2).sum())
    assert inf.n_postings == int(counts[counts >= 2].sum())


def test_heavy_rows_follow_the_sharing_level(lib_built):
    """The encoder's choice for the hybrid tile kernel: no heavy rows for independent random trees (nothing is shared by many trees),
    a few rows that cover almost all common splits for a clustered set."""
    from treecl_b200 import _lib, synth
    sparse = _lib.TreeSet(synth.uniform_trees(1500, 40, seed=3)).info
    assert sparse.heavy_rows == 0 and sparse.heavy_common == 0.0 and 0.0 < sparse.mean_common < 1.0
    dense = _lib.TreeSet(synth.structured_trees(1500, 40, n_classes=8, class_nni=6, lam=2.0, seed=4)).info
    assert dense.mean_common > 10.0
    assert 0 < dense.heavy_rows <= dense.dict_shared
    assert 0.9 * dense.mean_common <= dense.heavy_common <= dense.mean_common + 1e-9


def test_encoding_does_not_depend_on_the_thread_count(lib_built):
    """The encoder's workers write the big arrays of a set in full, padding included (raw_vector: nothing is zero-filled first), and the
    dictionary is built from hash buckets filled in parallel: the serialised set must be the same bytes for any number of threads, for
    uniform, clustered and non-uniform sets and for Kendall-Colijn vectors."""
    import treecl_b200 as td
    from treecl_b200 import _lib, synth
    uniform = synth.uniform_trees(300, 40, seed=3)
    clustered = synth.structured_trees(300, 70, n_classes=4, class_nni=6, lam=2.0, seed=4)
    labels = synth.taxon_labels(40)
    rng = np.random.default_rng(5)
    odd = [td.Tree(t).prune_to_subset({labels[i] for i in rng.choice(40, size=int(rng.integers(20, 41)), replace=False)}).newick if k % 3 else t
           for k, t in enumerate(uniform[:120])]
    for trees in (uniform, clustered, odd):
        blobs = [_lib.TreeSet(trees, n_threads=nt).serialize() for nt in (1, 2, 5, 8)]
        for b in blobs[1:]:
            assert b.size == blobs[0].size and np.array_equal(b, blobs[0])
        back = _lib.TreeSet.deserialize(blobs[0]).serialize()
        assert np.array_equal(back, blobs[0])
    kc = [_lib.TreeSet.kendall_colijn(uniform[:80], lbda=0.3, n_threads=nt).serialize() for nt in (1, 6)]
    assert np.array_equal(kc[0], kc[1])
