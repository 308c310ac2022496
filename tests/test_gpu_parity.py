"""Parity of the CUDA path (through the C ABI / the treeCl-shaped Python API) against the CPU oracle and the
reference's golden vectors.  Tolerances: RF bit-exact; wRF / Euclidean / geodesic <= 1e-9 relative (fp64; the order
of summation differs from the oracle's), as BASELINE.json's north_star states."""
import os

import numpy as np
import pandas as pd
import pytest
from scipy.spatial.distance import squareform

from conftest import GOLDEN, ROOT

pytestmark = pytest.mark.gpu
RTOL = 1e-9


def assert_close(got, want, rtol=RTOL, atol=1e-12):
    got, want = np.asarray(got), np.asarray(want)
    err = np.abs(got - want)
    bad = err > rtol * np.abs(want) + atol
    assert not bad.any(), 'max rel err {:.3e} at {}'.format((err / np.maximum(np.abs(want), 1e-300)).max(), np.argmax(err))


def condensed_index(n, i, j):
    """scipy-condensed position of the pair (i, j), i != j (vectorised)."""
    i, j = np.asarray(i, dtype=np.int64), np.asarray(j, dtype=np.int64)
    a, b = np.minimum(i, j), np.maximum(i, j)
    return a * n - a * (a + 1) // 2 + (b - a - 1)


def square_rows(cond, n, rows):
    """Rows `rows` of the square matrix behind the condensed vector `cond` (0 on the diagonal), shape (len(rows), n)."""
    out = np.zeros((len(rows), n), dtype=np.float64)
    cols = np.arange(n)
    for k, r in enumerate(rows):
        other = cols[cols != r]
        out[k, other] = cond[condensed_index(n, r, other)]
    return out


NTHREADS = os.cpu_count() or 4


@pytest.fixture(scope='module')
def td():
    import treecl_b200
    from treecl_b200 import _lib
    assert _lib.device_count() > 0, 'these tests need a GPU; the product has no CPU fallback'
    return treecl_b200


# ---- the reference's own tests, ported (reference tests/test.py:234-294) -------------------------------------------
class TestTreeDistanceGoldens:
    @pytest.fixture(autouse=True)
    def setup(self, td, fixture_trees):
        self.td = td
        self.c = td.Collection(trees_dir=os.path.join(GOLDEN, 'trees'), show_progress=False)
        self.tree1 = td.Tree(self.c[0].tree)
        self.tree2 = td.Tree(self.c[1].tree)

    @pytest.mark.parametrize('metric', ['geo', 'euc', 'wrf', 'rf'])
    @pytest.mark.parametrize('norm', [False, True])
    def test_scalar(self, golden_scalars, metric, norm):
        fn = getattr(self.td.treedist, metric + 'dist')
        want = golden_scalars[metric + ('_norm' if norm else '')]
        got = fn(self.tree1, self.tree2, norm)
        if metric == 'rf':
            assert got == want
        else:
            assert got == pytest.approx(want, rel=RTOL)

    def test_intertree(self):
        dm = self.c.get_inter_tree_distances('rf', show_progress=False)
        assert dm.values[0, 1] == self.td.treedist.rfdist(self.tree1, self.tree2, False)
        assert dm.get_names() == self.c.names and dm.values.shape == (15, 15)
        # the matrix path agrees with the scalar path for EVERY entry and every metric, not only [0, 1]
        trees = [self.td.Tree(r.tree) for r in self.c]
        for metric in ('rf', 'wrf', 'euc', 'geo'):
            vals = self.c.get_inter_tree_distances(metric, show_progress=False).values
            fn = getattr(self.td.treedist, metric + 'dist')
            for i in range(15):
                for j in range(i + 1, 15):
                    want = fn(trees[i], trees[j], False)
                    assert vals[i, j] == want if metric == 'rf' else vals[i, j] == pytest.approx(want, rel=1e-12), (metric, i, j)
        assert (dm.values == dm.values.T).all() and (np.diag(dm.values) == 0).all()


class TestPartialOverlapGoldens:
    """reference tests/test.py:280-294: trees pruned to different leaf sets -> distance on the intersection."""

    @pytest.fixture(autouse=True)
    def setup(self, td, golden_scalars):
        self.c = td.Collection(trees_dir=os.path.join(GOLDEN, 'trees'), show_progress=False)
        po = golden_scalars['partial_overlap']
        self.c[0].parameters.ml_tree = td.Tree(self.c[0].parameters.ml_tree).prune_to_subset(set(po['keep_0'])).newick
        self.c[1].parameters.ml_tree = td.Tree(self.c[1].parameters.ml_tree).prune_to_subset(set(po['keep_1'])).newick
        self.po = po

    def test_intertree_partial_overlap(self):
        dm = self.c.get_inter_tree_distances('geo', show_progress=False)
        assert dm.values[0, 1] == pytest.approx(self.po['geo'], rel=RTOL)

    def test_intertree_partial_overlap_value_replacement(self):
        dm = self.c.get_inter_tree_distances('geo', min_overlap=5, overlap_fail_value=-1, show_progress=False)
        assert dm.values[0, 1] == -1


class TestGeodesicMatrixGolden:
    """reference tests/test.py:320-326 and ParallelTests :437-467 (every jobhandler gives the same checksum)."""

    def test_matrix(self, td, cache_trees, golden_scalars):
        names, trees = cache_trees
        c = td.Collection(records=[td.Record(n, ml_tree=t) for n, t in zip(names, trees)])
        handlers = [None, td.parutils.SequentialJobHandler(), td.parutils.ProcesspoolJobHandler(2),
                    td.parutils.ThreadpoolJobHandler(2)]
        first = None
        for h in handlers:
            for bs in (1, 5):
                kw = dict(show_progress=False, batchsize=bs)
                if h is not None:
                    kw['jobhandler'] = h
                dm = c.get_inter_tree_distances('geo', **kw)
                assert dm.df.values.sum() == pytest.approx(golden_scalars['geo_matrix_checksum'], abs=5e-8)
                if first is None:
                    first = dm.values.copy()
                assert np.array_equal(first, dm.values)
        want = pd.read_csv(os.path.join(GOLDEN, 'geo_dm.csv'), index_col=0).values
        iu = np.triu_indices(15, 1)
        assert_close(first[iu], want[iu])

    def test_matrix_functions_return_ndarray(self, td, cache_trees, oracle_mod):
        names, trees = cache_trees
        objs = [td.Tree(t) for t in trees]
        for metric in ('rf', 'wrf', 'euc', 'geo'):
            sq = getattr(td.treedist, metric + 'dist_matrix')(objs, False, show_progress=False)
            assert isinstance(sq, np.ndarray) and sq.shape == (15, 15)
            want = squareform(oracle_mod.matrix(trees, metric))
            if metric == 'rf':
                assert np.array_equal(sq, want)
            else:
                assert_close(sq, want)


# ---- randomised parity against the oracle ---------------------------------------------------------------------------
CASES = [
    # (n_trees, n_taxa, kind)
    (2, 4, 'uniform'), (3, 5, 'uniform'), (65, 8, 'uniform'), (200, 20, 'uniform'), (130, 50, 'uniform'),
    (97, 64, 'uniform'), (70, 65, 'uniform'), (66, 100, 'uniform'), (40, 200, 'uniform'),
    (150, 30, 'structured'), (100, 100, 'structured'),
    (333, 12, 'uniform'),          # five tile rows, N not a multiple of 16 / 32 / 64, many common splits per pair
    # more than 255 shared splits per tree (byte counters of the tile kernel: rf from the incidence contraction) and lists too long
    # for the shared-memory merge kernel (big-tree kernel as the last resort): no tree size is refused
    (40, 300, 'structured'), (24, 300, 'uniform'), (20, 520, 'structured'),
]


def make(kind, n_trees, n_taxa, seed):
    from treecl_b200 import synth
    if kind == 'uniform':
        return synth.uniform_trees(n_trees, n_taxa, seed=seed)
    return synth.structured_trees(n_trees, n_taxa, n_classes=3, class_nni=4, lam=1.5, seed=seed)


@pytest.fixture
def pair_kernel(request, monkeypatch):
    """Force one of the rf/wrf/euc kernels (split = postings pipeline + fp64 tensor-core tile kernel, join / join_tile = fused hash
    join, merge = shared-memory merge-join) or let the library choose."""
    monkeypatch.delenv('TREEDIST_NO_HYBRID', raising=False)
    if request.param == 'split_events_only':          # postings pipeline without heavy rows: every common split is an event
        monkeypatch.setenv('TREEDIST_PAIR_KERNEL', 'split')
        monkeypatch.setenv('TREEDIST_NO_HYBRID', '1')
    elif request.param != 'auto':
        monkeypatch.setenv('TREEDIST_PAIR_KERNEL', request.param)
    else:
        monkeypatch.delenv('TREEDIST_PAIR_KERNEL', raising=False)
    return request.param


@pytest.mark.parametrize('n_trees,n_taxa,kind', CASES)
@pytest.mark.parametrize('norm', [False, True])
@pytest.mark.parametrize('pair_kernel', ['auto', 'merge', 'join', 'join_tile', 'split', 'split_events_only', 'extended'], indirect=True)
def test_rf_wrf_euc_vs_oracle(td, oracle_mod, n_trees, n_taxa, kind, norm, pair_kernel):
    from treecl_b200 import _lib
    trees = make(kind, n_trees, n_taxa, seed=100 + n_taxa)
    ts = _lib.TreeSet(trees)
    got = ts.pairwise(('rf', 'wrf', 'euc'), normalise=norm)
    for metric in ('rf', 'wrf', 'euc'):
        want = oracle_mod.matrix(trees, metric, normalise=norm, nthreads=4)
        if metric == 'rf':
            assert np.array_equal(got[metric], want), metric          # bit exact (also when normalised)
        else:
            assert_close(got[metric], want)
    # single-metric launches use other template instantiations (same answers, bit for bit) - or, when the library chooses, even
    # another kernel (rf alone -> incidence contraction, euc alone -> split-extended tile kernel on densely sharing sets)
    for metric in ('rf', 'wrf', 'euc'):
        alone = ts.pairwise((metric,), normalise=norm)[metric]
        if metric == 'rf' or pair_kernel not in ('auto', 'extended', 'split'):     # (hybrid: rf from another kernel, wrf / euc identical masks differ in chunking only)
            assert np.array_equal(alone, got[metric]), metric
        else:
            assert_close(alone, got[metric])
    both = ts.pairwise(('rf', 'euc'), normalise=norm)
    assert np.array_equal(both['rf'], got['rf'])
    assert_close(both['euc'], got['euc'])


@pytest.mark.parametrize('n_trees,n_taxa,kind', [(2, 4, 'uniform'), (3, 5, 'uniform'), (120, 10, 'uniform'),
                                                 (150, 20, 'uniform'), (200, 30, 'uniform'), (60, 35, 'uniform'),
                                                 (50, 50, 'uniform'), (30, 67, 'uniform'), (24, 100, 'uniform'), (16, 131, 'uniform'),
                                                 (150, 30, 'structured'), (60, 60, 'structured'), (40, 120, 'structured'),
                                                 # more than 128 internal edges per tree: the big-tree kernel (no size limit, as in the reference)
                                                 (16, 132, 'uniform'), (16, 200, 'uniform'), (8, 300, 'uniform'), (24, 200, 'structured'), (6, 520, 'structured')])
@pytest.mark.parametrize('norm', [False, True])
def test_geodesic_vs_oracle(td, oracle_mod, n_trees, n_taxa, kind, norm):
    from treecl_b200 import _lib
    trees = make(kind, n_trees, n_taxa, seed=300 + n_taxa)
    ts = _lib.TreeSet(trees)
    got = ts.pairwise(('geo',), normalise=norm)['geo']
    want = oracle_mod.matrix(trees, 'geo', normalise=norm, nthreads=8)
    assert_close(got, want)
    st = ts.geo_stats()
    assert st['pairs'] == n_trees * (n_trees - 1) // 2 and st['ratios'] >= 0


@pytest.mark.parametrize('n_trees,n_taxa', [(300, 30), (200, 50), (60, 100)])
def test_flow_matrix_in_shared_or_global_memory(td, monkeypatch, n_trees, n_taxa):
    """The GTP kernels keep the per-warp flow matrix in shared memory or in a global scratch slice, whichever keeps more warps resident
    (GeoPlan); both forms run the same arithmetic: identical bytes, uniform and non-uniform sets."""
    from treecl_b200 import _lib, synth
    trees = synth.uniform_trees(n_trees, n_taxa, seed=900 + n_taxa)
    labels = synth.taxon_labels(n_taxa)
    rng = np.random.default_rng(n_taxa)
    odd = [td.Tree(t).prune_to_subset({labels[i] for i in rng.choice(n_taxa, size=n_taxa - 3, replace=False)}).newick if k % 3 else t
           for k, t in enumerate(trees[:40])]
    res = {}
    for form in ('smem', 'global'):
        monkeypatch.setenv('TREEDIST_GEO_FLOW', form)
        res[form] = (_lib.TreeSet(trees).pairwise(('geo',))['geo'], _lib.TreeSet(odd).pairwise(('geo', 'euc')))
    assert np.array_equal(res['smem'][0], res['global'][0])
    assert np.array_equal(res['smem'][1]['geo'], res['global'][1]['geo']) and np.array_equal(res['smem'][1]['euc'], res['global'][1]['euc'])


@pytest.mark.parametrize('n_taxa,n_trees', [(12, 60), (30, 80), (40, 50), (66, 30), (110, 16), (160, 14), (260, 10)])
@pytest.mark.parametrize('norm', [False, True])
def test_different_leaf_sets_vs_oracle(td, oracle_mod, n_taxa, n_trees, norm):
    """Pairs with different leaf sets (restrict-then-join kernel) against the oracle's tree-level pruning.
    UNPINNED by the reference for rf/wrf/euc (only the geodesic has a golden)."""
    from treecl_b200 import _lib, synth
    rng = np.random.default_rng(n_taxa)
    full = synth.uniform_trees(n_trees, n_taxa, seed=77 + n_taxa)
    labels = synth.taxon_labels(n_taxa)
    trees = []
    for k, t in enumerate(full):
        if k % 4 == 0:
            trees.append(t)                                   # some complete trees
        else:
            keep = rng.choice(n_taxa, size=int(rng.integers(max(4, n_taxa // 2), n_taxa + 1)), replace=False)
            trees.append(td.Tree(t).prune_to_subset({labels[i] for i in keep}).newick)
    ts = _lib.TreeSet(trees)
    assert ts.info.uniform == 0
    mo = max(4, n_taxa // 2)                                 # some pairs fall under min_overlap
    got = ts.pairwise(('rf', 'wrf', 'euc', 'geo'), normalise=norm, min_overlap=mo, overlap_fail_value=-7.5)
    n_fail = 0
    for metric in ('rf', 'wrf', 'euc', 'geo'):
        want = oracle_mod.matrix(trees, metric, normalise=norm, min_overlap=mo, overlap_fail_value=-7.5, nthreads=8)
        n_fail += int((want == -7.5).sum())
        if metric == 'rf':
            assert np.array_equal(got[metric], want)
        else:
            assert_close(got[metric], want)
    # sharded rows give the same bytes
    a = ts.pairwise(('geo',), normalise=norm, min_overlap=mo, overlap_fail_value=-7.5, row_begin=3, row_end=n_trees - 2)['geo']
    p0, p1 = _lib.condensed_offset(n_trees, 3), _lib.condensed_offset(n_trees, n_trees - 2)
    assert np.array_equal(a, got['geo'][p0:p1])


def test_leaf_set_groups(td, oracle_mod, monkeypatch):
    """treeCl/treedist.py:63-68 prunes only the pairs whose leaf sets differ.  A set in which a few trees miss taxa: the trees that share a
    leaf set form groups served by the uniform kernels (own dictionary), only the pairs across groups take the restrict-then-join kernel.
    Every entry against the oracle, all four metrics; the grouped result equals the ungrouped one (TREEDIST_NO_GROUPS) to rounding."""
    from treecl_b200 import _lib, synth
    n, taxa = 700, 20
    rng = np.random.default_rng(5)
    full = synth.structured_trees(n, taxa, n_classes=4, class_nni=5, lam=1.0, seed=15)
    labels = synth.taxon_labels(taxa)
    trees = list(full)
    miss3 = {labels[i] for i in range(taxa) if i not in (2, 7, 11)}
    for k in rng.choice(n, size=120, replace=False)[:60]:
        trees[k] = td.Tree(full[k]).prune_to_subset(miss3).newick                  # a second group: 60 trees without three taxa
    for k in rng.choice(n, size=12, replace=False):
        keep = rng.choice(taxa, size=int(rng.integers(10, taxa)), replace=False)
        trees[k] = td.Tree(full[k]).prune_to_subset({labels[i] for i in keep}).newick   # odd trees, each with a leaf set of its own
    ts = _lib.TreeSet(trees)
    assert ts.info.uniform == 0
    metrics = ('rf', 'wrf', 'euc', 'geo')
    got = ts.pairwise(metrics, normalise=True, min_overlap=11, overlap_fail_value=-2.0)
    assert ts.info.n_groups == 2
    for m in metrics:
        want = oracle_mod.matrix(trees, m, normalise=True, min_overlap=11, overlap_fail_value=-2.0, nthreads=NTHREADS)
        assert (want == -2.0).any()
        assert np.array_equal(got[m], want) if m == 'rf' else assert_close(got[m], want) is None
    # row ranges cut through the groups: same bytes as the one-shot call for the same metrics
    whole = ts.pairwise(('wrf', 'geo'), normalise=True, min_overlap=11, overlap_fail_value=-2.0)
    assert_close(whole['wrf'], got['wrf'], rtol=1e-12)
    for a, b in ((0, 250), (250, 251), (251, n)):
        part = ts.pairwise(('wrf', 'geo'), normalise=True, min_overlap=11, overlap_fail_value=-2.0, row_begin=a, row_end=b)
        p0, p1 = _lib.condensed_offset(n, a), _lib.condensed_offset(n, b)
        assert np.array_equal(part['wrf'], whole['wrf'][p0:p1]) and np.array_equal(part['geo'], whole['geo'][p0:p1])
    # the pairs across groups normally come from lists; without them the kernel skips same-group pairs chunk by chunk: same bytes
    monkeypatch.setenv('TREEDIST_NO_CROSS_LISTS', '1')
    skipping = _lib.TreeSet(trees)
    sk = skipping.pairwise(metrics, normalise=True, min_overlap=11, overlap_fail_value=-2.0)
    assert skipping.info.n_groups == 2
    for m in metrics:
        assert np.array_equal(sk[m], got[m])
    monkeypatch.delenv('TREEDIST_NO_CROSS_LISTS')
    # ungrouped: every pair through the restrict-then-join kernel
    monkeypatch.setenv('TREEDIST_NO_GROUPS', '1')
    plain = _lib.TreeSet(trees)
    ref = plain.pairwise(metrics, normalise=True, min_overlap=11, overlap_fail_value=-2.0)
    assert plain.info.n_groups == 0
    for m in metrics:
        assert np.array_equal(got[m], ref[m]) if m == 'rf' else assert_close(got[m], ref[m], rtol=1e-12) is None


def test_mixed_rooting_is_refused(td):
    from treecl_b200 import _lib
    ts = _lib.TreeSet(['((A:1,B:1):1,(C:1,D:1):1);', '((A:1,B:1):1,C:1,D:1);', '((A:1,C:1):1,B:1,D:1);'])
    with pytest.raises(ValueError):
        ts.pairwise(('rf',))


def test_async_entry_points_surface_the_error_flag(td):
    """td_pairwise_device cannot return what a kernel finds: the mixed-rooting flag is read with td_check_errors (and cleared)."""
    import torch
    from treecl_b200 import _lib
    ts = _lib.TreeSet(['((A:1,B:1):1,(C:1,D:1):1);', '((A:1,B:1):1,C:1,D:1);', '((A:1,C:1):1,B:1,D:1);']).upload(0)
    out = torch.zeros(3, dtype=torch.float64, device='cuda:0')
    before = torch.cuda.current_device()
    ts.pairwise_device(('rf',), {'rf': out.data_ptr()}, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert torch.cuda.current_device() == before
    with pytest.raises(ValueError):
        ts.check_errors(0)
    ts.check_errors(0)                                       # the flag was cleared by the read
    ok = _lib.TreeSet(['((A:1,B:1):1,C:1,D:1);', '((A:1,C:1):1,B:1,D:1);']).upload(0)
    ok.pairwise_device(('rf',), {'rf': out.data_ptr()}, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    ok.check_errors(0)
    assert out[0].item() == 2.0


def test_first_rf_only_calls_race_on_two_streams(td, oracle_mod):
    """The incidence matrix of the int8 tcgen05 path is built lazily by the first rf-only call: a second call on another stream must
    not see it half built."""
    import threading
    import torch
    from treecl_b200 import _lib
    trees = make('structured', 600, 60, seed=11)
    want = oracle_mod.matrix(trees, 'rf', nthreads=NTHREADS)
    for _ in range(3):
        ts = _lib.TreeSet(trees).upload(0)
        streams = [torch.cuda.Stream() for _ in range(2)]
        outs = [torch.full((want.size,), -1.0, dtype=torch.float64, device='cuda:0') for _ in streams]
        torch.cuda.synchronize()
        go = threading.Barrier(2)

        def work(k):
            go.wait()
            ts.pairwise_device(('rf',), {'rf': outs[k].data_ptr()}, stream=streams[k].cuda_stream)

        th = [threading.Thread(target=work, args=(k,)) for k in range(2)]
        for t in th:
            t.start()
        for t in th:
            t.join()
        torch.cuda.synchronize()
        for o in outs:
            assert np.array_equal(o.cpu().numpy(), want)
        ts.close()


def test_host_threads_with_sets_of_different_sizes(td):
    """Host threads running the same kernels for sets of different sizes at the same time: the dynamic shared memory a kernel may use
    is a property of the function, not of a launch - it is opted in once per kernel (allow_full_smem), not resized by every launch."""
    import threading
    from treecl_b200 import _lib
    cases = [(400, 12), (300, 40), (250, 64), (120, 100)]
    sets = [_lib.TreeSet(make('uniform', n, taxa, seed=40 + taxa)) for n, taxa in cases]
    metrics = ('rf', 'wrf', 'euc', 'geo')
    want = [ts.pairwise(metrics) for ts in sets]
    errors = []
    go = threading.Barrier(len(sets))

    def work(k):
        try:
            go.wait()
            for _ in range(6):
                got = sets[k].pairwise(metrics)
                for m in metrics:
                    assert np.array_equal(got[m], want[k][m]), (cases[k], m)
        except Exception as exc:                       # noqa: BLE001 - reported below, in the main thread
            errors.append(exc)

    th = [threading.Thread(target=work, args=(k,)) for k in range(len(sets))]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errors, errors


def test_polytomies_and_rooted_trees(td, oracle_mod):
    """UNPINNED by the reference (no golden): extended common-edge rule and clade (rooted) mode, oracle vs CUDA."""
    from treecl_b200 import _lib
    poly = ['(A:1,B:1,(C:1,D:1,E:1):2,F:1,G:0.5);', '((A:1,B:1):1,(C:1,D:1):1,E:1,F:1,G:0.25);',
            '((A:1,C:1):1,(B:2,D:1):0.7,(E:1,F:1):0.3,G:1);', '(A:1,B:1,C:1,D:1,E:1,F:1,G:1);',
            '(((A:1,B:1):1,C:2):1,D:1,(E:1,(F:0.1,G:1):0.0):1);']
    rooted = ['((A:1,B:2):0.5,((C:1,D:1):0.25,E:3):0.75);', '(((A:1,C:2):0.5,B:1):0.3,(D:1,E:1):0.4);',
              '((A:1,(B:1,(C:1,D:1):0.5):0.5):1,E:2);', '((A:1,B:1):0.1,(C:1,(D:1,E:1):2):0.2);']
    for trees in (poly, rooted):
        ts = _lib.TreeSet(trees)
        got = ts.pairwise(('rf', 'wrf', 'euc', 'geo'))
        for metric in ('rf', 'wrf', 'euc', 'geo'):
            want = oracle_mod.matrix(trees, metric)
            if metric == 'rf':
                assert np.array_equal(got[metric], want)
            else:
                assert_close(got[metric], want)


@pytest.mark.parametrize('pair_kernel', ['merge', 'join', 'split'], indirect=True)
def test_identical_trees_give_exact_zero(td, pair_kernel):
    from treecl_b200 import _lib, synth
    trees = synth.uniform_trees(5, 40, seed=9)
    ts = _lib.TreeSet([trees[0], trees[1], trees[0], trees[1], trees[0]])
    got = ts.pairwise(('rf', 'wrf', 'euc', 'geo'))
    sq = {m: squareform(v) for m, v in got.items()}
    for m in sq:
        assert sq[m][0, 2] == 0.0 and sq[m][1, 3] == 0.0 and sq[m][2, 4] == 0.0 and sq[m][0, 1] > 0


@pytest.mark.parametrize('pair_kernel', ['merge', 'join', 'split'], indirect=True)
def test_row_sharding_is_byte_identical(td, pair_kernel):
    """Multi-GPU path: any partition of the rows reproduces the one-shot matrix bit for bit."""
    from treecl_b200 import _lib, engine, synth
    trees = synth.uniform_trees(333, 24, seed=21)
    ts = _lib.TreeSet(trees)
    full = ts.pairwise(('rf', 'wrf', 'euc', 'geo'), normalise=True)
    n = len(trees)
    for parts in (2, 3, 8):
        for m in full:
            pieces = [ts.pairwise((m,), normalise=True, row_begin=a, row_end=b)[m]
                      for a, b in engine.balanced_row_ranges(n, parts)]
            assert np.array_equal(np.concatenate(pieces), full[m])
    # odd, non tile-aligned ranges
    a = ts.pairwise(('euc',), normalise=True, row_begin=5, row_end=131)['euc']
    p0, p1 = _lib.condensed_offset(n, 5), _lib.condensed_offset(n, 131)
    assert np.array_equal(a, full['euc'][p0:p1])
    out = engine.pairwise_condensed(ts, ('rf', 'geo'), normalise=True, devices=[0] * 3)
    assert np.array_equal(out['rf'], full['rf']) and np.array_equal(out['geo'], full['geo'])
    # a rank's two folded blocks in ONE call (td_pairwise_device2: shared events and fix-up kernels) = the same bytes
    import torch
    ts.upload(0)
    st = torch.cuda.current_stream().cuda_stream
    metrics = ('rf', 'wrf', 'euc', 'geo')
    for parts in (2, 4):
        for rank_blocks in engine.folded_row_ranges(n, parts, align=16):
            bufs = [(a, b, {m: torch.full((max(_lib.condensed_offset(n, b) - _lib.condensed_offset(n, a), 1),), -1.0, dtype=torch.float64, device='cuda:0')
                            for m in metrics}) for a, b in rank_blocks]
            ts.pairwise_device_blocks(metrics, [(a, b, {m: t[m].data_ptr() for m in metrics}) for a, b, t in bufs], normalise=True, stream=st)
            torch.cuda.synchronize()
            ts.check_errors(0)
            for a, b, t in bufs:
                p0, p1 = _lib.condensed_offset(n, a), _lib.condensed_offset(n, b)
                for m in metrics:
                    assert np.array_equal(t[m][:p1 - p0].cpu().numpy(), full[m][p0:p1]), (parts, a, b, m)


@pytest.mark.parametrize('pair_kernel', ['auto', 'merge', 'join', 'split'], indirect=True)
@pytest.mark.parametrize('kind', ['uniform', 'structured'])
def test_rectangular_variant(td, oracle_mod, pair_kernel, kind):
    """query x reference distances (treeCl/bootstrap.py:174-218) against the ORACLE's block of the square matrix,
    geodesic included."""
    import torch
    from treecl_b200 import _lib
    trees = make(kind, 150, 30, seed=33)
    split = 70
    ts = _lib.TreeSet(trees).upload(0)
    metrics = ('rf', 'wrf', 'euc', 'geo')
    outs = {m: torch.zeros(split * (150 - split), dtype=torch.float64, device='cuda:0') for m in metrics}
    ts.pairwise_rect_device(split, metrics, {m: outs[m].data_ptr() for m in metrics},
                            stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    for m in metrics:
        want = squareform(oracle_mod.matrix(trees, m, nthreads=NTHREADS))[:split, split:]
        got = outs[m].cpu().numpy().reshape(split, 150 - split)
        if m == 'rf':
            assert np.array_equal(got, want), m
        else:
            assert_close(got, want)


@pytest.mark.parametrize('kind', ['uniform', 'structured'])
def test_host_calls_in_row_bands(td, oracle_mod, monkeypatch, kind):
    """td_pairwise / td_pairwise_square / td_rf_incidence stream their result in row bands (kernels of band k under the copy of band
    k - 1), into page-locked buffers directly and into pageable ones through a staging block: every way gives the same bytes."""
    from treecl_b200 import _lib
    n = 900
    trees = make(kind, n, 24, seed=5)
    ts = _lib.TreeSet(trees)
    metrics = ('rf', 'wrf', 'euc', 'geo')
    pairs = n * (n - 1) // 2
    one = ts.pairwise(metrics, normalise=True)                                 # 3.2 MB per metric: a single band
    for m in metrics:
        want = oracle_mod.matrix(trees, m, normalise=True, nthreads=NTHREADS)
        assert np.array_equal(one[m], want) if m == 'rf' else assert_close(one[m], want) is None
    one_rf_euc = ts.pairwise(('rf', 'euc'), normalise=True)                    # (another template instantiation: compare like with like)
    monkeypatch.setenv('TREEDIST_BAND_BYTES', str(1 << 18))                    # 13 bands
    pinned = ts.pairwise(metrics, normalise=True)
    pageable = ts.pairwise(metrics, normalise=True, out={m: np.full(pairs, -1.0) for m in metrics})
    mixed_out = {'rf': np.full(pairs, -1.0), 'euc': _lib.pinned_empty(pairs)}
    mixed = ts.pairwise(('rf', 'euc'), normalise=True, out=mixed_out)
    for m in metrics:
        assert np.array_equal(pinned[m], one[m]) and np.array_equal(pageable[m], one[m]), m
    assert np.array_equal(mixed['rf'], one_rf_euc['rf']) and np.array_equal(mixed['euc'], one_rf_euc['euc'])
    p0, p1 = _lib.condensed_offset(n, 37), _lib.condensed_offset(n, 801)
    part = ts.pairwise(('wrf', 'geo'), normalise=True, row_begin=37, row_end=801)
    assert np.array_equal(part['wrf'], one['wrf'][p0:p1]) and np.array_equal(part['geo'], one['geo'][p0:p1])
    # the square form (squareform() on the device), banded, both kinds of destination
    for m in ('rf', 'euc', 'geo'):
        want = squareform(ts.pairwise((m,), normalise=True)[m])               # (a metric alone may run another kernel than in a group: like with like)
        assert_close(want, squareform(one[m]))
        assert np.array_equal(ts.pairwise_square(m, normalise=True), want), m
        assert np.array_equal(ts.pairwise_square(m, normalise=True, out=np.full((n, n), -1.0)), want), m
    monkeypatch.delenv('TREEDIST_BAND_BYTES')
    assert np.array_equal(ts.pairwise_square('wrf', normalise=True), squareform(ts.pairwise(('wrf',), normalise=True)['wrf']))
    # uint16 RF streamed to the host
    if kind == 'structured':
        rf = oracle_mod.matrix(trees, 'rf', nthreads=NTHREADS)
        monkeypatch.setenv('TREEDIST_BAND_BYTES', str(1 << 16))
        assert np.array_equal(ts.rf_incidence().astype(np.float64), rf)
        assert np.array_equal(ts.rf_incidence(out=np.zeros(pairs, dtype=np.uint16)).astype(np.float64), rf)
        assert np.array_equal(ts.rf_incidence(row_begin=100, row_end=777).astype(np.float64),
                              rf[_lib.condensed_offset(n, 100):_lib.condensed_offset(n, 777)])


def test_drop_in_call_uses_the_square_entry(td, oracle_mod):
    """Collection.get_inter_tree_distances on one GPU: N x N straight from the library (no scipy squareform), same values."""
    trees = make('uniform', 257, 20, seed=12)
    c = td.Collection(records=[td.Record('g%03d' % k, ml_tree=t) for k, t in enumerate(trees)])
    for metric in ('rf', 'geo'):
        dm = c.get_inter_tree_distances(metric, normalise=(metric == 'geo'), show_progress=False)
        want = squareform(oracle_mod.matrix(trees, metric, normalise=(metric == 'geo'), nthreads=NTHREADS))
        assert dm.values.shape == (257, 257) and dm.get_names() == c.names
        assert np.array_equal(dm.values, want) if metric == 'rf' else assert_close(dm.values, want) is None
        two = c.get_inter_tree_distances(metric, normalise=(metric == 'geo'), show_progress=False, devices=[0, 0])
        assert np.array_equal(two.values, dm.values)


class TestClassicalMDS:
    """SURVEY 8(f)-3, first slice: DistanceMatrix.embedding(dims, 'cmds') with the matrix on the device - double-centring and the leading
    eigenpairs by blocked subspace iteration on the fp64 tensor cores.  Eigenvalues <= 1e-8 relative, coordinates up to the sign of each
    axis."""

    @staticmethod
    def check(coords, vals, want_coords, want_vals, dims):
        want_vals = np.asarray(want_vals)[:dims]
        assert np.all(np.abs(vals - want_vals) <= 1e-8 * np.abs(want_vals)), (vals, want_vals)
        scale = np.abs(want_coords).max()
        for j in range(dims):
            s = np.sign(np.dot(coords[:, j], want_coords[:, j]))
            assert np.allclose(s * coords[:, j], want_coords[:, j], rtol=0, atol=1e-6 * scale), j

    def test_reference_goldens(self, td):
        """tests/golden/mds_golden.json: outputs of the reference's own functions (treeCl/distance_matrix.py:58-74,285-315)."""
        import json
        from treecl_b200 import _lib
        with open(os.path.join(GOLDEN, 'mds_golden.json')) as fh:
            cases = json.load(fh)
        for name, c in cases.items():
            m, dims = np.array(c['matrix']), c['dimensions']
            coords, vals, info = _lib.cmds_host(m, dims)
            self.check(coords, vals, np.array(c['coords']), c['eigenvalues'], dims)
            dm = td.DistanceMatrix.from_array(m, ['n%d' % k for k in range(len(m))])
            emb = dm.embedding(dims, 'cmds')
            assert emb.values.shape == (len(m), dims) and list(emb.df.index) == dm.get_names()
            assert np.array_equal(emb.values, coords)
        with pytest.raises(NotImplementedError):
            dm.embedding(3, 'spectral')
        with pytest.raises(td.OptionError):
            dm.embedding(3, 'nope')

    @pytest.mark.parametrize('n_trees,n_taxa,metric,dims', [(300, 24, 'euc', 3), (777, 40, 'geo', 4), (1500, 30, 'rf', 5)])
    def test_tree_sets_never_leave_the_device(self, td, oracle_mod, n_trees, n_taxa, metric, dims):
        """td_cmds: distances -> square -> double-centre -> eigenpairs all on the GPU, against the oracle chain on the oracle's matrix."""
        from treecl_b200 import _lib, synth
        trees = synth.structured_trees(n_trees, n_taxa, n_classes=dims + 2, class_nni=n_taxa // 3, lam=0.7, seed=41 + dims)
        ts = _lib.TreeSet(trees)
        coords, vals, info = ts.cmds(metric, dims)
        want_coords, want_vals = oracle_mod.mds.classical_mds(squareform(oracle_mod.matrix(trees, metric, nthreads=NTHREADS)), dims)
        assert info['iterations'] < 5000, info
        self.check(coords, vals, want_coords, want_vals, dims)
        # the same through the device-pointer entry on the library's own condensed matrix
        import torch
        pairs = n_trees * (n_trees - 1) // 2
        d = torch.empty(pairs, dtype=torch.float64, device='cuda:0')
        ts.upload(0)
        st = torch.cuda.current_stream().cuda_stream
        ts.pairwise_device((metric,), {metric: d.data_ptr()}, stream=st)
        out = torch.empty(n_trees * dims, dtype=torch.float64, device='cuda:0')
        v2 = np.empty(dims)
        import ctypes
        it, res = ctypes.c_int(), ctypes.c_double()
        _lib.check(_lib.load().td_cmds_device(d.data_ptr(), n_trees, dims, out.data_ptr(), v2.ctypes.data, 0, 0.0, ctypes.byref(it), ctypes.byref(res), st))
        assert np.array_equal(v2, vals) and np.array_equal(out.cpu().numpy().reshape(n_trees, dims), coords)


@pytest.mark.parametrize('kind', ['uniform', 'structured'])
def test_query_reference_block_of_two_separate_sets(td, oracle_mod, kind):
    """SURVEY 8(f)-1 as the reference uses it (treeCl/bootstrap.py:174-218): bootstrap trees against reference trees, the two lists
    encoded SEPARATELY, host buffers in and out (td_pairwise_rect).  Every entry against the oracle's pair function."""
    from treecl_b200 import _lib
    trees = make(kind, 90, 25, seed=52)
    q, r = trees[:35], trees[35:]
    want = {m: squareform(oracle_mod.matrix(trees, m, nthreads=NTHREADS))[:35, 35:] for m in ('rf', 'wrf', 'euc', 'geo')}
    tq, tr = _lib.TreeSet(q), _lib.TreeSet(r)
    got = _lib.pairwise_rect(tq, tr, ('rf', 'wrf', 'euc', 'geo'))
    for m in want:
        assert got[m].shape == (35, 55)
        assert np.array_equal(got[m], want[m]) if m == 'rf' else assert_close(got[m], want[m]) is None
    block = td.bootstrap.query_reference_distances(q, r, 'geo')
    assert_close(block, want['geo'])
    # query trees on another leaf set: the block comes from the square path (restrict-then-join), still the oracle's values
    labels = ['t%03d' % k for k in range(25)]
    q2 = [td.Tree(t).prune_to_subset(set(labels[:20])).newick for t in q[:6]]
    block2 = td.bootstrap.query_reference_distances(q2, r, 'euc')
    want2 = squareform(oracle_mod.matrix(q2 + r, 'euc', nthreads=NTHREADS))[:6, 6:]
    assert_close(block2, want2)
    with pytest.raises(NotImplementedError):
        _lib.pairwise_rect(_lib.TreeSet(q2), tr, ('rf',))                       # different taxon tables: encode together instead


class TestSpectralFrontEnd:
    """kNN distances / mask, local scale, affinity, Laplacian and its leading eigenvectors on the GPU
    (treeCl/distance_matrix.py:32-55,153-207; treeCl/clustering.py:125-303)."""

    @staticmethod
    def same_rows(coords, want, atol=1e-6):
        # (nearly) equal leading eigenvalues leave a rotation of the eigenvectors free: the inner products of the unit rows are invariant
        assert np.allclose(coords.dot(coords.T), want.dot(want.T), rtol=0, atol=atol)

    def test_reference_goldens(self, td):
        import json
        opt = td.spectral.options
        with open(os.path.join(GOLDEN, 'mds_golden.json')) as fh:
            cases = json.load(fh)
        for name, c in cases.items():
            m, dims = np.array(c['matrix']), c['dimensions']
            dm = td.DistanceMatrix.from_array(m, ['n%d' % k for k in range(len(m))])
            for sp in c['spectral']:
                s = td.spectral.Spectral(dm, opt.PRUNING_MANUAL if sp['prune_k'] else opt.PRUNING_NONE,
                                         opt.LOCAL_SCALE_MANUAL if sp['scale_k'] else opt.LOCAL_SCALE_MEDIAN,
                                         manual_pruning=sp['prune_k'] or None, manual_scale=sp['scale_k'] or None, logic=sp['logic'])
                assert np.allclose(s.affinity, np.array(sp['affinity']), rtol=1e-12, atol=1e-300), (name, sp)
                emb = s.spectral_embedding_(dims)
                assert emb.values.shape == (len(m), dims) and np.allclose(np.linalg.norm(emb.values, axis=1), 1.0)
                assert np.all(np.abs(s.eigenvalues - np.array(sp['eigenvalues'][:dims])) <= 1e-8), (name, sp['prune_k'], s.eigenvalues)
                self.same_rows(emb.values, np.array(sp['coords']))
        with pytest.raises(NotImplementedError):
            td.spectral.Spectral(dm, opt.PRUNING_ESTIMATE)

    @pytest.mark.parametrize('n_trees,n_taxa,metric,dims,prune_k,logic,scale_k', [(400, 24, 'euc', 3, 0, 'or', 0), (901, 30, 'geo', 4, 60, 'or', 7),
                                                                                    (640, 40, 'wrf', 5, 300, 'and', 0)])
    def test_tree_sets_never_leave_the_device(self, td, oracle_mod, n_trees, n_taxa, metric, dims, prune_k, logic, scale_k):
        from treecl_b200 import _lib, synth
        trees = synth.structured_trees(n_trees, n_taxa, n_classes=dims, class_nni=n_taxa // 3, lam=0.7, seed=60 + dims)
        ts = _lib.TreeSet(trees)
        coords, vals, info = ts.spectral_embedding(metric, dims, prune_k, logic, scale_k)
        sq = squareform(oracle_mod.matrix(trees, metric, nthreads=NTHREADS))
        want, want_vals = oracle_mod.mds.spectral_embedding(sq, dims, prune_k, logic, scale_k)
        assert info['iterations'] < 5000, info
        assert np.all(np.abs(vals - want_vals[:dims]) <= 1e-8), (vals, want_vals[:dims])
        self.same_rows(coords, want)


def test_baseline_config_2_full_vs_oracle(td, oracle_mod):
    """BASELINE configs[1] = the bench workload (5,000 x 50, seed of bench.py): EVERY entry of the rf, euc and wrf matrices against
    the oracle (12.5 M pairs per metric), plus the size-independent properties."""
    from treecl_b200 import _lib, synth
    n, taxa = 5000, 50
    trees = synth.uniform_trees(n, taxa, seed=0xCA55E77E + 1)
    ts = _lib.TreeSet(trees)
    got = ts.pairwise(('rf', 'euc', 'wrf'))
    rf, euc, wrf = got['rf'], got['euc'], got['wrf']
    assert rf.size == n * (n - 1) // 2
    assert np.array_equal(rf, oracle_mod.matrix(trees, 'rf', nthreads=NTHREADS))
    assert_close(euc, oracle_mod.matrix(trees, 'euc', nthreads=NTHREADS))
    assert_close(wrf, oracle_mod.matrix(trees, 'wrf', nthreads=NTHREADS))
    # the bench launch itself (rf + euc in one call) gives the same bytes
    both = ts.pairwise(('rf', 'euc'))
    assert np.array_equal(both['rf'], rf) and np.array_equal(both['euc'], euc)
    assert (rf == np.round(rf)).all() and (rf % 2 == 0).all() and rf.min() >= 0 and rf.max() <= 2 * (taxa - 3)
    assert (euc > 0).all() and (wrf >= euc).all()                   # L1 >= L2
    # order reversal: d(i,j) computed with the trees listed backwards is the same matrix transposed
    ts_rev = _lib.TreeSet(trees[::-1])
    rev = ts_rev.pairwise(('rf', 'euc'))
    idx = np.random.default_rng(0).integers(0, n, size=(2000, 2))
    idx = idx[idx[:, 0] < idx[:, 1]]
    fwd = condensed_index(n, idx[:, 0], idx[:, 1]); bwd = condensed_index(n, n - 1 - idx[:, 1], n - 1 - idx[:, 0])
    assert np.array_equal(rf[fwd], rev['rf'][bwd])
    assert_close(euc[fwd], rev['euc'][bwd], rtol=1e-12)


@pytest.mark.parametrize('kind,n_trees,n_taxa', [('structured', 300, 40), ('structured', 257, 100), ('structured', 130, 200),
                                                 ('uniform', 200, 20), ('uniform', 129, 64)])
def test_rf_incidence_tcgen05(td, oracle_mod, kind, n_trees, n_taxa):
    """int8 tcgen05 contraction X * X^T over the shared-split dictionary (f64, u16 and row-shard forms): bit-identical to the ORACLE."""
    import torch
    from treecl_b200 import _lib
    trees = make(kind, n_trees, n_taxa, seed=900 + n_taxa)
    ts = _lib.TreeSet(trees)
    want = oracle_mod.matrix(trees, 'rf', nthreads=NTHREADS)                 # never the library's own answer
    want_n = oracle_mod.matrix(trees, 'rf', normalise=True, nthreads=NTHREADS)
    assert np.array_equal(ts.pairwise(('rf',))['rf'], want) and np.array_equal(ts.pairwise(('rf',), normalise=True)['rf'], want_n)
    pairs = want.size
    st = torch.cuda.current_stream().cuda_stream
    out = torch.full((pairs,), -1.0, dtype=torch.float64, device='cuda:0')
    ts.rf_incidence_device(out.data_ptr(), stream=st)
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy(), want)
    ts.rf_incidence_device(out.data_ptr(), normalise=True, stream=st)
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy(), want_n)
    out16 = torch.zeros(pairs, dtype=torch.int16, device='cuda:0')
    ts.rf_incidence_device(out16.data_ptr(), out_is_u16=True, stream=st)
    torch.cuda.synchronize()
    assert np.array_equal(out16.cpu().numpy().astype(np.float64), want)
    # a row shard
    r0, r1 = 37, n_trees - 5
    p0, p1 = _lib.condensed_offset(n_trees, r0), _lib.condensed_offset(n_trees, r1)
    part = torch.zeros(p1 - p0, dtype=torch.float64, device='cuda:0')
    ts.rf_incidence_device(part.data_ptr(), row_begin=r0, row_end=r1, stream=st)
    torch.cuda.synchronize()
    assert np.array_equal(part.cpu().numpy(), want[p0:p1])


def test_baseline_config_1_full(td, oracle_mod):
    """BASELINE configs[0]: 200 random 20-taxon trees, rf / wrf / euc - small enough for the oracle in full."""
    from treecl_b200 import synth
    trees = synth.uniform_trees(200, 20, seed=0xCA55E77E)
    c = td.Collection(records=[td.Record('t%03d' % k, ml_tree=t) for k, t in enumerate(trees)])
    for metric in ('rf', 'wrf', 'euc'):
        for norm in (False, True):
            dm = c.get_inter_tree_distances(metric, jobhandler=td.parutils.ProcesspoolJobHandler(4), normalise=norm,
                                            batchsize=100, show_progress=False)
            want = squareform(oracle_mod.matrix(trees, metric, normalise=norm, nthreads=4))
            if metric == 'rf':
                assert np.array_equal(dm.values, want)
            else:
                assert_close(dm.values, want)
            assert dm.get_names() == c.names


def test_baseline_config_5_full_size(td, oracle_mod):
    """BASELINE configs[4]: 2,000 random 30-taxon trees, geodesic: all 1,999,000 pairs against the oracle."""
    from treecl_b200 import _lib, synth
    n = 2000
    trees = synth.uniform_trees(n, 30, seed=0xCA55E77E + 4)
    ts = _lib.TreeSet(trees)
    got = ts.pairwise(('geo', 'euc'))
    geo, euc = got['geo'], got['euc']
    assert_close(geo, oracle_mod.matrix(trees, 'geo', nthreads=NTHREADS))
    assert np.isfinite(geo).all() and (geo >= euc * (1 - 1e-12)).all()          # the geodesic is never shorter than the chord
    st = ts.geo_stats()
    assert st['pairs'] == n * (n - 1) // 2 and 2.0 < st['ratios'] / st['pairs'] < 4.0
    # order reversal gives the transposed matrix
    rev = _lib.TreeSet(trees[::-1]).pairwise(('geo',))['geo']
    idx = np.random.default_rng(1).integers(0, n, size=(3000, 2)); idx = idx[idx[:, 0] < idx[:, 1]]
    assert_close(geo[condensed_index(n, idx[:, 0], idx[:, 1])], rev[condensed_index(n, n - 1 - idx[:, 1], n - 1 - idx[:, 0])], rtol=1e-10)


def test_baseline_config_3_full_size(td, oracle_mod):
    """BASELINE configs[2] at its own size: 20,000 random 100-taxon trees, wrf + euc (200 M pairs per metric).  The oracle checks
    128 random rows of the SQUARE matrix (2.56 M pairs per metric, spread over the whole triangle); 8-way row sharding must give
    the same bytes."""
    from treecl_b200 import _lib, engine, synth
    n = 20000
    g = synth.baseline_config(2)
    trees = g.strings()
    ts = g.treeset()
    got = ts.pairwise(('wrf', 'euc'))
    rows = np.sort(np.random.default_rng(3).choice(n, size=128, replace=False))
    for m in ('wrf', 'euc'):
        assert_close(square_rows(got[m], n, rows), oracle_mod.matrix_rows(trees, m, rows, nthreads=NTHREADS))
    assert (got['wrf'] >= got['euc']).all()
    parts = [ts.pairwise(('wrf',), row_begin=a, row_end=b)['wrf'] for a, b in engine.balanced_row_ranges(n, 8)]
    assert np.array_equal(np.concatenate(parts), got['wrf'])


def test_baseline_config_4_shape(td, oracle_mod):
    """BASELINE configs[3] at its own size: 100,000 clustered 200-taxon trees, RF through the int8 tcgen05 split-incidence
    contraction.  u16: the FULL condensed matrix (5 G pairs, 10 GB) on the device; f64: a band of rows through td_pairwise (whose
    choice for rf alone is the same kernel) and through td_rf_incidence_device.  The oracle checks 24 random rows of the square
    matrix (2.4 M pairs) of the u16 matrix, and the band in f64."""
    import torch
    from treecl_b200 import _lib, synth
    n = 100000
    g = synth.baseline_config(3)
    ts = g.treeset().upload(0)
    assert ts.info.uniform == 1 and ts.info.max_splits == 197
    st = torch.cuda.current_stream().cuda_stream
    pairs = n * (n - 1) // 2
    out16 = torch.zeros(pairs, dtype=torch.int16, device='cuda:0')
    ts.rf_incidence_device(out16.data_ptr(), out_is_u16=True, stream=st)
    torch.cuda.synchronize()
    rows = np.sort(np.random.default_rng(4).choice(n, size=24, replace=False))
    r0 = int(rows[11])
    rows = np.concatenate([rows, np.arange(r0, r0 + 8)])                       # + the band [r0, r0 + 8) for the f64 forms
    trees = g.strings()
    want = oracle_mod.matrix_rows(trees, 'rf', rows, nthreads=NTHREADS)
    del trees
    cols = torch.arange(n, device='cuda:0')
    for k, r in enumerate(rows[:24]):
        other = cols[cols != int(r)]
        a, b = torch.clamp(other, max=int(r)), torch.clamp(other, min=int(r))
        idx = a * n - a * (a + 1) // 2 + (b - a - 1)
        got = out16[idx].cpu().numpy().astype(np.float64)
        assert np.array_equal(got, np.delete(want[k], int(r))), int(r)
    # f64 band, both entry points
    p0, p1 = _lib.condensed_offset(n, r0), _lib.condensed_offset(n, r0 + 8)
    band = torch.zeros(p1 - p0, dtype=torch.float64, device='cuda:0')
    ts.rf_incidence_device(band.data_ptr(), row_begin=r0, row_end=r0 + 8, stream=st)
    torch.cuda.synchronize()
    host = ts.pairwise(('rf',), row_begin=r0, row_end=r0 + 8)['rf']
    at = 0
    for k in range(8):
        i = r0 + k
        w = want[24 + k, i + 1:]
        assert np.array_equal(band[at:at + w.size].cpu().numpy(), w) and np.array_equal(host[at:at + w.size], w), i
        assert np.array_equal(out16[p0 + at:p0 + at + w.size].cpu().numpy().astype(np.float64), w)
        at += w.size
    assert at == p1 - p0


def test_concurrent_launches_on_one_set_share_the_event_buffer_safely(td, monkeypatch):
    """The postings pipeline keeps one event buffer per resident set: launches from different streams / host threads on the same set
    are ordered behind each other (cudaStreamWaitEvent), so they must give the same bytes as a lone launch."""
    import threading
    import torch
    from treecl_b200 import _lib, synth
    monkeypatch.setenv('TREEDIST_PAIR_KERNEL', 'split')
    trees = synth.uniform_trees(700, 24, seed=77)
    ts = _lib.TreeSet(trees)
    ts.upload(0)
    pairs = 700 * 699 // 2
    ref = {m: torch.empty(pairs, dtype=torch.float64, device='cuda') for m in ('rf', 'wrf', 'euc')}
    ts.pairwise_device(('rf', 'wrf', 'euc'), {m: ref[m].data_ptr() for m in ref})
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream() for _ in range(4)]
    outs = [{m: torch.full((pairs,), -1.0, dtype=torch.float64, device='cuda') for m in ('rf', 'wrf', 'euc')} for _ in streams]
    torch.cuda.synchronize()

    def work(k):
        for _ in range(5):
            ts.pairwise_device(('rf', 'wrf', 'euc'), {m: outs[k][m].data_ptr() for m in outs[k]}, stream=streams[k].cuda_stream)

    threads = [threading.Thread(target=work, args=(k,)) for k in range(len(streams))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    torch.cuda.synchronize()
    for k in range(len(streams)):
        for m in ('rf', 'wrf', 'euc'):
            assert torch.equal(outs[k][m], ref[m]), (k, m)


def test_randomised_cross_check_of_the_pair_kernels(td):
    """A few seconds of tools/fuzz_parity.py: random sizes / taxa / sharing / normalisation / row ranges / rectangular splits,
    postings pipeline against the merge kernel (rf bit for bit, wrf / euc to 1e-9) and against the oracle for the small sets."""
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'tools', 'fuzz_parity.py'), '8', '7'], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert 'fuzz ok' in out.stdout
