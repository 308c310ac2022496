"""Synthetic tree sets for benchmarks and parity tests.

uniform_trees     emulates treeCl.tree.RandomTree.new(n, rooted=False) (treeCl/tree.py:1234-1276): start from a
                  3-leaf star, attach each remaining leaf to a uniformly random existing edge => binary tree with a
                  trifurcating root, n-3 internal edges; branch lengths i.i.d. Exp(1) = randomise_branch_lengths
                  defaults gammavariate(1,1) (treeCl/tree.py:1003-1025).
structured_trees  follows Simulator.generate_class_trees (treeCl/simulator.py:127-190): K class trees = master tree +
                  random NNIs; every gene tree = its class tree + Poisson(lam) extra NNIs, lengths redrawn.
Lengths are written with repr() (shortest round trip), so every consumer parses identical doubles.
"""
import numpy as np

from .constants import RANDOM_SEED


def taxon_labels(n):
    width = max(3, len(str(n - 1)))
    return ['t{:0{w}d}'.format(i, w=width) for i in range(n)]


class _Topo(object):
    """Rooted-at-a-trifurcation topology: node 0 is the root; children lists; leaves carry a taxon index."""

    __slots__ = ('children', 'parent', 'taxon')

    def __init__(self):
        self.children, self.parent, self.taxon = [[]], [-1], [-1]

    def add(self, parent, taxon=-1):
        self.children.append([])
        self.parent.append(parent)
        self.taxon.append(taxon)
        k = len(self.parent) - 1
        if parent >= 0:
            self.children[parent].append(k)
        return k

    def copy(self):
        t = _Topo()
        t.children = [list(c) for c in self.children]
        t.parent = list(self.parent)
        t.taxon = list(self.taxon)
        return t


def _random_topology(n, rng):
    order = rng.permutation(n)
    t = _Topo()
    for k in range(3):
        t.add(0, int(order[k]))
    for k in range(3, n):
        v = int(rng.integers(1, len(t.parent)))          # edge above node v, uniform over existing edges
        p = t.parent[v]
        mid = t.add(-1)
        t.parent[mid] = p
        t.children[p][t.children[p].index(v)] = mid
        t.children[mid].append(v)
        t.parent[v] = mid
        t.add(mid, int(order[k]))
    return t


def _nni(t, rng):
    """One random nearest-neighbour interchange around a random internal edge."""
    internal = [v for v in range(1, len(t.parent)) if t.children[v]]
    if not internal:
        return
    v = internal[int(rng.integers(len(internal)))]
    p = t.parent[v]
    sibs = [c for c in t.children[p] if c != v]
    a = sibs[int(rng.integers(len(sibs)))]                 # a sibling of v
    b = t.children[v][int(rng.integers(len(t.children[v])))]   # a child of v
    t.children[p][t.children[p].index(a)] = b
    t.children[v][t.children[v].index(b)] = a
    t.parent[a], t.parent[b] = v, p


def _newick(t, labels, lengths):
    out = []
    # iterative post-order serialisation
    stack = [(0, 0)]
    while stack:
        v, i = stack.pop()
        ch = t.children[v]
        if not ch:
            out.append(labels[t.taxon[v]])
            out.append(':' + repr(float(lengths[v])))
            continue
        if i == 0:
            out.append('(')
        if i < len(ch):
            if i > 0:
                out.append(',')
            stack.append((v, i + 1))
            stack.append((ch[i], 0))
        else:
            out.append(')')
            if v != 0:
                out.append(':' + repr(float(lengths[v])))
    out.append(';')
    return ''.join(out)


def uniform_trees(n_trees, n_taxa, seed=RANDOM_SEED, labels=None):
    """n_trees independent uniform-attachment binary unrooted trees on one taxon set, Exp(1) branch lengths."""
    labels = labels or taxon_labels(n_taxa)
    trees = []
    for k in range(n_trees):
        rng = np.random.default_rng([int(seed), k])
        t = _random_topology(n_taxa, rng)
        lengths = rng.exponential(1.0, size=len(t.parent))
        trees.append(_newick(t, labels, lengths))
    return trees


def structured_trees(n_trees, n_taxa, n_classes=8, class_nni=10, lam=1.0, seed=RANDOM_SEED, labels=None):
    """Clustered gene-tree set: shared structure => small global split dictionary (int8 incidence regime)."""
    labels = labels or taxon_labels(n_taxa)
    rng0 = np.random.default_rng([int(seed), 0xC1A55])
    master = _random_topology(n_taxa, rng0)
    classes = []
    for _ in range(n_classes):
        c = master.copy()
        for _ in range(class_nni):
            _nni(c, rng0)
        classes.append(c)
    trees = []
    for k in range(n_trees):
        rng = np.random.default_rng([int(seed), 1 + k])
        t = classes[k % n_classes].copy()
        for _ in range(int(rng.poisson(lam))):
            _nni(t, rng)
        lengths = rng.exponential(1.0, size=len(t.parent))
        trees.append(_newick(t, labels, lengths))
    return trees
