"""Tree: the slice of treeCl.Tree (treeCl/tree.py:742-1001) that the distance path uses.

The reference wraps a dendropy tree and lazily a C++ PhyloTree; every pair re-parses both.  Here a Tree is a Newick
string plus what the host encoder of libtreedist_cuda derives from it; nothing is computed per pair on the host.
"""
from . import _lib


class Tree(object):
    def __init__(self, newick=None, name=None, **kwargs):
        if newick is not None and not isinstance(newick, str):
            newick = newick.decode()
        self._newick = newick.strip() if newick else None
        self.name = name
        self._labels = None
        self._rooted = None

    def __repr__(self):
        return '{0}{1}'.format(self.__class__.__name__, self.newick if self.newick else '(None)')

    def __str__(self):
        return 'Tree Object: {}\n{}'.format(self.name, self.newick)

    def _scan(self):
        if self._labels is None and self._newick:
            labels, rooted = _lib.newick_labels(self._newick)     # ValueError on malformed Newick
            self._labels, self._rooted = set(labels), rooted

    def __len__(self):
        """Number of leaves (tree.py:777-780)."""
        return len(self.labels)

    def __and__(self, other):
        return self.intersection(other)

    def __xor__(self, other):
        return self.labels ^ other.labels

    @property
    def labels(self):
        """Leaf label set (tree.py:792-796)."""
        self._scan()
        return set(self._labels) if self._labels is not None else set()

    @property
    def newick(self):
        return self._newick

    @property
    def rooted(self):
        """Rooted <=> the root has exactly two children (tree.py:858-862)."""
        if not self._newick:
            return None
        self._scan()
        return self._rooted

    def intersection(self, other):
        """Leaves in common (tree.py:960-964)."""
        return self.labels & other.labels

    def copy(self):
        return self.__class__(self._newick, self.name)

    def prune_to_subset(self, subset, inplace=False):
        """Prune to the taxa in `subset`; merged edges add their lengths (tree.py:989-1001)."""
        subset = set(subset)
        if not subset.issubset(self.labels):
            print('"subset" is not a subset')
            return
        pruned = _lib.newick_prune(self._newick, sorted(subset))
        if inplace:
            self._newick, self._labels, self._rooted = pruned, None, None
            return self
        return self.__class__(pruned, self.name)
