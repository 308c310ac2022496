"""Task interfaces of the distance path (treeCl/tasks.py:41-75,648-718), kept for API compatibility.

In the reference a task is a per-pair function taking two Newick strings; get_inter_tree_distances maps it over all
pairs through a JobHandler.  The GPU engine computes the whole matrix in one call, so these classes only name the
metric (`.metric`) and still offer the per-pair callables for code that used them directly.
"""
import itertools
from abc import ABCMeta, abstractmethod

from . import treedist


def eucdist_task(newick_string_a, newick_string_b, normalise, min_overlap=4, overlap_fail_value=0):
    return treedist.eucdist(newick_string_a, newick_string_b, normalise, min_overlap, overlap_fail_value)


def geodist_task(newick_string_a, newick_string_b, normalise, min_overlap=4, overlap_fail_value=0):
    return treedist.geodist(newick_string_a, newick_string_b, normalise, min_overlap, overlap_fail_value)


def rfdist_task(newick_string_a, newick_string_b, normalise, min_overlap=4, overlap_fail_value=0):
    return treedist.rfdist(newick_string_a, newick_string_b, normalise, min_overlap, overlap_fail_value)


def wrfdist_task(newick_string_a, newick_string_b, normalise, min_overlap=4, overlap_fail_value=0):
    return treedist.wrfdist(newick_string_a, newick_string_b, normalise, min_overlap, overlap_fail_value)


class TreeDistanceTaskInterface(metaclass=ABCMeta):
    _name = 'TreeDistance'
    metric = None

    @property
    def name(self):
        return self._name

    def scrape_args(self, trees, normalise, min_overlap, overlap_fail_value):
        for (t1, t2) in itertools.combinations(trees, 2):
            yield (t1, t2, normalise, min_overlap, overlap_fail_value)

    @abstractmethod
    def get_task(self):
        pass


class GeodesicTreeDistance(TreeDistanceTaskInterface):
    _name, metric = 'GeodesicDistance', 'geo'

    def get_task(self):
        return geodist_task


class RobinsonFouldsTreeDistance(TreeDistanceTaskInterface):
    _name, metric = 'RFDistance', 'rf'

    def get_task(self):
        return rfdist_task


class WeightedRobinsonFouldsTreeDistance(TreeDistanceTaskInterface):
    _name, metric = 'WeightedRFDistance', 'wrf'

    def get_task(self):
        return wrfdist_task


class EuclideanTreeDistance(TreeDistanceTaskInterface):
    _name, metric = 'EuclideanDistance', 'euc'

    def get_task(self):
        return eucdist_task


# 'fast*' aliases: "same leaf set asserted" variants.  They are dead code in the reference (collection.py:448-455
# builds a generator and then calls len() on it); here they are plain aliases of the checked versions.
class EqualLeafSetGeodesicTreeDistance(GeodesicTreeDistance):
    pass


class EqualLeafSetEuclideanTreeDistance(EuclideanTreeDistance):
    pass


class EqualLeafSetRobinsonFouldsTreeDistance(RobinsonFouldsTreeDistance):
    _name = 'RobinsonFouldsDistance'


class EqualLeafSetWeightedRobinsonFouldsTreeDistance(WeightedRobinsonFouldsTreeDistance):
    _name = 'WeightedRobinsonFouldsDistance'
