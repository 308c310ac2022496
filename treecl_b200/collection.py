"""Collection: the slice of treeCl.Collection that feeds and owns get_inter_tree_distances.

Mirrors treeCl/collection.py:75-153 (record loading in SORT_KEY order, `trees`, `names`) and :421-456
(get_inter_tree_distances).  Alignment parsing, tree inference, scoring and the CEM optimiser are not on the distance
path and stay in treeCl; a record here only carries `name` and `parameters.{ml_tree,nj_tree}`.
"""
import bz2
import glob
import gzip
import itertools
import json
import logging
import os

from scipy.spatial.distance import squareform

from . import _lib, engine, tasks
from .constants import SORT_KEY
from .distance_matrix import DistanceMatrix
from .errors import optioncheck
from .parutils import SequentialJobHandler

logger = logging.getLogger(__name__)
default_jobhandler = SequentialJobHandler()


def _read_text(filename):
    if filename.endswith('.gz'):
        with gzip.open(filename) as fh:
            return fh.read().decode('utf-8')
    if filename.endswith('.bz2'):
        with bz2.open(filename) as fh:
            return fh.read().decode('utf-8')
    with open(filename) as fh:
        return fh.read()


def strip_extensions(filename):
    """treeCl/utils/fileIO.py:229-233"""
    toplevel = os.path.splitext(os.path.basename(filename))
    while toplevel[1] > '':
        toplevel = os.path.splitext(toplevel[0])
    return toplevel[0]


class Parameters(object):
    """ml_tree / nj_tree slots of treeCl/parameters.py:110-168 (the only ones the distance path reads)."""

    def __init__(self, ml_tree=None, nj_tree=None):
        self.ml_tree = ml_tree
        self.nj_tree = nj_tree

    def construct_from_dict(self, d):
        for key in ('ml_tree', 'nj_tree'):
            if d.get(key) is not None:
                setattr(self, key, str(d[key]))


class Record(object):
    """Stand-in for treeCl.Alignment on this path: `.name`, `.parameters`, `.tree` (treeCl/alignment.py:108-124)."""

    def __init__(self, name, ml_tree=None, nj_tree=None):
        self.name = name
        self.parameters = Parameters(ml_tree, nj_tree)

    @property
    def tree(self):
        if self.parameters.ml_tree is not None:
            return self.parameters.ml_tree
        elif self.parameters.nj_tree is not None:
            return self.parameters.nj_tree
        else:
            raise AttributeError('No tree')


class NoRecordsError(Exception):
    def __init__(self, file_format, input_dir):
        self.file_format = file_format
        self.input_dir = input_dir

    def __str__(self):
        return 'No records were found in {0} matching\n\tfile_format = {1}\n'.format(self.input_dir, self.file_format)


class Collection(object):
    """Collection(records=... | input_dir=..., param_dir=... | trees_dir=..., file_format='phylip')

    input_dir is only listed for alignment file names (they define the record names and their SORT_KEY order, as in
    treeCl/collection.py:207-280); the alignments themselves are not parsed.  If input_dir is omitted, the names come
    from the files of trees_dir / param_dir.
    """

    def __init__(self, records=None, input_dir=None, param_dir=None, trees_dir=None, file_format='phylip',
                 header_grep=None, show_progress=True):
        self.show_progress = show_progress
        if records is not None:
            self.records = list(records)
        elif input_dir is not None:
            if not os.path.isdir(input_dir):
                raise IOError('{} is not a directory'.format(input_dir))
            optioncheck(file_format, ['fasta', 'phylip'])
            self.records = [Record(n) for n in self._alignment_names(os.path.abspath(input_dir), file_format)]
        elif param_dir is not None or trees_dir is not None:
            d, pat = (param_dir, '*.json*') if param_dir is not None else (trees_dir, '*.nwk*')
            names = sorted({strip_extensions(f) for f in glob.glob(os.path.join(d, pat))}, key=SORT_KEY)
            self.records = [Record(n) for n in names]
        else:
            raise ValueError('Provide a list of records, or the path to a set of alignments')

        if param_dir is not None and trees_dir is None:
            self.read_parameters(param_dir)
        elif trees_dir is not None and param_dir is None:
            self.read_trees(trees_dir)
        if not self.records:
            raise NoRecordsError(file_format, input_dir)

    @staticmethod
    def _alignment_names(input_dir, file_format):
        extensions = ['phy'] if file_format == 'phylip' else ['fa', 'fas', 'fasta']
        extensions = ['.'.join([x, y]) if y else x for x, y in itertools.product(extensions, ['', 'gz', 'bz2'])]
        files = []
        for ex in extensions:
            files.extend(glob.glob('{0}/*.{1}'.format(input_dir, ex)))
        files.sort(key=SORT_KEY)
        return [strip_extensions(f) for f in files]

    def __len__(self):
        return len(self.records)

    def __getitem__(self, i):
        return self.records[i]

    def __iter__(self):
        return iter(self.records)

    @property
    def trees(self):
        """Newick strings in record order (treeCl/collection.py:137-143)."""
        return [rec.tree for rec in self.records]

    @property
    def names(self):
        return [rec.name for rec in self.records]

    def read_trees(self, input_dir):
        """treeCl/collection.py:283-310"""
        for rec in self.records:
            filename = glob.glob(os.path.join(input_dir, '{}.nwk*'.format(rec.name)))
            try:
                rec.parameters.construct_from_dict(dict(ml_tree=_read_text(filename[0])))
            except (IOError, IndexError):
                continue

    def read_parameters(self, input_dir):
        """treeCl/collection.py:312-334"""
        for rec in self.records:
            filename = glob.glob(os.path.join(input_dir, '{}.json*'.format(rec.name)))
            try:
                rec.parameters.construct_from_dict(json.loads(_read_text(filename[0])))
            except (IOError, IndexError):
                continue

    def get_inter_tree_distances(self, metric, jobhandler=default_jobhandler, normalise=False, min_overlap=4,
                                 overlap_fail_value=0, batchsize=1, show_progress=True, devices=None):
        """Generate a distance matrix from a fully-populated Collection (treeCl/collection.py:421-456).

        :param metric: 'euc', 'geo', 'rf', 'wrf' (and the 'fast*' aliases).
        :param jobhandler, batchsize, show_progress: accepted for compatibility, ignored - the matrix is one GPU job.
        :param normalise, min_overlap, overlap_fail_value: as in the reference (overlap_fail_value is coerced to float).
        :param devices: optional list of CUDA device ordinals to shard the matrix over (extension).
        :return: DistanceMatrix with names == self.names.

        Needs a CUDA device (no CPU fallback).  Trees of any size are accepted: up to 128 internal edges per tree run on the fast
        kernels, larger ones (to 1,024 internal edges) on a slower warp-per-pair kernel; trees that differ in leaf set are grouped by
        leaf set and only the pairs across groups are pruned to their common leaves (treeCl/treedist.py:63-68).
        """
        metrics = {'euc': tasks.EuclideanTreeDistance,
                   'geo': tasks.GeodesicTreeDistance,
                   'rf': tasks.RobinsonFouldsTreeDistance,
                   'wrf': tasks.WeightedRobinsonFouldsTreeDistance,
                   'fasteuc': tasks.EqualLeafSetEuclideanTreeDistance,
                   'fastgeo': tasks.EqualLeafSetGeodesicTreeDistance,
                   'fastrf': tasks.EqualLeafSetRobinsonFouldsTreeDistance,
                   'fastwrf': tasks.EqualLeafSetWeightedRobinsonFouldsTreeDistance}
        optioncheck(metric, list(metrics.keys()))
        task_interface = metrics[metric]()
        trees = self.trees
        ts = _lib.TreeSet(trees)
        try:
            if devices is None or len(devices) == 1:
                # one GPU: squareform() (collection.py:456) is done on the device, the N x N matrix lands in page-locked memory
                square = ts.pairwise_square(task_interface.metric, normalise, min_overlap, float(overlap_fail_value),
                                            engine.default_device() if devices is None else devices[0])
            else:
                square = squareform(engine.pairwise_condensed(ts, (task_interface.metric,), normalise, min_overlap,
                                                              float(overlap_fail_value), devices)[task_interface.metric])
        finally:
            ts.close()
        return DistanceMatrix.from_array(square, self.names)
