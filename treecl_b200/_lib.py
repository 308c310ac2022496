"""ctypes binding of libtreedist_cuda.so (include/treedist_cuda.h).

There is deliberately no fallback: if the shared library is missing or no B200 is visible, the compute entry
points raise.  The CPU oracle under oracle/ is test infrastructure and is never imported from here.
"""
import ctypes
import os
import weakref

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('TREEDIST_LIB') or os.path.join(_HERE, 'libtreedist_cuda.so')     # (override: A/B of builds)

TD_RF, TD_WRF, TD_EUC, TD_GEO = 0, 1, 2, 3
METRIC_INDEX = {'rf': TD_RF, 'wrf': TD_WRF, 'euc': TD_EUC, 'geo': TD_GEO}

TD_OK, TD_ERR_PARSE, TD_ERR_MIXED_ROOTING, TD_ERR_ARG, TD_ERR_NOMEM, TD_ERR_CUDA, TD_ERR_NO_DEVICE, \
    TD_ERR_UNSUPPORTED, TD_ERR_NOT_UPLOADED = range(9)

(TD_GET_LABELS, TD_GET_NSPLITS, TD_GET_MASKS, TD_GET_LENGTHS, TD_GET_LEAFLEN, TD_GET_IDS, TD_GET_LEAFMASK,
 TD_GET_ROOTED) = range(8)


class TreedistError(RuntimeError):
    """CUDA / library failure (no CPU fallback exists)."""


class NoDeviceError(TreedistError):
    pass


class td_info(ctypes.Structure):
    _fields_ = [('n_trees', ctypes.c_int64), ('n_taxa', ctypes.c_int32), ('words', ctypes.c_int32),
                ('max_splits', ctypes.c_int32), ('max_shared', ctypes.c_int32), ('uniform', ctypes.c_int32),
                ('rooted', ctypes.c_int32), ('dict_size', ctypes.c_int64), ('dict_shared', ctypes.c_int64),
                ('total_splits', ctypes.c_int64), ('device', ctypes.c_int32), ('n_postings', ctypes.c_int32),
                ('heavy_rows', ctypes.c_int32), ('n_groups', ctypes.c_int32), ('mean_common', ctypes.c_double),
                ('heavy_common', ctypes.c_double)]


class td_geo_stats(ctypes.Structure):
    _fields_ = [('pairs', ctypes.c_uint64), ('maxflow_solves', ctypes.c_uint64), ('ratios', ctypes.c_uint64),
                ('augmentations', ctypes.c_uint64), ('launches', ctypes.c_uint64)]


_lib = None
_c_void_p4 = ctypes.c_void_p * 4


def load():
    """Load the shared library; raises if it was not built (run `python -c 'import __graft_entry__ as g; g.build()'`)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise TreedistError('libtreedist_cuda.so is missing at {} - build it (make -C treecl_b200/csrc); '
                            'there is no CPU fallback'.format(LIB_PATH))
    lib = ctypes.CDLL(LIB_PATH)
    vp, i64, i32, u32, dbl = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_uint32, ctypes.c_double
    lib.td_last_error.restype = ctypes.c_char_p
    lib.td_version.restype = ctypes.c_char_p
    lib.td_launch_count.restype = ctypes.c_uint64
    lib.td_encode_newick.argtypes = [ctypes.POINTER(ctypes.c_char_p), i64, i32, ctypes.POINTER(vp)]
    lib.td_encode_newick_joined.argtypes = [ctypes.c_char_p, i64, i64, i32, ctypes.POINTER(vp)]
    lib.td_encode_kendall_colijn.argtypes = [ctypes.POINTER(ctypes.c_char_p), i64, dbl, i32, ctypes.POINTER(vp)]
    lib.td_treeset_free.argtypes = [vp]
    lib.td_treeset_free.restype = None
    lib.td_treeset_info.argtypes = [vp, ctypes.POINTER(td_info)]
    lib.td_treeset_get.argtypes = [vp, i32, vp, i64, ctypes.POINTER(i64)]
    lib.td_newick_labels.argtypes = [ctypes.c_char_p, ctypes.c_char_p, i64, ctypes.POINTER(i64), ctypes.POINTER(i32)]
    lib.td_newick_prune.argtypes = [ctypes.c_char_p, ctypes.POINTER(ctypes.c_char_p), i64, ctypes.c_char_p, i64,
                                    ctypes.POINTER(i64)]
    lib.td_device_count.argtypes = [ctypes.POINTER(i32)]
    lib.td_upload.argtypes = [vp, i32, vp]
    lib.td_condensed_offset.argtypes = [i64, i64]
    lib.td_condensed_offset.restype = i64
    lib.td_pairwise_device.argtypes = [vp, u32, i32, i32, dbl, i64, i64, _c_void_p4, vp]
    lib.td_pairwise.argtypes = [vp, u32, i32, i32, dbl, i64, i64, _c_void_p4, i32]
    if hasattr(lib, 'td_pairwise_rect'):
        lib.td_pairwise_rect.argtypes = [vp, vp, u32, i32, _c_void_p4, i32]
    if hasattr(lib, 'td_pairwise_device2'):
        lib.td_pairwise_device2.argtypes = [vp, u32, i32, i32, dbl, i64 * 2, i64 * 2, _c_void_p4, _c_void_p4, vp]
    lib.td_pairwise_rect_device.argtypes = [vp, i64, u32, i32, _c_void_p4, vp]
    lib.td_rf_incidence_device.argtypes = [vp, i32, i64, i64, vp, i32, vp]
    lib.td_geo_get_stats.argtypes = [vp, ctypes.POINTER(td_geo_stats)]
    if hasattr(lib, 'td_treeset_serialize'):
        lib.td_treeset_serialize.argtypes = [vp, vp, i64, ctypes.POINTER(i64)]
        lib.td_treeset_deserialize.argtypes = [vp, i64, ctypes.POINTER(vp)]
    if hasattr(lib, 'td_pairwise_square'):
        lib.td_pairwise_square.argtypes = [vp, i32, i32, i32, dbl, vp, i32]
        lib.td_rf_incidence.argtypes = [vp, i64, i64, vp, i32]
        lib.td_alloc_pinned.argtypes = [i64, ctypes.POINTER(vp)]
        lib.td_free_pinned.argtypes = [vp]
    if hasattr(lib, 'td_cmds'):
        pd_, pi_ = ctypes.POINTER(dbl), ctypes.POINTER(i32)
        lib.td_cmds_device.argtypes = [vp, i64, i32, vp, vp, i32, dbl, pi_, pd_, vp]
        lib.td_cmds.argtypes = [vp, i32, i32, i32, dbl, i32, vp, vp, i32, dbl, pi_, pd_, i32]
        lib.td_cmds_host.argtypes = [vp, i64, i32, vp, vp, i32, dbl, pi_, pd_, i32]
    if hasattr(lib, 'td_spectral_embedding'):
        pd_, pi_ = ctypes.POINTER(dbl), ctypes.POINTER(i32)
        lib.td_affinity_host.argtypes = [vp, i64, i32, i32, i32, vp, i32]
        lib.td_spectral_embedding_host.argtypes = [vp, i64, i32, i32, i32, i32, vp, vp, i32, dbl, pi_, pd_, i32]
        lib.td_spectral_embedding.argtypes = [vp, i32, i32, i32, i32, i32, i32, vp, vp, i32, dbl, pi_, pd_, i32]
    if hasattr(lib, 'td_check_errors'):         # (absent from older builds loaded through TREEDIST_LIB for A/B timing)
        lib.td_check_errors.argtypes = [vp, i32]
        lib.td_trim_cache.argtypes = [i32]
        lib.td_synth_trees.argtypes = [i32, i64, i32, ctypes.c_uint64, i32, i32, dbl, i32, ctypes.POINTER(vp), ctypes.POINTER(i64)]
        lib.td_synth_free.argtypes = [vp]
        lib.td_synth_free.restype = None
    _lib = lib
    return lib


_glue = False


def _pyglue():
    """The optional CPython helper (csrc/pyglue.c); None when it was not built or TREEDIST_NO_PYGLUE is set."""
    global _glue
    if _glue is False:
        _glue = None
        if not os.environ.get('TREEDIST_NO_PYGLUE'):
            try:
                from . import _pyglue as mod
                _glue = mod
            except ImportError:
                pass
    return _glue


def check(rc):
    if rc == TD_OK:
        return
    msg = load().td_last_error().decode(errors='replace')
    if rc in (TD_ERR_PARSE, TD_ERR_MIXED_ROOTING, TD_ERR_ARG):
        raise ValueError(msg)                 # tree_distance surfaces C++ exceptions as ValueError (treeCl/tree.py:842)
    if rc == TD_ERR_NOMEM:
        raise MemoryError(msg)
    if rc == TD_ERR_NO_DEVICE:
        raise NoDeviceError(msg)
    if rc == TD_ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    raise TreedistError(msg)


def _b(s):
    return s.encode() if isinstance(s, str) else s


PINNED_RESULT_MAX = int(os.environ.get('TREEDIST_PINNED_RESULT_MAX', 4 << 30))


def pinned_empty(shape, dtype=np.float64):
    """Uninitialised ndarray in page-locked host memory from the library's cache (td_alloc_pinned): the device-to-host copies of
    td_pairwise / td_pairwise_square go straight into it at PCIe rate.  The block returns to the cache when the array (and every view
    of it) is gone."""
    lib = load()
    count = int(np.prod(shape))
    nbytes = max(count * np.dtype(dtype).itemsize, 1)
    ptr = ctypes.c_void_p()
    check(lib.td_alloc_pinned(nbytes, ctypes.byref(ptr)))
    buf = (ctypes.c_char * nbytes).from_address(ptr.value)
    weakref.finalize(buf, lib.td_free_pinned, ctypes.c_void_p(ptr.value))
    return np.frombuffer(buf, dtype=dtype, count=count).reshape(shape)


def result_empty(shape):
    """Result buffer of a host call: page-locked when a GPU is there and the result is not huge, plain numpy otherwise."""
    if int(np.prod(shape)) * 8 <= PINNED_RESULT_MAX:
        try:
            return pinned_empty(shape)
        except (NoDeviceError, MemoryError):
            pass
    return np.empty(shape, dtype=np.float64)


def condensed_offset(n, row):
    return int(row) * int(n) - int(row) * (int(row) + 1) // 2


class TreeSet(object):
    """N trees encoded once: canonical split bitmasks + branch lengths over a shared taxon index.

    Replaces the N*(N-1) PhyloTree(newick, rooted) constructions of the reference (treeCl/tree.py:830-845 reached
    from treeCl/tasks.py:41-75 for every pair).
    """

    def __init__(self, newicks, n_threads=0):
        lib = load()
        newicks = list(newicks)
        handle = ctypes.c_void_p()
        glue = _pyglue()
        if glue is not None and newicks:
            # the interpreter's own UTF-8 buffers go to the library: no join, no copies
            rc, h = glue.encode_list(newicks, ctypes.cast(lib.td_encode_newick, ctypes.c_void_p).value, int(n_threads))
            check(rc)
            handle = ctypes.c_void_p(h)
        elif newicks and all(isinstance(s, str) for s in newicks):
            # one join + one encode instead of one bytes object and one pointer per tree
            joined = ('\0'.join(newicks) + '\0').encode()        # the library checks that it holds exactly len(newicks) strings
            check(lib.td_encode_newick_joined(joined, len(joined), len(newicks), int(n_threads), ctypes.byref(handle)))
        else:
            raw = [_b(s) for s in newicks]
            arr = (ctypes.c_char_p * len(raw))(*raw)
            check(lib.td_encode_newick(arr, len(raw), int(n_threads), ctypes.byref(handle)))
        self._h = handle
        self._lib = lib
        self.n_trees = len(newicks)

    @classmethod
    def from_joined(cls, buffer, total_bytes, n_trees, n_threads=0):
        """Encode n_trees NUL-terminated Newick strings stored back to back (bytes object or an address + size)."""
        lib = load()
        handle = ctypes.c_void_p()
        src = buffer if isinstance(buffer, (bytes, bytearray)) else ctypes.cast(ctypes.c_void_p(buffer), ctypes.c_char_p)
        check(lib.td_encode_newick_joined(src, int(total_bytes), int(n_trees), int(n_threads), ctypes.byref(handle)))
        self = cls.__new__(cls)
        self._h, self._lib, self.n_trees = handle, lib, int(n_trees)
        return self

    def serialize(self):
        """The encoded set as a numpy uint8 array (td_treeset_serialize): broadcast it and TreeSet.deserialize() on the other ranks."""
        need = ctypes.c_int64()
        check(self._lib.td_treeset_serialize(self._h, None, 0, ctypes.byref(need)))
        buf = np.empty(need.value, dtype=np.uint8)
        check(self._lib.td_treeset_serialize(self._h, buf.ctypes.data, need.value, None))
        return buf

    @classmethod
    def deserialize(cls, buf):
        lib = load()
        buf = np.ascontiguousarray(buf, dtype=np.uint8)
        handle = ctypes.c_void_p()
        check(lib.td_treeset_deserialize(buf.ctypes.data, buf.size, ctypes.byref(handle)))
        self = cls.__new__(cls)
        self._h, self._lib = handle, lib
        self.n_trees = int(self.info.n_trees)
        return self

    @classmethod
    def kendall_colijn(cls, newicks, lbda=0.5, n_threads=0):
        """Set of Kendall-Colijn vectors v = (1 - lbda) m + lbda M (treeCl/utils/kendallcolijn.py:53-79): its 'euc' matrix is the
        Kendall-Colijn distance matrix.  All trees must have one leaf set."""
        lib = load()
        raw = [_b(s) for s in newicks]
        arr = (ctypes.c_char_p * len(raw))(*raw)
        handle = ctypes.c_void_p()
        check(lib.td_encode_kendall_colijn(arr, len(raw), float(lbda), int(n_threads), ctypes.byref(handle)))
        self = cls.__new__(cls)
        self._h, self._lib, self.n_trees = handle, lib, len(raw)
        return self

    def close(self):
        if getattr(self, '_h', None):
            self._lib.td_treeset_free(self._h)
            self._h = None

    __del__ = close

    @property
    def info(self):
        inf = td_info()
        check(self._lib.td_treeset_info(self._h, ctypes.byref(inf)))
        return inf

    def get(self, what, dtype):
        need = ctypes.c_int64()
        check(self._lib.td_treeset_get(self._h, what, None, 0, ctypes.byref(need)))
        if what == TD_GET_LABELS:
            buf = ctypes.create_string_buffer(max(need.value, 1))
            check(self._lib.td_treeset_get(self._h, what, buf, need.value, None))
            raw = buf.raw[:need.value].decode()
            return raw.split('\n') if raw else []
        out = np.empty(need.value // np.dtype(dtype).itemsize, dtype=dtype)
        check(self._lib.td_treeset_get(self._h, what, out.ctypes.data, need.value, None))
        return out

    def upload(self, device=0, stream=0):
        check(self._lib.td_upload(self._h, int(device), ctypes.c_void_p(stream)))
        return self

    def pairwise_device(self, metrics, d_out_ptrs, normalise=False, min_overlap=4, overlap_fail_value=0.0,
                        row_begin=0, row_end=None, stream=0):
        """Kernels only; d_out_ptrs maps metric name -> device pointer (int). Asynchronous on `stream`."""
        mask, table = 0, [None] * 4
        for m in metrics:
            mask |= 1 << METRIC_INDEX[m]
            table[METRIC_INDEX[m]] = d_out_ptrs[m]
        if row_end is None:
            row_end = self.n_trees
        check(self._lib.td_pairwise_device(self._h, mask, int(bool(normalise)), int(min_overlap),
                                           float(overlap_fail_value), int(row_begin), int(row_end),
                                           _c_void_p4(*table), ctypes.c_void_p(stream)))

    def pairwise_device_blocks(self, metrics, blocks, normalise=False, min_overlap=4, overlap_fail_value=0.0, stream=0):
        """Kernels only, for one or two disjoint row blocks [(row_begin, row_end, {metric: device pointer}), ...] - a rank's share under
        engine.folded_row_ranges.  Two blocks go through td_pairwise_device2 (one events kernel, one fix-up kernel for both)."""
        if len(blocks) == 1:
            a, b, ptrs = blocks[0]
            return self.pairwise_device(metrics, ptrs, normalise, min_overlap, overlap_fail_value, a, b, stream)
        assert len(blocks) == 2
        mask, tables = 0, [[None] * 4, [None] * 4]
        for m in metrics:
            mask |= 1 << METRIC_INDEX[m]
            for k in range(2):
                tables[k][METRIC_INDEX[m]] = blocks[k][2][m]
        i64x2 = ctypes.c_int64 * 2
        check(self._lib.td_pairwise_device2(self._h, mask, int(bool(normalise)), int(min_overlap), float(overlap_fail_value),
                                            i64x2(int(blocks[0][0]), int(blocks[1][0])), i64x2(int(blocks[0][1]), int(blocks[1][1])),
                                            _c_void_p4(*tables[0]), _c_void_p4(*tables[1]), ctypes.c_void_p(stream)))

    def pairwise_rect_device(self, split, metrics, d_out_ptrs, normalise=False, stream=0):
        mask, table = 0, [None] * 4
        for m in metrics:
            mask |= 1 << METRIC_INDEX[m]
            table[METRIC_INDEX[m]] = d_out_ptrs[m]
        check(self._lib.td_pairwise_rect_device(self._h, int(split), mask, int(bool(normalise)), _c_void_p4(*table),
                                                ctypes.c_void_p(stream)))

    def rf_incidence_device(self, d_out_ptr, out_is_u16=False, normalise=False, row_begin=0, row_end=None, stream=0):
        """Exact RF through the int8 tcgen05 split-incidence contraction; d_out_ptr: device pointer (u16 or f64 condensed)."""
        if row_end is None:
            row_end = self.n_trees
        check(self._lib.td_rf_incidence_device(self._h, int(bool(normalise)), int(row_begin), int(row_end),
                                               ctypes.c_void_p(d_out_ptr), int(bool(out_is_u16)), ctypes.c_void_p(stream)))

    def rf_incidence(self, row_begin=0, row_end=None, device=0, out=None):
        """Exact RF of rows [row_begin, row_end) as uint16, streamed in row bands into host memory (td_rf_incidence)."""
        if row_end is None:
            row_end = self.n_trees
        n = self.n_trees
        pairs = condensed_offset(n, row_end) - condensed_offset(n, row_begin)
        arr = out if out is not None else pinned_empty(pairs, np.uint16)
        assert arr.dtype == np.uint16 and arr.size >= pairs and arr.flags.c_contiguous
        check(self._lib.td_rf_incidence(self._h, int(row_begin), int(row_end), arr.ctypes.data, int(device)))
        return arr

    def pairwise(self, metrics, normalise=False, min_overlap=4, overlap_fail_value=0.0, row_begin=0, row_end=None,
                 device=0, out=None):
        """End to end with host buffers: returns {metric: condensed float64 ndarray} for rows [row_begin,row_end)."""
        if row_end is None:
            row_end = self.n_trees
        n = self.n_trees
        pairs = condensed_offset(n, row_end) - condensed_offset(n, row_begin)
        mask, table, res = 0, [None] * 4, {}
        for m in metrics:
            mask |= 1 << METRIC_INDEX[m]
            arr = out[m] if out is not None else result_empty(pairs)
            assert arr.dtype == np.float64 and arr.size >= pairs and arr.flags.c_contiguous
            res[m] = arr
            table[METRIC_INDEX[m]] = arr.ctypes.data
        check(self._lib.td_pairwise(self._h, mask, int(bool(normalise)), int(min_overlap), float(overlap_fail_value),
                                    int(row_begin), int(row_end), _c_void_p4(*table), int(device)))
        return res

    def pairwise_square(self, metric, normalise=False, min_overlap=4, overlap_fail_value=0.0, device=0, out=None):
        """One metric as the N x N symmetric matrix with a zero diagonal (td_pairwise_square): squareform() done on the device."""
        n = self.n_trees
        arr = out if out is not None else result_empty((n, n))
        assert arr.dtype == np.float64 and arr.shape == (n, n) and arr.flags.c_contiguous
        check(self._lib.td_pairwise_square(self._h, METRIC_INDEX[metric], int(bool(normalise)), int(min_overlap), float(overlap_fail_value),
                                           arr.ctypes.data, int(device)))
        return arr

    def cmds(self, metric, dimensions=3, normalise=False, min_overlap=4, overlap_fail_value=0.0, max_iter=0, tol=0.0, device=0):
        """Classical MDS embedding of the `metric` matrix, computed where the matrix is: distances, double-centring and the leading
        eigenpairs all stay on the GPU, only the N x dimensions coordinates come back (td_cmds; treeCl/distance_matrix.py:301-315).
        Returns (coords, eigenvalues, info)."""
        coords = np.empty((self.n_trees, int(dimensions)), dtype=np.float64)
        vals = np.empty(int(dimensions), dtype=np.float64)
        it, res = ctypes.c_int(), ctypes.c_double()
        check(self._lib.td_cmds(self._h, METRIC_INDEX[metric], int(bool(normalise)), int(min_overlap), float(overlap_fail_value), int(dimensions),
                                coords.ctypes.data, vals.ctypes.data, int(max_iter), float(tol), ctypes.byref(it), ctypes.byref(res), int(device)))
        return coords, vals, {'iterations': it.value, 'last_relative_change': res.value}

    def spectral_embedding(self, metric, dimensions=3, prune_k=0, logic='or', scale_k=0, normalise=False, device=0):
        """Spectral(dm, ...).spectral_embedding_(dimensions) with the distance matrix of `metric` computed, turned into an affinity matrix and
        decomposed on the GPU (td_spectral_embedding; treeCl/clustering.py:171-228,288-303).  Returns (coords with unit rows, eigenvalues, info)."""
        coords = np.empty((self.n_trees, int(dimensions)), dtype=np.float64)
        vals = np.empty(int(dimensions), dtype=np.float64)
        it, res = ctypes.c_int(), ctypes.c_double()
        check(self._lib.td_spectral_embedding(self._h, METRIC_INDEX[metric], int(bool(normalise)), int(dimensions), int(prune_k),
                                              int(logic in ('and', '&')), int(scale_k), coords.ctypes.data, vals.ctypes.data, 0, 0.0,
                                              ctypes.byref(it), ctypes.byref(res), int(device)))
        return coords, vals, {'iterations': it.value, 'last_relative_change': res.value}

    def check_errors(self, device=0):
        """After pairwise_device / pairwise_rect_device and a stream synchronise: raises if a launch flagged an error (mixed rooting,
        geodesic iteration cap); clears the flag."""
        check(self._lib.td_check_errors(self._h, int(device)))

    def geo_stats(self):
        st = td_geo_stats()
        check(self._lib.td_geo_get_stats(self._h, ctypes.byref(st)))
        return dict(pairs=st.pairs, maxflow_solves=st.maxflow_solves, ratios=st.ratios, augmentations=st.augmentations)


def newick_labels(newick):
    lib = load()
    need, rooted = ctypes.c_int64(), ctypes.c_int()
    check(lib.td_newick_labels(_b(newick), None, 0, ctypes.byref(need), ctypes.byref(rooted)))
    buf = ctypes.create_string_buffer(need.value + 1)
    check(lib.td_newick_labels(_b(newick), buf, need.value + 1, None, None))
    s = buf.value.decode()
    return (s.split('\n') if s else []), bool(rooted.value)


def newick_prune(newick, keep):
    lib = load()
    keep = [_b(k) for k in keep]
    arr = (ctypes.c_char_p * len(keep))(*keep)
    need = ctypes.c_int64()
    check(lib.td_newick_prune(_b(newick), arr, len(keep), None, 0, ctypes.byref(need)))
    buf = ctypes.create_string_buffer(need.value + 1)
    check(lib.td_newick_prune(_b(newick), arr, len(keep), buf, need.value + 1, None))
    return buf.value.decode()


def device_count():
    c = ctypes.c_int()
    load().td_device_count(ctypes.byref(c))
    return c.value


def pairwise_rect(query, reference, metrics, normalise=False, device=0):
    """{metric: (n_query, n_reference) ndarray} for two separately encoded TreeSets over the same taxa (td_pairwise_rect)."""
    lib = load()
    nq, nr = query.n_trees, reference.n_trees
    mask, table, res = 0, [None] * 4, {}
    for m in metrics:
        mask |= 1 << METRIC_INDEX[m]
        res[m] = result_empty((nq, nr))
        table[METRIC_INDEX[m]] = res[m].ctypes.data
    check(lib.td_pairwise_rect(query._h, reference._h, mask, int(bool(normalise)), _c_void_p4(*table), int(device)))
    return res


def cmds_host(square, dimensions=3, max_iter=0, tol=0.0, device=0):
    """Classical MDS of a square symmetric distance matrix held in host memory (td_cmds_host).  Returns (coords, eigenvalues, info)."""
    lib = load()
    sq = np.ascontiguousarray(square, dtype=np.float64)
    assert sq.ndim == 2 and sq.shape[0] == sq.shape[1]
    n = sq.shape[0]
    coords = np.empty((n, int(dimensions)), dtype=np.float64)
    vals = np.empty(int(dimensions), dtype=np.float64)
    it, res = ctypes.c_int(), ctypes.c_double()
    check(lib.td_cmds_host(sq.ctypes.data, n, int(dimensions), coords.ctypes.data, vals.ctypes.data, int(max_iter), float(tol),
                           ctypes.byref(it), ctypes.byref(res), int(device)))
    return coords, vals, {'iterations': it.value, 'last_relative_change': res.value}


def trim_cache(device=-1):
    check(load().td_trim_cache(int(device)))


def launch_count():
    return int(load().td_launch_count())
