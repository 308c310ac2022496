"""ctypes binding of oracle/treedist_oracle.c (plain-C restatement of the reference path; see its header).

Parity status: PINNED against the reference's golden vectors for geo/euc/wrf/rf on 10-taxon unrooted binary
trees and for geodesic partial overlap (tests/test_oracle_golden.py); UNPINNED for rooted trees, polytomies,
zero/missing lengths, and rf/wrf/euc under partial overlap (no reference test covers them); UNPINNED for the
Kendall-Colijn metric (kc_*): the reference has no test for it and dendropy is not installed here, the restatement
is checked against hand-derived values of the published definition only (tests/test_kendall_colijn.py).
"""
import ctypes
import os
import subprocess

import numpy as np

__all__ = ['build', 'pair', 'matrix', 'matrix_range', 'matrix_rows', 'prune_newick', 'encode_summary', 'stats', 'stats_reset', 'kc_pair', 'kc_vector', 'kc_matrix',
           'METRICS', 'OracleError']

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, 'libtreedist_oracle.so')
METRICS = {'rf': 0, 'wrf': 1, 'euc': 2, 'geo': 3}
_lib = None


class OracleError(ValueError):
    pass


def build(force=False):
    src = os.path.join(_HERE, 'treedist_oracle.c')
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(['make', '-s', '-C', _HERE, 'libtreedist_oracle.so'] + (['-B'] if force else []))
    return _SO


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        lib = ctypes.CDLL(_SO)
        lib.orc_last_error.restype = ctypes.c_char_p
        lib.orc_pair.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                 ctypes.c_double, ctypes.POINTER(ctypes.c_double)]
        lib.orc_matrix_range.argtypes = [ctypes.POINTER(ctypes.c_char_p), ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                         ctypes.c_int, ctypes.c_double, ctypes.c_int, ctypes.c_int,
                                         ctypes.c_long, ctypes.c_long, ctypes.c_void_p]
        lib.orc_matrix_rows.argtypes = [ctypes.POINTER(ctypes.c_char_p), ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                        ctypes.c_double, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
        lib.orc_prune_newick.argtypes = [ctypes.c_char_p, ctypes.POINTER(ctypes.c_char_p), ctypes.c_int,
                                         ctypes.c_char_p, ctypes.c_size_t]
        lib.orc_encode_summary.argtypes = [ctypes.c_char_p] + [ctypes.POINTER(ctypes.c_int)] * 3 + \
                                          [ctypes.POINTER(ctypes.c_double)] * 2
        lib.orc_stats.argtypes = [ctypes.POINTER(ctypes.c_long)] * 3
        lib.orc_kc_pair.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_double, ctypes.c_int, ctypes.POINTER(ctypes.c_double)]
        lib.orc_kc_vector.argtypes = [ctypes.c_char_p, ctypes.c_double, ctypes.c_void_p, ctypes.c_long, ctypes.POINTER(ctypes.c_long)]
        lib.orc_kc_matrix.argtypes = [ctypes.POINTER(ctypes.c_char_p), ctypes.c_int, ctypes.c_double, ctypes.c_int, ctypes.c_void_p]
        _lib = lib
    return _lib


def _check(rc):
    if rc:
        raise OracleError(_load().orc_last_error().decode())


def _b(s):
    return s.encode() if isinstance(s, str) else s


def pair(nwk_a, nwk_b, metric, normalise=False, min_overlap=4, overlap_fail_value=0.0):
    """Reference-shaped pair distance: two Newick strings in, one float out (treeCl/tasks.py:41-75)."""
    out = ctypes.c_double()
    _check(_load().orc_pair(_b(nwk_a), _b(nwk_b), METRICS[metric], int(bool(normalise)), int(min_overlap),
                            float(overlap_fail_value), ctypes.byref(out)))
    return out.value


def matrix_range(newicks, metric, pair_begin, pair_end, normalise=False, min_overlap=4, overlap_fail_value=0.0,
                 nthreads=1, reparse_per_pair=False):
    n = len(newicks)
    arr = (ctypes.c_char_p * n)(*[_b(s) for s in newicks])
    out = np.zeros(max(pair_end - pair_begin, 0), dtype=np.float64)
    _check(_load().orc_matrix_range(arr, n, METRICS[metric], int(bool(normalise)), int(min_overlap),
                                    float(overlap_fail_value), int(nthreads), int(bool(reparse_per_pair)),
                                    int(pair_begin), int(pair_end), out.ctypes.data))
    return out


def matrix(newicks, metric, normalise=False, min_overlap=4, overlap_fail_value=0.0, nthreads=1,
           reparse_per_pair=False):
    """Condensed all-pairs vector in scipy order (treeCl/tasks.py:651-653, collection.py:455-456)."""
    n = len(newicks)
    return matrix_range(newicks, metric, 0, n * (n - 1) // 2, normalise, min_overlap, overlap_fail_value,
                        nthreads, reparse_per_pair)


def matrix_rows(newicks, metric, rows, normalise=False, min_overlap=4, overlap_fail_value=0.0, nthreads=1):
    """Rows `rows` of the SQUARE matrix, shape (len(rows), n): spot checks of matrices too large for matrix()."""
    n = len(newicks)
    arr = (ctypes.c_char_p * n)(*[_b(s) for s in newicks])
    rows = np.ascontiguousarray(rows, dtype=np.int64)
    out = np.zeros((rows.size, n), dtype=np.float64)
    _check(_load().orc_matrix_rows(arr, n, METRICS[metric], int(bool(normalise)), int(min_overlap), float(overlap_fail_value),
                                   int(nthreads), rows.ctypes.data, int(rows.size), out.ctypes.data))
    return out


def prune_newick(nwk, keep):
    """Tree.prune_to_subset(...).newick restated (treeCl/tree.py:989-1001)."""
    keep = list(keep)
    arr = (ctypes.c_char_p * len(keep))(*[_b(k) for k in keep])
    buf = ctypes.create_string_buffer(4 * len(_b(nwk)) + 64 * len(keep) + 256)
    _check(_load().orc_prune_newick(_b(nwk), arr, len(keep), buf, len(buf)))
    return buf.value.decode()


def encode_summary(nwk):
    nl, ns, rooted = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    si, sl = ctypes.c_double(), ctypes.c_double()
    _check(_load().orc_encode_summary(_b(nwk), ctypes.byref(nl), ctypes.byref(ns), ctypes.byref(rooted),
                                      ctypes.byref(si), ctypes.byref(sl)))
    return dict(n_leaves=nl.value, n_splits=ns.value, rooted=bool(rooted.value), sum_internal=si.value,
                sum_leaf=sl.value)


def stats():
    a, b, c = ctypes.c_long(), ctypes.c_long(), ctypes.c_long()
    _load().orc_stats(ctypes.byref(a), ctypes.byref(b), ctypes.byref(c))
    return dict(maxflows=a.value, ratios=b.value, pairs=c.value)


def stats_reset():
    _load().orc_stats_reset()


def kc_pair(nwk_a, nwk_b, lbda=0.5, min_overlap=4):
    """KendallColijn(t1).get_distance(KendallColijn(t2), lbda, min_overlap) restated (treeCl/utils/kendallcolijn.py:81-95)."""
    out = ctypes.c_double()
    _check(_load().orc_kc_pair(_b(nwk_a), _b(nwk_b), float(lbda), int(min_overlap), ctypes.byref(out)))
    return out.value


def kc_vector(nwk, lbda=0.5):
    """KendallColijn(tree).get_vector(lbda) restated (treeCl/utils/kendallcolijn.py:53-79)."""
    need = ctypes.c_long()
    _check(_load().orc_kc_vector(_b(nwk), float(lbda), None, 0, ctypes.byref(need)))
    out = np.zeros(need.value, dtype=np.float64)
    _check(_load().orc_kc_vector(_b(nwk), float(lbda), out.ctypes.data, need.value, None))
    return out


def kc_matrix(newicks, lbda=0.5, min_overlap=4):
    n = len(newicks)
    arr = (ctypes.c_char_p * n)(*[_b(s) for s in newicks])
    out = np.zeros(n * (n - 1) // 2, dtype=np.float64)
    _check(_load().orc_kc_matrix(arr, n, float(lbda), int(min_overlap), out.ctypes.data))
    return out
