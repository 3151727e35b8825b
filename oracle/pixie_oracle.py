"""ctypes loader for oracle/libpixie_oracle.so (the C restatement) plus the small host-side pieces
of the pyFlowSOM semantics that live in Python there: the default radius
range and the seeded codebook initialisation.  TEST INFRASTRUCTURE ONLY."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libpixie_oracle.so")
_lib = None
TILE = 128  # mini-batch interleave unit of the batch SOM (must match PIXIE_TILE in the C file)

_f64p = ctypes.POINTER(ctypes.c_double)
_f32p = ctypes.POINTER(ctypes.c_float)
_i32p = ctypes.POINTER(ctypes.c_int32)


def build(force=False):
    """Compile the oracle with gcc (see oracle/Makefile for the flags)."""
    src = os.path.join(_HERE, "pixie_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or \
            os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "libpixie_oracle.so"])
    return _LIB_PATH


def _load():
    global _lib
    if _lib is None:
        build()
        lib = ctypes.CDLL(_LIB_PATH)
        lib.oracle_map_data_to_nodes.argtypes = [
            _f64p, ctypes.c_int, _f64p, ctypes.c_int64, ctypes.c_int, _i32p, _f64p]
        lib.oracle_map_data_to_nodes_f32.argtypes = [
            _f32p, ctypes.c_int, _f32p, ctypes.c_int64, ctypes.c_int, ctypes.c_int64, _i32p, _f64p]
        lib.oracle_map_data_to_nodes_mt.argtypes = [
            _f64p, ctypes.c_int, _f64p, ctypes.c_int64, ctypes.c_int, _i32p, _f64p, ctypes.c_int]
        lib.oracle_grid_chebyshev.argtypes = [ctypes.c_int, ctypes.c_int, _f64p]
        lib.oracle_som_online.argtypes = [
            _f64p, ctypes.c_int64, ctypes.c_int, _f64p, ctypes.c_int, _f64p, ctypes.c_double,
            ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_int, ctypes.c_uint]
        lib.oracle_som_online.restype = ctypes.c_int64
        lib.oracle_som_batch.argtypes = [
            _f32p, ctypes.c_int64, ctypes.c_int, ctypes.c_int64, _f64p, ctypes.c_int, ctypes.c_int,
            ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_double,
            ctypes.c_double]
        lib.oracle_cluster_sums_f32.argtypes = [
            _f32p, ctypes.c_int64, ctypes.c_int, ctypes.c_int64, _i32p, ctypes.c_int, _f64p, _f64p]
        _lib = lib
    return _lib


def _p(a, t):
    return a.ctypes.data_as(t)


def grid_chebyshev(xdim, ydim):
    K = xdim * ydim
    D = np.empty((K, K), np.float64)
    _load().oracle_grid_chebyshev(xdim, ydim, _p(D, _f64p))
    return D


def default_radius(xdim, ydim):
    """pyFlowSOM default radius_range = (quantile(nhbrdist, 0.67), 0) (Appendix A)."""
    return float(np.quantile(grid_chebyshev(xdim, ydim), 0.67)), 0.0


def init_codebook_indices(n, K, seed):
    """Seeded choice of K distinct data rows for the initial codebook: pyFlowSOM seeds numpy's
    LEGACY global RNG with ``seed`` and draws ``np.random.choice(n, K, replace=False)`` (SURVEY.md
    Appendix A; ark's comment at pixel_som_clustering.py:86 says the seed is consumed there).
    ``RandomState(seed)`` is the same generator without touching the global state.  Evidence beyond
    the recollection: the reference's own test_train_cell_som (cell_som_clustering_test.py:157,
    `assert not np.all(cell_weights < 1)`, data seeded 24 by pytest-randomly) holds for this init
    and FAILS for ``default_rng(seed).choice`` (tests/test_reference_suite.py)."""
    return np.random.RandomState(seed).choice(n, K, replace=False)


def map_data_to_nodes(nodes, newdata):
    """pyFlowSOM.map_data_to_nodes semantics: (labels 1..K int32, dists f64)."""
    nodes = np.ascontiguousarray(nodes, np.float64)
    newdata = np.ascontiguousarray(newdata, np.float64)
    m, C = newdata.shape if newdata.ndim == 2 else (0, nodes.shape[1])
    labels = np.empty(m, np.int32)
    dists = np.empty(m, np.float64)
    if m:
        _load().oracle_map_data_to_nodes(_p(nodes, _f64p), nodes.shape[0], _p(newdata, _f64p), m,
                                         C, _p(labels, _i32p), _p(dists, _f64p))
    return labels, dists


def map_data_to_nodes_mt(nodes, newdata, nthreads):
    nodes = np.ascontiguousarray(nodes, np.float64)
    newdata = np.ascontiguousarray(newdata, np.float64)
    m, C = newdata.shape
    labels = np.empty(m, np.int32)
    dists = np.empty(m, np.float64)
    _load().oracle_map_data_to_nodes_mt(_p(nodes, _f64p), nodes.shape[0], _p(newdata, _f64p), m, C,
                                        _p(labels, _i32p), _p(dists, _f64p), int(nthreads))
    return labels, dists


def map_data_to_nodes_f32(nodes32, data32):
    """Same arithmetic on fp32 inputs promoted to fp64 (the parity protocol)."""
    nodes32 = np.ascontiguousarray(nodes32, np.float32)
    assert data32.dtype == np.float32 and data32.ndim == 2
    if data32.shape[0] > 1 and data32.strides[1] != 4:
        data32 = np.ascontiguousarray(data32)
    m, C = data32.shape
    ld = data32.strides[0] // 4 if m > 1 else C
    labels = np.empty(m, np.int32)
    dists = np.empty(m, np.float64)
    if m:
        _load().oracle_map_data_to_nodes_f32(_p(nodes32, _f32p), nodes32.shape[0],
                                             _p(data32, _f32p), m, C, ld, _p(labels, _i32p),
                                             _p(dists, _f64p))
    return labels, dists


def som_online(data, xdim=10, ydim=10, rlen=10, alpha_range=(0.05, 0.01), radius_range=None,
               seed=42, init_idx=None):
    """pyFlowSOM.som semantics (online SOM, Appendix A -- UNVERIFIED).  Returns (K, C) f64."""
    data = np.ascontiguousarray(data, np.float64)
    n, C = data.shape
    K = xdim * ydim
    if radius_range is None:
        radius_range = default_radius(xdim, ydim)
    if init_idx is None:
        init_idx = init_codebook_indices(n, K, seed)
    nodes = np.ascontiguousarray(data[init_idx], np.float64).copy()
    D = grid_chebyshev(xdim, ydim)
    _load().oracle_som_online(_p(data, _f64p), n, C, _p(nodes, _f64p), K, _p(D, _f64p),
                              alpha_range[0], alpha_range[1], radius_range[0], radius_range[1],
                              int(rlen), int(seed) & 0xFFFFFFFF)
    return nodes


def som_batch(data32, xdim=10, ydim=10, rlen=1, alpha_range=(0.05, 0.01), radius_range=None,
              seed=42, batches_per_pass=None, init_idx=None):
    """fp64 restatement of the batch SOM the B200 path runs."""
    assert data32.dtype == np.float32 and data32.ndim == 2
    if data32.shape[0] > 1 and data32.strides[1] != 4:
        data32 = np.ascontiguousarray(data32)
    n, C = data32.shape
    ld = data32.strides[0] // 4 if n > 1 else C
    K = xdim * ydim
    if radius_range is None:
        radius_range = default_radius(xdim, ydim)
    if init_idx is None:
        init_idx = init_codebook_indices(n, K, seed)
    ntiles = (n + TILE - 1) // TILE
    B = batches_per_pass if batches_per_pass else max(1, min(32, ntiles))
    W = np.ascontiguousarray(data32[init_idx], np.float64).copy()
    _load().oracle_som_batch(_p(data32, _f32p), n, C, ld, _p(W, _f64p), xdim, ydim, int(rlen),
                             int(B), alpha_range[0], alpha_range[1], radius_range[0],
                             radius_range[1])
    return W


def cluster_sums_f32(data32, labels, K):
    assert data32.dtype == np.float32 and data32.ndim == 2
    if data32.shape[0] > 1 and data32.strides[1] != 4:
        data32 = np.ascontiguousarray(data32)
    n, C = data32.shape
    ld = data32.strides[0] // 4 if n > 1 else C
    labels = np.ascontiguousarray(labels, np.int32)
    S = np.empty((K, C), np.float64)
    cnt = np.empty(K, np.float64)
    _load().oracle_cluster_sums_f32(_p(data32, _f32p), n, C, ld, _p(labels, _i32p), K,
                                    _p(S, _f64p), _p(cnt, _f64p))
    return S, cnt
