"""CPU tests of the oracle (oracle/pixie_oracle.c) -- the checker every GPU parity test relies on.

The oracle restates pyFlowSOM's map_data_to_nodes / som (call sites
/root/reference/src/ark/phenotyping/cluster_helpers.py:106-109, :152-157).
It is checked here against (i) an independent plain-numpy restatement, (ii) the invariants the
reference's own tests pin for this path (labels in 1..K, same-seed determinism, shapes; reference
tests/phenotyping/cluster_helpers_test.py:304-404) and (iii) the committed golden vectors.
"""
import os

import numpy as np
import pytest

import oracle
from conftest import pixie_like

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def numpy_bmu(W, X):
    """argmin_k sqrt(sum_j (x - w)^2), sequential accumulation over j, first minimum wins."""
    W = W.astype(np.float64)
    X = X.astype(np.float64)
    acc = np.zeros((X.shape[0], W.shape[0]))
    for j in range(X.shape[1]):  # same summation order as the C loop
        t = X[:, j:j + 1] - W[None, :, j]
        acc = acc + t * t
    d = np.sqrt(acc)
    labels = np.zeros(X.shape[0], np.int32)
    dist = np.full(X.shape[0], np.finfo(np.float64).max)
    for k in range(W.shape[0]):  # strict '<' in node order; NaN never wins
        better = d[:, k] < dist
        dist = np.where(better, d[:, k], dist)
        labels = np.where(better, k + 1, labels)
    return labels, dist


@pytest.mark.parametrize("n,C,K", [(500, 16, 100), (300, 7, 30), (128, 32, 200)])
def test_map_data_to_nodes_matches_numpy(rng, n, C, K):
    X = rng.random((n, C))
    W = rng.random((K, C))
    labels, dists = oracle.map_data_to_nodes(W, X)
    ref_l, ref_d = numpy_bmu(W, X)
    assert labels.dtype == np.int32
    np.testing.assert_array_equal(labels, ref_l)
    np.testing.assert_array_equal(dists, ref_d)  # bit-exact: same operation order
    assert labels.min() >= 1 and labels.max() <= K  # reference cluster_helpers_test.py:388-391


def test_map_data_to_nodes_edge_cases():
    W = np.array([[0.0, 0.0], [1.0, 1.0], [0.0, 0.0]])
    X = np.array([[0.0, 0.0], [1.0, 1.0], [0.5, 0.5], [np.nan, 0.0]])
    labels, dists = oracle.map_data_to_nodes(W, X)
    # duplicate node: the lower index wins; equidistant: first minimum wins; NaN row: label 0
    np.testing.assert_array_equal(labels, [1, 2, 1, 0])
    assert dists[3] == np.finfo(np.float64).max
    labels0, dists0 = oracle.map_data_to_nodes(W, np.empty((0, 2)))
    assert labels0.shape == (0,) and dists0.shape == (0,)


def test_f32_entry_equals_f64_entry_on_same_values(rng):
    X = rng.random((400, 16)).astype(np.float32)
    W = rng.random((50, 16)).astype(np.float32)
    a, da = oracle.map_data_to_nodes_f32(W, X)
    b, db = oracle.map_data_to_nodes(W.astype(np.float64), X.astype(np.float64))
    np.testing.assert_array_equal(a, b)
    np.testing.assert_array_equal(da, db)
    c, _ = oracle.map_data_to_nodes_mt(W.astype(np.float64), X.astype(np.float64), 3)
    np.testing.assert_array_equal(a, c)


@pytest.mark.parametrize("name", ["bmu_u16_k100", "bmu_p32_k100", "bmu_ties_c8_k40",
                                  "bmu_nan_c15_k49"])
def test_golden_bmu(name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    labels, dists = oracle.map_data_to_nodes_f32(g["W"], g["X"])
    np.testing.assert_array_equal(labels, g["labels"])
    np.testing.assert_array_equal(dists, g["dists"])
    ref_l, _ = numpy_bmu(g["W"], g["X"])
    np.testing.assert_array_equal(labels, ref_l)


def test_grid_and_radius():
    D = oracle.grid_chebyshev(3, 2)
    assert D.shape == (6, 6) and D[0, 5] == 2 and D[0, 1] == 1 and D[1, 2] == 1
    # pyFlowSOM default radius = 0.67 quantile of the grid distances
    assert oracle.default_radius(10, 10) == (6.0, 0.0)
    assert oracle.default_radius(20, 20) == (11.0, 0.0)
    assert oracle.default_radius(20, 10) == (9.0, 0.0)


def numpy_som_batch(X32, W0, xdim, ydim, rlen, B, a, r):
    """Independent restatement of the batch SOM."""
    K, C = W0.shape
    n = X32.shape[0]
    D = oracle.grid_chebyshev(xdim, ydim)
    W = W0.astype(np.float64).copy()
    T = rlen * B
    tile = np.arange(n) // 128
    X = X32.astype(np.float64)
    for t in range(T):
        m = t % B
        rows = np.nonzero(tile % B == m)[0]
        W32 = W.astype(np.float32)
        lab, _ = numpy_bmu(W32, X32[rows])
        S = np.zeros((K, C))
        cnt = np.zeros(K)
        for i, b in zip(rows, lab - 1):
            S[b] += X[i]
            cnt[b] += 1
        frac = t / T
        rad = r[0] - (r[0] - r[1]) * frac
        sigma = 0.5 * (0.5 if rad < 1.0 else rad)
        alpha = a[0] - (a[0] - a[1]) * frac
        H = np.exp(-D * D / (2 * sigma * sigma))
        den = H @ cnt
        num = H @ S
        upd = den > 0
        beta = 1.0 - np.power(1.0 - alpha, den[upd])
        W[upd] += beta[:, None] * (num[upd] / den[upd, None] - W[upd])
    return W


def test_som_batch_matches_numpy():
    X = pixie_like(1500, 8, seed=3)
    idx = oracle.init_codebook_indices(1500, 12, 7)
    W = oracle.som_batch(X, 4, 3, rlen=2, batches_per_pass=5, init_idx=idx)
    ref = numpy_som_batch(X, X[idx], 4, 3, 2, 5, (0.05, 0.01), oracle.default_radius(4, 3))
    np.testing.assert_allclose(W, ref, rtol=1e-10, atol=1e-12)


def test_golden_som_batch_and_determinism():
    g = np.load(os.path.join(GOLD, "som_batch_p16_6x5.npz"))
    W = oracle.som_batch(g["X"], 6, 5, rlen=2, seed=42, init_idx=g["init_idx"])
    np.testing.assert_allclose(W, g["W"], rtol=1e-12, atol=0)
    # same seed -> same weights (reference cluster_helpers_test.py:323-332)
    W2 = oracle.som_batch(g["X"], 6, 5, rlen=2, seed=42)
    np.testing.assert_array_equal(oracle.init_codebook_indices(3000, 30, 42), g["init_idx"])
    np.testing.assert_allclose(W2, W, rtol=0, atol=0)
    assert W.shape == (30, 16)


def test_som_online_restatement():
    g = np.load(os.path.join(GOLD, "som_batch_p16_6x5.npz"))
    go = np.load(os.path.join(GOLD, "som_online_p16_6x5.npz"))
    X = g["X"].astype(np.float64)
    W = oracle.som_online(X, 6, 5, rlen=1, seed=42, init_idx=g["init_idx"])
    np.testing.assert_allclose(W, go["W"], rtol=1e-12, atol=0)
    # convex-combination updates keep weights inside the data range
    # (reference cell_som_clustering_test.py:97: trained weights < 1 on <= 1 data)
    assert W.min() >= X.min() - 1e-12 and W.max() <= X.max() + 1e-12
    # map quality: both trainers must beat the untrained codebook by a clear margin
    def qe(Wc):
        return oracle.map_data_to_nodes(Wc, X)[1].mean()
    init = X[g["init_idx"]]
    assert qe(W) < 0.95 * qe(init)
    assert qe(g["W"]) < 0.95 * qe(init)


def test_cluster_sums(rng):
    X = rng.random((1000, 5)).astype(np.float32)
    labels = rng.integers(0, 8, 1000).astype(np.int32)  # 0 = unlabelled, skipped
    S, cnt = oracle.cluster_sums_f32(X, labels, 7)
    for k in range(7):
        sel = labels == k + 1
        np.testing.assert_allclose(S[k], X[sel].astype(np.float64).sum(0), rtol=1e-12)
        assert cnt[k] == sel.sum()
