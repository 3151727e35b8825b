"""GPU parity tests of SOM training: the batch SOM kernels (accumulate + apply) through the C ABI
against the fp64 restatement of the same algorithm (oracle.som_batch).

Tolerance (BASELINE.json north_star): trained codebook weights within 1e-4 RELATIVE of the oracle
on identical seeds and inputs.  Measured differences are ~1e-8 (fp32 partial sums vs fp64)."""
import os

import numpy as np
import pytest
import torch

import oracle
from ark_analysis_b200 import som as S
from conftest import pixie_like

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
RTOL = 1e-4  # north_star tolerance


def rel_err(W, ref):
    return float(np.abs(W - ref).max() / np.abs(ref).max())


@pytest.mark.parametrize("kind,n,C,xd,yd,rlen", [
    ("U", 20000, 16, 10, 10, 1), ("P", 50000, 32, 10, 10, 1), ("P", 30000, 40, 20, 20, 1),
    ("P", 6554, 16, 10, 10, 2),   # cfg1: 10 % of one 256 x 256 FOV, 16 channels
    ("U", 1000, 6, 20, 10, 1),    # the reference tests' shape (cluster_helpers_test.py:304-317)
    ("P", 5000, 15, 7, 3, 3)])
def test_train_matches_oracle(kind, n, C, xd, yd, rlen):
    X = np.random.default_rng(0).random((n, C), dtype=np.float32) if kind == "U" \
        else pixie_like(n, C, seed=1)
    idx = oracle.init_codebook_indices(n, xd * yd, 42)
    ref = oracle.som_batch(X, xd, yd, rlen=rlen, init_idx=idx)
    Xd = S.to_device_matrix(X)
    W = S.train_som(Xd, X[idx].astype(np.float64), xd, yd, rlen=rlen).cpu().numpy()
    assert W.shape == (xd * yd, C) and W.dtype == np.float64
    assert rel_err(W, ref) < RTOL
    # same seed -> same weights, bit for bit (reference cluster_helpers_test.py:323-332)
    W2 = S.train_som(Xd, X[idx].astype(np.float64), xd, yd, rlen=rlen).cpu().numpy()
    np.testing.assert_array_equal(W, W2)


def test_golden_som_batch():
    g = np.load(os.path.join(GOLD, "som_batch_p16_6x5.npz"))
    Xd = S.to_device_matrix(g["X"])
    W = S.train_som(Xd, g["X"][g["init_idx"]].astype(np.float64), 6, 5, rlen=2).cpu().numpy()
    assert rel_err(W, g["W"]) < RTOL


def test_pyflowsom_shaped_som_function():
    X = pixie_like(8000, 12, seed=4).astype(np.float64)
    W = S.som(X, xdim=5, ydim=4, rlen=2, alpha_range=(0.05, 0.01), seed=42, algorithm="batch")
    assert W.shape == (20, 12) and W.dtype == np.float64
    ref = oracle.som_batch(X.astype(np.float32), 5, 4, rlen=2, seed=42)
    assert rel_err(W, ref) < RTOL
    # convex-combination updates stay inside the data range (reference
    # cell_som_clustering_test.py:97: trained weights < 1 on <= 1 data)
    assert W.min() >= X.min() - 1e-6 and W.max() <= X.max() + 1e-6
    with pytest.raises(ValueError):
        S.som(X[:10], xdim=5, ydim=4, rlen=1, seed=1)  # fewer rows than nodes (any algorithm)


def test_step_functions_accumulate_and_apply():
    """pixie_som_accum_f32 + pixie_som_apply_f64 driven step by step == the fused train call, and
    the per-step statistics equal the oracle's for that mini-batch."""
    n, C, xd, yd, B = 128 * 40 + 11, 16, 6, 6, 5
    X = pixie_like(n, C, seed=2)
    K = xd * yd
    idx = oracle.init_codebook_indices(n, K, 1)
    Xd = S.to_device_matrix(X)
    W64 = torch.from_numpy(X[idx].astype(np.float64)).cuda()
    W32 = torch.empty((K, C), dtype=torch.float32, device="cuda")
    SN = torch.zeros((K, C + 1), dtype=torch.float64, device="cuda")
    S.som_apply(W64, W32, SN, xd, yd, 1.0, 0.0)
    np.testing.assert_array_equal(W32.cpu().numpy(), X[idx])
    rr = S.default_radius(xd, yd)
    for t in range(B):
        S.som_accum(Xd, W32, t % B, B, SN=SN)
        # oracle statistics of the same mini-batch
        rows = np.nonzero((np.arange(n) // 128) % B == t % B)[0]
        lab, _ = oracle.map_data_to_nodes_f32(W32.cpu().numpy(), X[rows])
        Sref, cref = oracle.cluster_sums_f32(X[rows], lab, K)
        got = SN.cpu().numpy()
        np.testing.assert_array_equal(got[:, C], cref)
        np.testing.assert_allclose(got[:, :C], Sref, rtol=1e-5, atol=1e-6)
        sigma, alpha = S.step_schedule(t, B, (0.05, 0.01), rr)
        S.som_apply(W64, W32, SN, xd, yd, sigma, alpha)
    fused = S.train_som(Xd, X[idx].astype(np.float64), xd, yd, rlen=1, batches_per_pass=B)
    np.testing.assert_array_equal(W64.cpu().numpy(), fused.cpu().numpy())
    ref = oracle.som_batch(X, xd, yd, rlen=1, batches_per_pass=B, init_idx=idx)
    assert rel_err(fused.cpu().numpy(), ref) < RTOL


def test_map_quality_is_as_good_as_the_online_reference_rule():
    """The reference trains an ONLINE SOM; ours is a batch SOM (SURVEY.md section 7 hard part 1).
    Weight parity across the two is meaningless; map quality is what must hold."""
    X = pixie_like(30000, 16, seed=8)
    Wb = S.som(X, xdim=10, ydim=10, rlen=1, seed=42, algorithm="batch")
    Wo = oracle.som_online(X.astype(np.float64), 10, 10, rlen=1, seed=42)

    def qe(W):
        return oracle.map_data_to_nodes(W, X.astype(np.float64))[1].mean()
    print(f"quantisation error: batch(GPU) {qe(Wb):.5f}  online(oracle) {qe(Wo):.5f}")
    assert qe(Wb) < 1.05 * qe(Wo)


# ------------------------------------------------------------------------------------------------
# parity mode: the reference's own online rule on the device (pixie_som_online_f64)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,C,xd,yd,rlen", [(3000, 16, 10, 10, 1), (2000, 7, 5, 4, 3),
                                            (1500, 40, 20, 20, 1), (700, 100, 10, 10, 2),
                                            (64, 3, 2, 2, 5)])
def test_online_som_is_bit_identical_to_the_reference_rule(n, C, xd, yd, rlen):
    """Same fp32-representable inputs, same seed: the device codebook equals the C restatement of
    pyFlowSOM's C_SOM bit for bit (libc rand() sample stream, fp64 operation order, shrinking
    radius accumulated step by step, early stop between passes)."""
    rng = np.random.default_rng(n + C)
    X = rng.random((n, C)).astype(np.float32)
    idx = oracle.init_codebook_indices(n, xd * yd, 7)
    want = oracle.som_online(X.astype(np.float64), xd, yd, rlen=rlen, seed=7, init_idx=idx)
    Xd = S.to_device_matrix(X)
    got, iters = S.train_som_online(Xd, X[idx].astype(np.float64), xd, yd, rlen=rlen, seed=7)
    assert np.array_equal(got.cpu().numpy(), want)
    assert 1 <= iters <= rlen * n


def test_online_som_through_the_pyflowsom_shaped_call():
    rng = np.random.default_rng(3)
    X = rng.random((1200, 8)).astype(np.float32)
    W = S.som(X, xdim=4, ydim=5, rlen=2, seed=11, algorithm="online")
    idx = oracle.init_codebook_indices(1200, 20, 11)
    want = oracle.som_online(X.astype(np.float64), 4, 5, rlen=2, seed=11, init_idx=idx)
    assert W.shape == (20, 8) and np.array_equal(W, want)
    # "auto" (the default) takes the online rule for a table this small and the batch SOM beyond
    # ONLINE_MAX_ITERS samples
    assert np.array_equal(S.som(X, xdim=4, ydim=5, rlen=2, seed=11), want)
    big = np.tile(X, (60, 1))  # 72,000 rows > 65,536
    auto = S.som(big, xdim=4, ydim=5, rlen=1, seed=11)
    assert np.array_equal(auto, S.som(big, xdim=4, ydim=5, rlen=1, seed=11, algorithm="batch"))
    with pytest.raises(ValueError):
        S.som(X, xdim=4, ydim=5, algorithm="nope")


# ------------------------------------------------------------------------------------------------
# the BASELINE shapes that used to fall off the whole-pass kernel (round 1: step-by-step path)
# ------------------------------------------------------------------------------------------------
def test_train_matches_oracle_cfg4_shape():
    """cfg4's shape: 100 features, 10 x 10 SOM (group tables in global memory, red.v4 accumulate)."""
    n, C, xd, yd = 128 * 400 + 37, 100, 10, 10
    X = pixie_like(n, C, seed=5)
    idx = oracle.init_codebook_indices(n, xd * yd, 42)
    ref = oracle.som_batch(X, xd, yd, rlen=1, init_idx=idx)
    Xd = S.to_device_matrix(X)
    W = S.train_som(Xd, X[idx].astype(np.float64), xd, yd, rlen=1).cpu().numpy()
    assert rel_err(W, ref) < RTOL
    W2 = S.train_som(Xd, X[idx].astype(np.float64), xd, yd, rlen=1).cpu().numpy()
    np.testing.assert_array_equal(W, W2)  # red.global.add in a fixed order: still deterministic


def test_train_matches_oracle_cfg3_shape_global_tables_forced_shared_shape():
    """40 channels, 20 x 20 SOM: every node class, both accumulate forms (shared-memory tables are
    forced for a shape that fits them through PIXIE_TAB_GLOBAL in test_table_kinds below)."""
    n, C, xd, yd = 128 * 300 + 5, 40, 20, 20
    X = pixie_like(n, C, seed=6)
    idx = oracle.init_codebook_indices(n, xd * yd, 42)
    ref = oracle.som_batch(X, xd, yd, rlen=2, init_idx=idx)
    W = S.train_som(S.to_device_matrix(X), X[idx].astype(np.float64), xd, yd, rlen=2).cpu().numpy()
    assert rel_err(W, ref) < RTOL


@pytest.mark.parametrize("tab_global", ["0", "1"])
def test_table_kinds_agree(tab_global, monkeypatch):
    """Shared-memory and global group tables (PIXIE_TAB_GLOBAL forces either for a shape that fits
    both) give the oracle's codebook; the two differ only in fp32 summation order."""
    monkeypatch.setenv("PIXIE_TAB_GLOBAL", tab_global)
    n, C, xd, yd = 128 * 200 + 64, 24, 8, 8
    X = pixie_like(n, C, seed=9)
    idx = oracle.init_codebook_indices(n, xd * yd, 3)
    ref = oracle.som_batch(X, xd, yd, rlen=1, init_idx=idx)
    W = S.train_som(S.to_device_matrix(X), X[idx].astype(np.float64), xd, yd, rlen=1).cpu().numpy()
    assert rel_err(W, ref) < RTOL


def test_train_matches_oracle_cfg2_full_size():
    """cfg2's real training size: 5,241,600 rows x 32 channels, 10 x 10 SOM, one pass of 32
    mini-batches (the fp64 oracle takes ~20 s)."""
    n, C = 5241600, 32
    base = pixie_like(1 << 18, C, seed=12)
    X = np.ascontiguousarray(np.tile(base, (20, 1))[:n])
    X *= (1.0 + 1e-3 * np.random.default_rng(1).random((n, 1), dtype=np.float32))  # no exact repeats
    idx = oracle.init_codebook_indices(n, 100, 42)
    ref = oracle.som_batch(X, 10, 10, rlen=1, init_idx=idx)
    W = S.train_som(S.to_device_matrix(X), X[idx].astype(np.float64), 10, 10, rlen=1).cpu().numpy()
    assert rel_err(W, ref) < RTOL
