"""GPU parity tests of the BMU path (assignment): the CUDA kernels, reached through the C ABI
(include/pixie_b200.h via ark_analysis_b200.som), against the oracle on the same seeded inputs.

Bar: labels are BIT-EXACT (int32, 1-indexed, first minimum wins, 0 for NaN rows) with
pyFlowSOM.map_data_to_nodes semantics evaluated in fp64 on the same fp32-representable inputs
(/root/reference/src/ark/phenotyping/cluster_helpers.py:152-157); distances are bit-exact fp64.
"""
import os

import numpy as np
import pytest
import torch

import oracle
from ark_analysis_b200 import som as S
from conftest import pixie_like

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def data(kind, n, C, seed=0):
    if kind == "U":
        return np.random.default_rng(seed).random((n, C), dtype=np.float32)
    return pixie_like(n, C, seed + 12345)


def gpu_labels(X, W, flags=S.FLAG_AUTO, stats=None):
    Xd = S.to_device_matrix(X)
    Wd = torch.from_numpy(np.ascontiguousarray(W, np.float32)).cuda()
    lab = S.bmu(Xd, Wd, flags=flags, stats=stats)
    torch.cuda.synchronize()
    return lab.cpu().numpy()


# ------------------------------------------------------------------------------------------------
# golden vectors + seeded sweeps
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["bmu_u16_k100", "bmu_p32_k100", "bmu_ties_c8_k40",
                                  "bmu_nan_c15_k49"])
@pytest.mark.parametrize("flags", [S.FLAG_AUTO, S.FLAG_FORCE_EXACT])
def test_golden_vectors(name, flags):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    np.testing.assert_array_equal(gpu_labels(g["X"], g["W"], flags), g["labels"])


@pytest.mark.parametrize("kind", ["U", "P"])
@pytest.mark.parametrize("C,K", [(16, 100), (32, 100), (40, 400), (100, 100), (64, 100),
                                 (15, 200), (8, 16), (33, 49), (22, 144), (4, 2), (1, 3),
                                 (128, 100), (32, 512), (48, 256), (64, 400), (57, 300)])
def test_tensor_core_kernel_bit_exact(kind, C, K):
    n = 128 * 150 + 77  # ragged last tile, more tiles than SMs so the pipelines wrap
    X = data(kind, n, C, seed=C * 1000 + K)
    W = X[np.random.default_rng(1).choice(n, K, replace=False)].copy()
    stats = torch.zeros(S.NSTATS, dtype=torch.int64, device="cuda")
    lab = gpu_labels(X, W, S.FLAG_FORCE_TC, stats)
    ref, _ = oracle.map_data_to_nodes_f32(W, X)
    np.testing.assert_array_equal(lab, ref)
    assert int(stats[S._native.STAT_KERNEL]) == 1  # the tensor-core kernel really ran


@pytest.mark.parametrize("C,K", [(32, 100), (40, 400), (100, 100)])
def test_trained_codebooks_bit_exact(C, K):
    """Trained maps have close neighbours: many candidates per row, all three stages exercised."""
    n = 40000
    xd = int(round(np.sqrt(K)))
    for kind in ("U", "P"):
        X = data(kind, n, C, seed=5)
        W = oracle.som_batch(X[:20000], xd, K // xd, rlen=1).astype(np.float32)
        stats = torch.zeros(S.NSTATS, dtype=torch.int64, device="cuda")
        lab = gpu_labels(X, W, S.FLAG_FORCE_TC, stats)
        ref, _ = oracle.map_data_to_nodes_f32(W, X)
        np.testing.assert_array_equal(lab, ref)
        assert int(stats[S._native.STAT_ROWS_FLAGGED]) > 0  # the recheck path was taken


@pytest.mark.parametrize("C,K", [(16, 400), (32, 400), (24, 324), (40, 400), (64, 400)])
def test_two_chunk_variants_with_a_deep_pipeline(C, K):
    """K > 256 runs two accumulator chunks per tile whose TMEM buffers are shared by the two
    epilogue groups; with C <= 32 there is room for 6-8 X stages, and a warp that ran a tile ahead
    of its group used to fall through an aliased parity wait (32 wrong labels, then a trap).  The
    race needed many tiles per CTA and a trained map (uneven recheck work): 2^21 rows, three
    launches, every label compared with the exact fp64 kernel."""
    n = 1 << 21
    g = torch.Generator(device="cuda").manual_seed(C * 7 + K)
    X = torch.rand((n, C), device="cuda", generator=g)
    xd = int(round(np.sqrt(K)))
    idx = np.random.default_rng(42).choice(n, K, replace=False)
    W = S.train_som(X[:1 << 20], X[torch.from_numpy(idx).cuda()].double(), xd, K // xd,
                    rlen=1).float().contiguous()
    ref = S.bmu(X, W, flags=S.FLAG_FORCE_EXACT)
    for _ in range(3):
        lab = S.bmu(X, W, flags=S.FLAG_FORCE_TC)
        torch.cuda.synchronize()
        assert int((lab != ref).sum()) == 0
    # the exact kernel itself against the oracle on a slice
    sl = slice(12345, 12345 + 4000)
    o, _ = oracle.map_data_to_nodes_f32(W.cpu().numpy(), X[sl].cpu().numpy())
    np.testing.assert_array_equal(ref[sl].cpu().numpy(), o)


def test_exact_kernel_and_unsupported_shapes_fall_back():
    X = data("U", 3000, 130)  # C > 128: not a tensor-core shape
    W = X[:20].copy()
    ref, _ = oracle.map_data_to_nodes_f32(W, X)
    stats = torch.zeros(S.NSTATS, dtype=torch.int64, device="cuda")
    np.testing.assert_array_equal(gpu_labels(X, W, S.FLAG_AUTO, stats), ref)
    assert int(stats[S._native.STAT_KERNEL]) == 2
    with pytest.raises(S.PixieError):
        gpu_labels(X, W, S.FLAG_FORCE_TC)
    X2 = data("U", 2000, 16)
    W2 = data("U", 600, 16, seed=3)  # K > 512
    ref2, _ = oracle.map_data_to_nodes_f32(W2, X2)
    np.testing.assert_array_equal(gpu_labels(X2, W2), ref2)


# ------------------------------------------------------------------------------------------------
# edge cases the reference's tests and semantics imply
# ------------------------------------------------------------------------------------------------
def test_empty_single_row_and_tiny_inputs():
    W = torch.rand(10, 8, device="cuda")
    assert S.bmu(torch.empty((0, 8), device="cuda"), W).shape == (0,)
    for n in (1, 2, 127, 128, 129):
        X = data("U", n, 8, seed=n)
        ref, _ = oracle.map_data_to_nodes_f32(W.cpu().numpy(), X)
        np.testing.assert_array_equal(gpu_labels(X, W.cpu().numpy(), S.FLAG_FORCE_TC), ref)


def test_nan_inf_rows_and_nan_codebook_rows():
    X = data("U", 5000, 16)
    X[3, 2] = np.nan
    X[130, :] = np.inf
    X[131, 5] = -np.inf
    X[4999, 0] = np.nan
    W = X[1000:1100].copy()
    W[7, 3] = np.nan  # a NaN node can never win (its distance compares false)
    ref, _ = oracle.map_data_to_nodes_f32(W, X)
    lab = gpu_labels(X, W, S.FLAG_FORCE_TC)
    np.testing.assert_array_equal(lab, ref)
    assert lab[3] == 0 and lab[130] == 0 and lab[4999] == 0  # reference: minid = -1 -> label 0
    assert (lab != 8).all()


def test_exact_ties_first_minimum_wins():
    r = np.random.default_rng(0)
    X = r.integers(0, 3, (20000, 12)).astype(np.float32)
    W = X[:64].copy()
    W[32:] = W[:32]  # every node duplicated: the lower index must always win
    ref, _ = oracle.map_data_to_nodes_f32(W, X)
    lab = gpu_labels(X, W, S.FLAG_FORCE_TC)
    np.testing.assert_array_equal(lab, ref)
    assert lab.max() <= 32
    # degenerate codebooks: all nodes identical, all-zero codebook
    for Wd in (np.tile(X[:1], (50, 1)), np.zeros((50, 12), np.float32)):
        np.testing.assert_array_equal(gpu_labels(X[:3000], Wd), np.ones(3000, np.int32))


def test_negative_values_and_large_dynamic_range():
    r = np.random.default_rng(3)
    X = (r.standard_normal((30000, 24)) * np.exp(r.normal(0, 2, (30000, 1)))).astype(np.float32)
    W = X[r.choice(30000, 100, replace=False)].copy()
    ref, _ = oracle.map_data_to_nodes_f32(W, X)
    np.testing.assert_array_equal(gpu_labels(X, W, S.FLAG_FORCE_TC), ref)
    tiny = (X * 1e-22).astype(np.float32)  # fp32 squares underflow: must fall through to fp64
    ref, _ = oracle.map_data_to_nodes_f32(tiny[:100], tiny[:4000])
    np.testing.assert_array_equal(gpu_labels(tiny[:4000], tiny[:100]), ref)


def test_strided_rows_and_unaligned_inputs():
    base = torch.rand((5000, 40), device="cuda")
    X = base[:, :32]  # row pitch 40 floats
    W = X[:100].contiguous()
    ref, _ = oracle.map_data_to_nodes_f32(W.cpu().numpy(), X.cpu().numpy())
    np.testing.assert_array_equal(S.bmu(X, W, flags=S.FLAG_FORCE_TC).cpu().numpy(), ref)
    X2 = torch.rand((5000, 33), device="cuda")[:, 1:]  # 4-byte aligned only -> exact kernel
    W2 = X2[:50].contiguous()
    ref2, _ = oracle.map_data_to_nodes_f32(W2.cpu().numpy(), X2.contiguous().cpu().numpy())
    np.testing.assert_array_equal(S.bmu(X2, W2).cpu().numpy(), ref2)


def test_distances_and_cluster_sums():
    X = data("P", 20000, 32)
    W = X[:100].copy()
    ref_l, ref_d = oracle.map_data_to_nodes_f32(W, X)
    Xd = S.to_device_matrix(X)
    Wd = torch.from_numpy(W).cuda()
    lab, SN = S.cluster_sums(Xd, Wd)
    np.testing.assert_array_equal(lab.cpu().numpy(), ref_l)
    np.testing.assert_array_equal(S.bmu_dists(Xd, Wd, lab).cpu().numpy(), ref_d)  # fp64 bit-exact
    Sref, cref = oracle.cluster_sums_f32(X, ref_l, 100)
    SN = SN.cpu().numpy()
    np.testing.assert_array_equal(SN[:, 32], cref)  # counts are integers: exact
    # channel sums: fp32 partial sums folded in fp64; tolerance 1e-5 relative
    np.testing.assert_allclose(SN[:, :32], Sref, rtol=1e-5, atol=1e-6)
    # the standalone sums kernel adds in a different (equally fixed) order: same to 1e-6 relative
    SN2 = S.label_sums(Xd, lab, 100).cpu().numpy()
    np.testing.assert_allclose(SN2, SN, rtol=1e-6, atol=1e-6)
    np.testing.assert_array_equal(SN2[:, 32], cref)
    np.testing.assert_allclose(SN2[:, :32], Sref, rtol=1e-5, atol=1e-6)
    # each path is run-to-run deterministic (fixed summation order, no atomics)
    np.testing.assert_array_equal(S.label_sums(Xd, lab, 100).cpu().numpy(), SN2)
    np.testing.assert_array_equal(S.cluster_sums(Xd, Wd)[1].cpu().numpy(), SN)


def test_host_entry_points_match_device_path():
    X = data("U", 300000, 32)
    W = X[:100].copy()
    ref_l, ref_d = oracle.map_data_to_nodes_f32(W, X)
    lab, d = S.map_data_to_nodes(W, X, chunk_rows=65536)  # fp32 host buffers
    np.testing.assert_array_equal(lab, ref_l)
    np.testing.assert_array_equal(d, ref_d)
    lab64, d64 = S.map_data_to_nodes(W.astype(np.float64), X.astype(np.float64))
    np.testing.assert_array_equal(lab64, ref_l)
    np.testing.assert_array_equal(d64, ref_d)
    assert lab.dtype == np.int32 and d.dtype == np.float64
    # 15 channels: host path pads the device pitch to 16
    X15 = data("U", 70000, 15)
    ref15, _ = oracle.map_data_to_nodes_f32(X15[:30], X15)
    np.testing.assert_array_equal(S.map_data_to_nodes(X15[:30], X15)[0], ref15)
    np.testing.assert_array_equal(
        S.map_data_to_nodes(X15[:30].astype(np.float64), X15.astype(np.float64))[0], ref15)


def test_fp64_inputs_label_mismatch_rate_is_reported_not_asserted():
    """SURVEY.md section 7 hard part 4: the device matrix is fp32; feeding the oracle the original
    fp64 values can flip near-ties.  Recorded for the results table; only sanity-bounded here."""
    r = np.random.default_rng(9)
    X64 = r.random((100000, 32))
    W64 = X64[:100].copy()
    ref64, _ = oracle.map_data_to_nodes(W64, X64)
    lab, _ = S.map_data_to_nodes(W64, X64)
    rate = float((lab != ref64).mean())
    print(f"label mismatch rate vs fp64-input oracle: {rate:.2e}")
    assert rate < 1e-3


# ------------------------------------------------------------------------------------------------
# full-size, size-independent properties (BASELINE.json config 2: 50 x 1024^2 x 32, 10x10 SOM)
# ------------------------------------------------------------------------------------------------
def test_full_size_properties_cfg2():
    n, C, K = 50 * 1024 * 1024, 32, 100
    g = torch.Generator(device="cuda").manual_seed(42)
    X = torch.rand((n, C), device="cuda", generator=g)
    W = X[torch.randperm(n, device="cuda", generator=g)[:K]].contiguous()
    lab = S.bmu(X, W)
    lab2 = S.bmu(X, W)
    assert torch.equal(lab, lab2)  # deterministic
    assert int(lab.min()) >= 1 and int(lab.max()) <= K  # reference label range 1..K
    # a row that IS a codebook row maps to (the first copy of) itself with distance 0
    d = S.bmu_dists(X, W, lab)
    assert float(d.min()) == 0.0
    # idempotence under the exact kernel on random windows (tensor-core path == fp64 replica)
    for lo in (0, 7_000_003, n - 400_000):
        sl = slice(lo, lo + 400_000)
        ex = S.bmu(X[sl], W, flags=S.FLAG_FORCE_EXACT)
        assert torch.equal(lab[sl], ex)
    # permutation equivariance: permuting rows permutes labels
    perm = torch.randperm(2_000_000, device="cuda", generator=g)
    assert torch.equal(S.bmu(X[:2_000_000][perm].contiguous(), W), lab[:2_000_000][perm])
    # counts of the fused statistics sum to n and match a histogram of the labels
    _, SN = S.cluster_sums(X, W, labels=lab2)
    cnt = SN[:, C].cpu().numpy()
    assert cnt.sum() == n
    np.testing.assert_array_equal(cnt, torch.bincount(lab.long(), minlength=K + 1)[1:].cpu().numpy())
    # oracle spot check on a window the CPU finishes in seconds
    ref, _ = oracle.map_data_to_nodes_f32(W.cpu().numpy(), X[:300_000].cpu().numpy())
    np.testing.assert_array_equal(lab[:300_000].cpu().numpy(), ref)


def _full_size_properties(n, C, K, windows):
    g = torch.Generator(device="cuda").manual_seed(7)
    X = torch.rand((n, C), device="cuda", generator=g)
    W = X[torch.randperm(n, device="cuda", generator=g)[:K]].contiguous()
    lab, SN = S.cluster_sums(X, W)
    assert int(lab.min()) >= 1 and int(lab.max()) <= K
    assert torch.equal(lab, S.bmu(X, W))  # deterministic, with and without the fused sums
    cnt = SN[:, C].cpu().numpy()
    assert cnt.sum() == n
    np.testing.assert_array_equal(cnt, torch.bincount(lab.long(), minlength=K + 1)[1:].cpu().numpy())
    assert float(S.bmu_dists(X, W, lab).min()) == 0.0  # codebook rows map to themselves
    for lo in windows:
        sl = slice(lo, lo + 200_000)
        assert torch.equal(lab[sl], S.bmu(X[sl], W, flags=S.FLAG_FORCE_EXACT))
    ref, _ = oracle.map_data_to_nodes_f32(W.cpu().numpy(), X[:100_000].cpu().numpy())
    np.testing.assert_array_equal(lab[:100_000].cpu().numpy(), ref)


def test_full_size_properties_cfg4_cells():
    """BASELINE.json config 4: 5 M cells x 100 features, 10x10 SOM."""
    _full_size_properties(5_000_000, 100, 100, (0, 2_500_000, 4_800_000))


def test_full_size_properties_cfg3_shard_slice():
    """BASELINE.json config 3 shape (40 channels, 20x20 SOM) on 8 of a GPU's 62 FOVs of 2048^2."""
    _full_size_properties(8 * 2048 * 2048, 40, 400, (0, 16_000_000, 33_000_000))


def test_candidate_window_has_margin(monkeypatch):
    """The tensor-core candidate window is a PROVEN bound on the tf32 score error; this probes how
    much room it has: with the window shrunk to a quarter (PIXIE_DELTA_SCALE, a test knob read by
    the library) far fewer rows reach the recheck and the labels are still bit-identical to the
    exact kernel -- i.e. the bound is not the kind that only just holds."""
    X = S.to_device_matrix(pixie_like(1 << 20, 32, seed=3))
    W = X[:100].contiguous()
    ref = S.bmu(X, W, flags=S.FLAG_FORCE_EXACT)
    flagged = []
    for scale in ("1", "0.25"):
        monkeypatch.setenv("PIXIE_DELTA_SCALE", scale)
        stats = torch.zeros(S.NSTATS, dtype=torch.int64, device="cuda")
        lab = S.bmu(X, W, flags=S.FLAG_FORCE_TC, stats=stats)
        torch.cuda.synchronize()
        assert torch.equal(lab, ref)
        flagged.append(int(stats[0]))  # PIXIE_STAT_ROWS_FLAGGED
    assert flagged[1] < 0.6 * flagged[0]


@pytest.mark.parametrize("C,K", [(16, 100), (24, 64), (32, 100), (7, 33), (20, 104), (32, 96)])
def test_split_operand_kernel_hard_cases(monkeypatch, C, K):
    """The split-operand (3 x tf32) assign kernel (csrc/bmu_x3_kernel.cuh; taken for C <= 24,
    K <= 104, forced here up to C = 32) against the exact fp64 kernel on the inputs that stress a
    narrow candidate window: codebooks of near-identical node pairs (every row has two candidates a
    few ulp apart), rows equal to nodes, mixed signs, values x 1000, NaN / Inf rows, a ragged tail."""
    monkeypatch.setenv("PIXIE_X3", "2")
    n = 128 * 700 + 37
    X = pixie_like(n, C, seed=31 + C)
    W = X[np.random.default_rng(K).choice(n, K, replace=False)].copy()
    W[1::2] = W[:-1:2][: len(W[1::2])] * (1 + 2.0 ** -19)
    X[::53] = W[np.arange(len(X[::53])) % K]
    cases = {"pairs": (X, W), "x1000": (X * 1000, W * 1000), "signed": (X - 0.03, W - 0.03)}
    Xn = X.copy()
    Xn[11::401, C // 2] = np.nan
    Xn[12::401, 0] = np.inf
    cases["nan rows"] = (Xn, W)
    for name, (x, w) in cases.items():
        x = np.ascontiguousarray(x, dtype=np.float32)
        w = np.ascontiguousarray(w, dtype=np.float32)
        Xd, Wd = S.to_device_matrix(x), torch.from_numpy(w).cuda()
        want = S.bmu(Xd, Wd, flags=S.FLAG_FORCE_EXACT)
        stats = torch.zeros(S.NSTATS, dtype=torch.int64, device="cuda")
        got = S.bmu(Xd, Wd, flags=S.FLAG_FORCE_TC, stats=stats)
        torch.cuda.synchronize()
        assert torch.equal(got, want), name
        assert int(stats[4]) == 1
    # and against the oracle itself on a slice (the exact kernel is checked against it elsewhere)
    ref, _ = oracle.map_data_to_nodes_f32(W, X[:20000])
    np.testing.assert_array_equal(S.bmu(S.to_device_matrix(X[:20000]),
                                        torch.from_numpy(W).cuda()).cpu().numpy(), ref)


def test_split_operand_window_has_margin(monkeypatch):
    """The accumulation part of the split-operand kernel's window rests on a model of the tensor
    core's fp32 accumulator plus measurement (bmu_x3_kernel.cuh, scripts/x3_margin.py): labels must
    stay identical to the exact kernel with the window shrunk 16-fold, on plain rows and on a
    codebook of near-identical node pairs, and the plain kernel's recheck load must be gone."""
    X = S.to_device_matrix(pixie_like(1 << 21, 16, seed=8))
    W = X[:100].contiguous()
    W2 = W.clone()
    W2[1::2] = W2[::2] * (1 + 2.0 ** -18)
    for Wd in (W, W2):
        ref = S.bmu(X, Wd, flags=S.FLAG_FORCE_EXACT)
        for scale in ("1", "0.0625"):
            monkeypatch.setenv("PIXIE_DELTA_SCALE", scale)
            stats = torch.zeros(S.NSTATS, dtype=torch.int64, device="cuda")
            lab = S.bmu(X, Wd, flags=S.FLAG_FORCE_TC, stats=stats)
            torch.cuda.synchronize()
            assert torch.equal(lab, ref), scale
            if Wd is W and scale == "1":
                split_flagged = int(stats[0])
        monkeypatch.delenv("PIXIE_DELTA_SCALE")
    monkeypatch.setenv("PIXIE_X3", "0")
    stats = torch.zeros(S.NSTATS, dtype=torch.int64, device="cuda")
    S.bmu(X, W, flags=S.FLAG_FORCE_TC, stats=stats)
    torch.cuda.synchronize()
    assert split_flagged < 0.05 * int(stats[0])


def test_concurrent_host_threads_share_nothing():
    """cluster_pixels(multiprocess=True) labels FOVs from a thread pool; ctypes releases the GIL, so
    the memset -> prep -> BMU -> fix-up chains of several threads interleave on one stream.  Each
    thread must own its control block: a shared one lets another thread's memset zero the codebook
    norm (candidate window collapses, near-ties get the tf32 winner) or the fix-up counter (NaN rows
    keep the sentinel).  Near-tie rows and NaN rows against the exact kernel, many times over."""
    import threading
    C, K, n = 32, 100, 128 * 64 + 5
    X = pixie_like(n, C, seed=99)
    W = X[np.random.default_rng(3).choice(n, K, replace=False)].copy()
    X[::97] = W[np.arange(len(X[::97])) % K]                      # rows equal to nodes
    W[1::2] = W[::2] * (1 + 2.0 ** -20)                           # pairs of near-identical nodes
    X[5::211, 3] = np.nan                                         # rows for the fix-up kernel
    Xd = S.to_device_matrix(X)
    Wd = torch.from_numpy(W).cuda()
    want = S.bmu(Xd, Wd, flags=S.FLAG_FORCE_EXACT).cpu().numpy()
    np.testing.assert_array_equal(want, oracle.map_data_to_nodes_f32(W, X)[0])
    errors = []

    def worker():
        try:
            for _ in range(40):
                got = S.bmu(Xd, Wd).cpu().numpy()
                if not np.array_equal(got, want):
                    errors.append(int((got != want).sum()))
        except Exception as exc:  # noqa: BLE001
            errors.append(repr(exc))

    threads = [threading.Thread(target=worker) for _ in range(8)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors[:5]


def test_torch_library_ops():
    """torch.ops.pixie_b200.* are the same kernels behind torch.library custom ops."""
    import ark_analysis_b200.torch_ops  # noqa: F401  (registers the ops)
    X = pixie_like(128 * 20 + 3, 16, seed=4)
    W = X[:25].copy()
    Xd, Wd = S.to_device_matrix(X), torch.from_numpy(W).cuda()
    lab = torch.ops.pixie_b200.bmu(Xd, Wd)
    np.testing.assert_array_equal(lab.cpu().numpy(), oracle.map_data_to_nodes_f32(W, X)[0])
    lab2, SN = torch.ops.pixie_b200.bmu_sums(Xd, Wd)
    assert torch.equal(lab, lab2) and float(SN[:, -1].sum()) == X.shape[0]
    W64 = torch.ops.pixie_b200.som_train(Xd, Wd.double(), 5, 5, 1, 0.05, 0.01, 0)
    ref = oracle.som_batch(X, 5, 5, rlen=1, init_idx=np.arange(25))
    assert np.abs(W64.cpu().numpy() - ref).max() / np.abs(ref).max() < 1e-4
