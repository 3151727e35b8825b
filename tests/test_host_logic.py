"""CPU tests of the host-side logic: the helper shims, the algorithm's
host-side definitions, the compat aliases, API errors raised before any kernel is needed, and
the rule that the product path fails loudly without a GPU (no CPU fallback)."""
import inspect
import os
import sys

import numpy as np
import pandas as pd
import pytest
import torch

import oracle
from ark_analysis_b200 import (cell_som_clustering, cluster_helpers, compat, distributed,
                               io_utils, pixel_cluster_utils, pixel_som_clustering, som)


def test_validate_paths_and_list_files(tmp_path):
    (tmp_path / "fov1.feather").touch()
    (tmp_path / "fov0.feather").touch()
    (tmp_path / ".hidden.feather").touch()
    (tmp_path / "notes.txt").touch()
    io_utils.validate_paths([str(tmp_path), str(tmp_path / "fov0.feather")])
    io_utils.validate_paths(str(tmp_path))
    with pytest.raises(FileNotFoundError):
        io_utils.validate_paths([str(tmp_path), str(tmp_path / "nope")])
    assert io_utils.list_files(str(tmp_path), substrs=".feather") == ["fov0.feather", "fov1.feather"]
    assert io_utils.list_files(str(tmp_path)) == ["fov0.feather", "fov1.feather", "notes.txt"]
    assert io_utils.remove_file_extensions(["fov0.feather", "a.b.csv"]) == ["fov0", "a.b"]


def test_verify_in_list_and_same_elements():
    assert io_utils.verify_in_list(a=["x"], b=np.array(["x", "y"]))
    assert io_utils.verify_in_list(a="x", b=["x"])
    with pytest.raises(ValueError):
        io_utils.verify_in_list(provided=["x", "z"], valid=["x", "y"])
    assert io_utils.verify_same_elements(a=[1, 2], b=[2, 1])
    with pytest.raises(ValueError):
        io_utils.verify_same_elements(a=[1, 2], b=[1, 3])
    with pytest.raises(ValueError):
        io_utils.verify_same_elements(enforce_order=True, a=[1, 2], b=[2, 1])
    assert io_utils.verify_same_elements(enforce_order=True, a=[1, 2], b=np.array([1, 2]))


def test_feather_roundtrip_and_corruption(tmp_path):
    df = pd.DataFrame({"a": np.arange(5.0), "fov": ["f"] * 5})
    p = str(tmp_path / "t.feather")
    io_utils.write_dataframe(df, p)
    pd.testing.assert_frame_equal(io_utils.read_dataframe(p), df)
    io_utils.write_dataframe(df, p, compression="uncompressed")
    assert io_utils.read_table(p).column_names == ["a", "fov"]
    with open(p, "w") as f:  # the reference's tests corrupt files exactly like this
        f.write("baddatabaddatabaddata")
    with pytest.raises((io_utils.ArrowInvalid, OSError)):
        io_utils.read_dataframe(p)


def test_host_side_algorithm_definitions_agree_with_the_oracle():
    for xd, yd in [(10, 10), (20, 20), (20, 10), (3, 2)]:
        np.testing.assert_array_equal(som.grid_chebyshev(xd, yd), oracle.grid_chebyshev(xd, yd))
        assert som.default_radius(xd, yd) == oracle.default_radius(xd, yd)
    np.testing.assert_array_equal(som.init_codebook_indices(1000, 100, 42),
                                  oracle.init_codebook_indices(1000, 100, 42))
    with pytest.raises(ValueError):
        som.init_codebook_indices(10, 100, 42)  # fewer rows than nodes, as in the reference
    assert som.default_batches(1000) == 8 and som.default_batches(10 ** 7) == 32
    assert som.default_batches(1) == 1
    assert som.TILE == oracle.TILE == distributed.TILE == 128


def test_step_schedule():
    T = 8
    sig0, a0 = distributed.step_schedule(0, T, (0.05, 0.01), (6.0, 0.0))
    assert sig0 == 3.0 and a0 == 0.05
    sig, a = distributed.step_schedule(7, T, (0.05, 0.01), (6.0, 0.0))
    assert sig == 0.25 and abs(a - (0.05 - 0.04 * 7 / 8)) < 1e-15  # radius < 1 -> 0.5 -> sigma .25


def test_sharding_helpers():
    assert distributed.fov_shards(500, 8) == [(0, 63), (63, 126), (126, 189), (189, 252),
                                              (252, 314), (314, 376), (376, 438), (438, 500)]
    assert distributed.fov_shards(3, 4) == [(0, 1), (1, 2), (2, 3), (3, 3)]
    shards = distributed.tile_aligned_row_shards(1000, 3)
    assert shards == [(0, 384), (384, 768), (768, 1000)]
    assert all(lo % 128 == 0 for lo, _ in shards)
    # every global tile of mini-batch m is visited by exactly one rank, whatever the sharding
    n, B = 128 * 37 + 5, 8
    ntiles = -(-n // 128)
    for world in (1, 2, 3, 5):
        seen = []
        for lo, hi in distributed.tile_aligned_row_shards(n, world):
            off = lo // 128
            local_tiles = -(-(hi - lo) // 128)
            for m in range(B):
                first = distributed.first_local_tile(m, B, off)
                seen += [(m, t + off) for t in range(first, local_tiles, B)]
        assert sorted(seen) == sorted((t % B, t) for t in range(ntiles))


def test_signatures_match_the_reference_api():
    """Argument names and defaults the notebooks rely on (reference pixel_som_clustering.py:16-21,
    :139-141, :308-311; cell_som_clustering.py:8-11, :78-79; cluster_helpers.py:167-170, :305-308)."""
    def params(f):
        return [(p.name, p.default) for p in inspect.signature(f).parameters.values()]
    E = inspect.Parameter.empty
    assert params(pixel_som_clustering.train_pixel_som) == [
        ("fovs", E), ("channels", E), ("base_dir", E), ("subset_dir", 'pixel_mat_subsetted'),
        ("norm_vals_name", 'post_rowsum_chan_norm.feather'),
        ("som_weights_name", 'pixel_som_weights.feather'), ("xdim", 10), ("ydim", 10),
        ("lr_start", 0.05), ("lr_end", 0.01), ("num_passes", 1), ("seed", 42),
        ("overwrite", False)]
    assert params(pixel_som_clustering.cluster_pixels) == [
        ("fovs", E), ("base_dir", E), ("pixel_pysom", E), ("data_dir", 'pixel_mat_data'),
        ("multiprocess", False), ("batch_size", 5), ("num_parallel_pixels", 1000000),
        ("overwrite", False)]
    assert params(pixel_som_clustering.generate_som_avg_files) == [
        ("fovs", E), ("channels", E), ("base_dir", E), ("pixel_pysom", E),
        ("data_dir", 'pixel_data_dir'),
        ("pc_chan_avg_som_cluster_name", 'pixel_channel_avg_som_cluster.csv'),
        ("num_fovs_subset", 100), ("require_all_som_clusters", True), ("seed", 42),
        ("overwrite", False)]
    assert params(cell_som_clustering.train_cell_som) == [
        ("fovs", E), ("base_dir", E), ("cell_table_path", E), ("cell_som_cluster_cols", E),
        ("cell_som_input_data", E), ("som_weights_name", 'cell_som_weights.feather'),
        ("xdim", 10), ("ydim", 10), ("lr_start", 0.05), ("lr_end", 0.01), ("num_passes", 1),
        ("seed", 42), ("overwrite", False), ("normalize", True)]
    assert params(cell_som_clustering.cluster_cells) == [
        ("base_dir", E), ("cell_pysom", E), ("cell_som_cluster_cols", E),
        ("num_parallel_cells", 1000000), ("overwrite", False)]
    assert [n for n, _ in params(cluster_helpers.PixelSOMCluster.__init__)] == [
        "self", "pixel_subset_folder", "norm_vals_path", "weights_path", "fovs", "columns",
        "num_passes", "xdim", "ydim", "lr_start", "lr_end", "seed"]
    assert [n for n, _ in params(cluster_helpers.CellSOMCluster.__init__)] == [
        "self", "cell_data", "weights_path", "fovs", "columns", "num_passes", "xdim", "ydim",
        "lr_start", "lr_end", "seed", "normalize"]


def test_compat_aliases():
    for name in list(sys.modules):
        if name == "pyFlowSOM" or name == "ark" or name.startswith("ark."):
            del sys.modules[name]
    names = compat.install()
    assert "pyFlowSOM" in names and "ark.phenotyping.pixel_som_clustering" in names
    import pyFlowSOM
    from ark.phenotyping import cluster_helpers as ch
    assert pyFlowSOM.som is som.som and pyFlowSOM.map_data_to_nodes is som.map_data_to_nodes
    assert ch is cluster_helpers
    # the rows either side of the SOM resolve under the reference's module names too
    from ark.phenotyping import cell_cluster_utils as ccu, pixie_preprocessing as pp
    from ark.utils import data_utils as du
    assert callable(pp.create_fov_pixel_data) and callable(ccu.create_c2pc_data)
    assert callable(du.generate_pixel_cluster_mask)


def _make_pixel_dirs(base, fovs, chans, n=200, seed=0):
    r = np.random.default_rng(seed)
    os.mkdir(os.path.join(base, "pixel_mat_subsetted"))
    os.mkdir(os.path.join(base, "pixel_mat_data"))
    for fov in fovs:
        for d, rows in (("pixel_mat_subsetted", n // 2), ("pixel_mat_data", n)):
            df = pd.DataFrame(r.random((rows, len(chans))), columns=chans)
            df["fov"] = fov
            df["row_index"] = r.integers(0, 10, rows)
            df["column_index"] = r.integers(0, 10, rows)
            df["segmentation_label"] = r.integers(1, 5, rows)
            io_utils.write_dataframe(df, os.path.join(base, d, fov + ".feather"))
    io_utils.write_dataframe(pd.DataFrame(np.full((1, len(chans)), 0.5), columns=chans),
                             os.path.join(base, "post_rowsum_chan_norm.feather"))


def test_api_errors_that_need_no_gpu(tmp_path):
    base = str(tmp_path)
    fovs, chans = ["fov0", "fov1"], ["Marker1", "Marker2", "Marker3"]
    _make_pixel_dirs(base, fovs, chans)
    # reference pixel_som_clustering_test.py:96-139
    with pytest.raises(FileNotFoundError):
        pixel_som_clustering.train_pixel_som(fovs, chans, base, subset_dir="bad_path")
    with pytest.raises(FileNotFoundError):
        pixel_som_clustering.train_pixel_som(fovs, chans, base, norm_vals_name="bad.feather")
    with pytest.raises(ValueError):
        pixel_som_clustering.train_pixel_som(["fov2", "fov3"], chans, base)
    with pytest.raises(ValueError):
        pixel_som_clustering.train_pixel_som(fovs, ["Marker4"], base)

    pysom = cluster_helpers.PixelSOMCluster(
        os.path.join(base, "pixel_mat_subsetted"),
        os.path.join(base, "post_rowsum_chan_norm.feather"),
        os.path.join(base, "weights.feather"), fovs, chans)
    assert pysom.weights is None and pysom.som_clusters_seen == set()
    assert pysom.train_data.shape[0] == 200
    # normalisation: values / 0.5 (reference cluster_helpers_test.py:286-302)
    raw = io_utils.read_dataframe(os.path.join(base, "pixel_mat_subsetted", "fov0.feather"))
    np.testing.assert_allclose(pysom.train_data[chans].values[:100], raw[chans].values / 0.5)
    # untrained object (reference pixel_som_clustering_test.py:224-226)
    with pytest.raises(ValueError, match="untrained pixel_pysom"):
        pixel_som_clustering.cluster_pixels(fovs, base, pysom)
    with pytest.raises(ValueError, match="untrained pixel_pysom"):
        pixel_som_clustering.generate_som_avg_files(fovs, chans, base, pysom)
    # num_parallel_obs <= 0 (reference cluster_helpers_test.py:406-420)
    pysom.weights = pd.DataFrame(np.zeros((100, 3)), columns=chans)
    with pytest.raises(ValueError, match="greater than 0"):
        pysom.generate_som_clusters(pysom.train_data, num_parallel_obs=0)
    # empty table -> empty label array (reference cluster_helpers.py:160-161)
    assert pysom.generate_som_clusters(pysom.train_data.iloc[:0]).shape == (0,)
    # restart protocol (reference pixel_cluster_utils.py:419-478)
    assert pixel_cluster_utils.find_fovs_missing_col(base, "pixel_mat_data",
                                                     "pixel_som_cluster") == fovs
    assert os.path.isdir(os.path.join(base, "pixel_mat_data_temp"))
    io_utils.make_blank_file(os.path.join(base, "pixel_mat_data_temp"), "fov0.feather")
    assert pixel_cluster_utils.find_fovs_missing_col(base, "pixel_mat_data",
                                                     "pixel_som_cluster") == ["fov1"]


def test_cell_api_errors_that_need_no_gpu(tmp_path):
    df = pd.DataFrame(np.random.default_rng(0).random((50, 3)), columns=["a", "b", "c"])
    df["fov"] = ["fov0"] * 25 + ["fov1"] * 25
    df["label"] = np.arange(50)
    table = str(tmp_path / "cell_table.csv")
    df.to_csv(table)
    with pytest.raises(FileNotFoundError):
        cell_som_clustering.train_cell_som(["fov0"], str(tmp_path), "bad.csv", ["a"], df)
    with pytest.raises(ValueError):
        cell_som_clustering.train_cell_som(["fov0"], str(tmp_path), table, ["zzz"], df)
    pysom = cluster_helpers.CellSOMCluster(df, str(tmp_path / "w.feather"), ["fov0"],
                                           ["a", "b", "c"])
    assert pysom.cell_data.shape[0] == 25
    # 99.9 % normalisation, zeros ignored (reference cluster_helpers.py:366-369)
    q = pysom.cell_data[["a", "b", "c"]].replace(0, np.nan).quantile(q=0.999, axis=0)
    np.testing.assert_allclose(q.values, 1.0, rtol=1e-12)
    with pytest.raises(ValueError, match="untrained cell_pysom"):
        cell_som_clustering.cluster_cells(str(tmp_path), pysom, ["a", "b", "c"])
    with pytest.raises(ValueError, match="does not have SOM labels"):
        cell_som_clustering.generate_som_avg_files(str(tmp_path), df, ["a"], "avg.csv")


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_product_path_fails_loudly_without_a_gpu():
    """No CPU fallback: without a CUDA device the operators raise instead of computing."""
    X = np.random.default_rng(0).random((10, 4))
    with pytest.raises(som.PixieError):
        som.som(X, 2, 2, rlen=1, seed=1)
    with pytest.raises(som.PixieError):
        som.map_data_to_nodes(X[:4], X)
    with pytest.raises(som.PixieError):
        som.bmu(torch.zeros(4, 4), torch.zeros(2, 4))
    # the rows either side of the SOM (N3 / N4) have no CPU route either
    from ark_analysis_b200 import pixie_preprocessing as PP
    z = torch.zeros(4, dtype=torch.int32)
    with pytest.raises(som.PixieError):
        som.label_histogram(z, z, 2, 2)
    with pytest.raises(som.PixieError):
        som.scatter_labels(z, z, z, 2, 2)
    with pytest.raises(som.PixieError):
        PP.preprocess_fov_device(np.zeros((4, 4, 2), np.float32))


def test_gaussian_taps_are_scipys_kernel_bit_for_bit():
    """The blur weights are computed on the host and handed to the kernel: they must be the very
    numbers scipy.ndimage uses (its private _gaussian_kernel1d when importable, else the public
    filter's impulse response)."""
    from scipy import ndimage
    from ark_analysis_b200 import pixie_preprocessing as PP
    for sigma in (1, 2, 2.5, 3.5):
        taps, radius = PP.gaussian_taps(sigma)
        assert radius == int(4.0 * sigma + 0.5) and taps.shape == (radius + 1,)
        try:
            from scipy.ndimage._filters import _gaussian_kernel1d
            ref = _gaussian_kernel1d(float(sigma), 0, radius)[::-1]
            np.testing.assert_array_equal(taps, ref[radius:])
        except ImportError:
            pass
        impulse = np.zeros(4 * radius + 1)
        impulse[2 * radius] = 1.0
        resp = ndimage.gaussian_filter1d(impulse, sigma)
        np.testing.assert_array_equal(resp[2 * radius:3 * radius + 1], taps)
    assert PP.gaussian_taps(0)[1] == 0
    with pytest.raises(som.PixieError):
        PP.gaussian_taps(20)      # radius 80 > the kernel's 32 taps
    names = ['chan10', 'chan2', 'chan1', 'CD45', 'CD4']
    names.sort(key=PP.natural_key)
    assert names == ['CD4', 'CD45', 'chan1', 'chan2', 'chan10']


def test_product_package_never_imports_the_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "ark_analysis_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f
                assert "pixie_oracle" not in text.replace("oracle/pixie_oracle.c", ""), f


def test_libc_sample_indices_reproduce_the_reference_stream():
    """pixie_libc_sample_indices (host helper of the online parity mode) must draw exactly what
    C_SOM draws: i = (int)(n * rand() / (RAND_MAX + 1.0)) after srand(seed) -- checked against libc
    itself and against the oracle's own use of that stream (same seed, same trained codebook is
    asserted on the GPU in tests/test_train_gpu.py)."""
    import ctypes
    from ark_analysis_b200 import _native
    lib = ctypes.CDLL(_native.build())
    lib.pixie_libc_sample_indices.argtypes = [ctypes.c_uint32, ctypes.c_int64, ctypes.c_int64,
                                              ctypes.c_void_p]
    n, count, seed = 12345, 5000, 42
    out = np.empty(count, np.int64)
    assert lib.pixie_libc_sample_indices(seed, n, count, out.ctypes.data_as(ctypes.c_void_p)) == 0
    libc = ctypes.CDLL("libc.so.6")
    libc.srand(seed)
    rand_max = 2147483647
    want = np.array([int(n * (libc.rand() / (rand_max + 1.0))) for _ in range(count)], np.int64)
    assert np.array_equal(out, want)
    assert out.min() >= 0 and out.max() < n
    assert lib.pixie_libc_sample_indices(seed, 0, 1, out.ctypes.data_as(ctypes.c_void_p)) < 0
