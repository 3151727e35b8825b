"""GPU tests of the N2 row: Feather column buffers -> device matrix
(`pixie_columns_to_rows_f32`) and the Arrow-native per-FOV assignment built on it.  The oracle for
the kernel is numpy's float64 division followed by the float32 cast -- the two host passes
(cluster_helpers.py:244-246 normalize_data, :153 astype) it replaces; the oracle for the file path
is this package's own DataFrame path, which mirrors the reference line by line."""
import os

import numpy as np
import pandas as pd
import pytest
import torch

from ark_analysis_b200 import cluster_helpers, io_utils, pixel_som_clustering
from ark_analysis_b200 import som as S
from test_api_gpu import make_pixel_data

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,C", [(1, 1), (127, 7), (128, 32), (129, 40), (100003, 32), (5000, 300),
                                 (70000, 16), (33, 100)])
@pytest.mark.parametrize("with_div", [True, False])
def test_columns_to_rows_is_bit_exact(n, C, with_div):
    r = np.random.default_rng(n * 31 + C)
    cols = r.random((C, n)) * r.choice([1e-3, 1.0, 250.0], size=(C, 1))
    cols[:, ::17] = 0.0
    div = r.random(C) * 3 + 1e-3 if with_div else None
    want = (cols / div[:, None] if with_div else cols).T.astype(np.float32)
    X = S.columns_to_rows(torch.from_numpy(cols).cuda(),
                          torch.from_numpy(div).cuda() if with_div else None)
    assert X.shape == (n, C) and X.stride(0) % 4 == 0
    got = X.cpu().numpy()
    assert np.array_equal(got, want)
    # the matrix feeds the BMU kernel as is
    W = X[:min(n, 5)].contiguous()
    lab = S.bmu(X, W).cpu().numpy()
    assert lab.min() >= 1 and lab.max() <= W.shape[0]


def test_columns_to_rows_strided_columns_and_caller_buffer():
    r = np.random.default_rng(5)
    n, C = 4000, 12
    big = torch.from_numpy(r.random((C, n + 96))).cuda()   # column stride > n
    cols = big[:, :n]
    out = torch.full((n + 7, 16), -1.0, dtype=torch.float32, device="cuda")
    X = S.columns_to_rows(cols, out=out)
    assert X.data_ptr() == out.data_ptr() and X.stride(0) == 16
    assert np.array_equal(X.cpu().numpy(), big[:, :n].T.float().cpu().numpy())
    assert float(out[:, C:].max()) == -1.0 and float(out[n:].max()) == -1.0  # nothing else written


def test_columns_to_rows_rejects_bad_arguments():
    with pytest.raises(S.PixieError):
        S.columns_to_rows(torch.zeros((3, 10), dtype=torch.float32, device="cuda"))
    with pytest.raises(S.PixieError):
        S.columns_to_rows(torch.zeros((3, 10), dtype=torch.float64, device="cuda"),
                          divisor=torch.ones(4, dtype=torch.float64))
    assert S.columns_to_rows(torch.zeros((3, 0), dtype=torch.float64, device="cuda")).shape == (0, 3)


def test_arrow_path_writes_the_same_file_as_the_dataframe_path(tmp_path):
    chans = [f"Marker{i}" for i in range(6)]
    fovs = ["fov0", "fov1"]
    base = str(tmp_path)
    norm_path, weights_path = make_pixel_data(base, fovs, chans, rows=3000, with_weights=100)
    pysom = cluster_helpers.PixelSOMCluster(
        os.path.join(base, "pixel_mat_subsetted"), norm_path, weights_path, fovs, chans[:4])
    data_dir = os.path.join(base, "pixel_mat_data")
    os.makedirs(data_dir + "_temp")
    for fov in fovs:
        # Arrow-native path
        assert pixel_som_clustering.run_pixel_som_assignment(data_dir, pysom, False, 1000000,
                                                             fov) == (fov, 0)
        fast = io_utils.read_dataframe(os.path.join(data_dir + "_temp", fov + ".feather"))
        # DataFrame path (the line-by-line mirror of the reference)
        df = io_utils.read_dataframe(os.path.join(data_dir, fov + ".feather"))
        slow = pysom.assign_som_clusters(df, normalize_data=True)
        pd.testing.assert_frame_equal(fast, slow, check_exact=True)
        assert fast["pixel_som_cluster"].dtype == np.int32
        # channels outside the SOM columns are normalised too (normalize_data divides them all)
        assert np.array_equal(fast[chans[5]].values, df[chans[5]].values / 0.5)
    assert pysom.som_clusters_seen


def test_arrow_path_falls_back_for_tables_it_does_not_take(tmp_path):
    chans = ["a", "b", "c"]
    base = str(tmp_path)
    norm_path, weights_path = make_pixel_data(base, ["fov0"], chans, rows=500, with_weights=9)
    pysom = cluster_helpers.PixelSOMCluster(
        os.path.join(base, "pixel_mat_subsetted"), norm_path, weights_path, ["fov0"], chans,
        xdim=3, ydim=3)
    table = io_utils.read_table(os.path.join(base, "pixel_mat_data", "fov0.feather"))
    import pyarrow as pa
    f32 = table.set_column(0, "a", table.column("a").cast(pa.float32()))
    assert pysom.assign_som_clusters_table(f32) is None           # not float64
    assert pysom.assign_som_clusters_table(table.slice(0, 0)) is None  # no rows
    with pytest.raises(ValueError):
        pysom.assign_som_clusters_table(table.drop_columns(["b"]))
    out = pysom.assign_som_clusters_table(table, normalize_data=False)
    assert out.column_names[-1] == "pixel_som_cluster"
    assert np.array_equal(out.column("a").to_numpy(), table.column("a").to_numpy())
