"""GPU tests of the reference-facing API (the ark-level boundary).

These read like the reference's own tests for this path and assert what those pin
(/root/reference/tests/phenotyping/cluster_helpers_test.py:286-517,
pixel_som_clustering_test.py:17-438, cell_som_clustering_test.py:17-252,
pixel_cluster_utils_test.py:355-487): files written, shapes, column names, label range 1..K,
warnings and printed messages, restart / overwrite / corrupted-FOV behaviour -- plus, beyond the
reference, label parity with the oracle through the whole file-level pipeline."""
import os
import warnings

import numpy as np
import pandas as pd
import pytest

import oracle
from ark_analysis_b200 import (cell_som_clustering, cluster_helpers, io_utils,
                               pixel_cluster_utils, pixel_som_clustering)

pytestmark = pytest.mark.gpu


def make_pixel_data(base, fovs, chans, rows=1000, subset_rows=300, seed=0, with_weights=None):
    r = np.random.default_rng(seed)
    os.makedirs(os.path.join(base, "pixel_mat_data"))
    os.makedirs(os.path.join(base, "pixel_mat_subsetted"))
    for fov in fovs:
        for d, n in (("pixel_mat_data", rows), ("pixel_mat_subsetted", subset_rows)):
            df = pd.DataFrame(r.random((n, len(chans))), columns=chans)
            df["fov"] = fov
            df["row_index"] = r.integers(0, 40, n)
            df["column_index"] = r.integers(0, 40, n)
            df["segmentation_label"] = r.integers(1, 9, n)
            io_utils.write_dataframe(df, os.path.join(base, d, fov + ".feather"))
    norm_path = os.path.join(base, "post_rowsum_chan_norm.feather")
    io_utils.write_dataframe(pd.DataFrame(np.full((1, len(chans)), 0.5), columns=chans), norm_path)
    weights_path = os.path.join(base, "pixel_som_weights.feather")
    if with_weights is not None:
        io_utils.write_dataframe(pd.DataFrame(r.random((with_weights, len(chans))), columns=chans),
                                 weights_path)
    return norm_path, weights_path


# ------------------------------------------------------------------------------------------------
# PixelSOMCluster (reference cluster_helpers_test.py:256-420)
# ------------------------------------------------------------------------------------------------
class TestPixelSOMCluster:
    chans = [f"Marker{i}" for i in range(6)]
    fovs = ["fov0", "fov1", "fov2"]

    def make(self, base, columns=None, **kw):
        norm_path, weights_path = make_pixel_data(base, self.fovs, self.chans)
        return cluster_helpers.PixelSOMCluster(
            os.path.join(base, "pixel_mat_subsetted"), norm_path, weights_path, self.fovs,
            columns or self.chans, xdim=20, ydim=10, **kw)

    def test_train_restart_overwrite_new_cols(self, tmp_path):
        pysom = self.make(str(tmp_path))
        pysom.train_som()
        assert os.path.exists(pysom.weights_path)
        assert list(pysom.weights.columns) == self.chans
        assert pysom.weights.shape == (200, 6)
        saved = io_utils.read_dataframe(pysom.weights_path)
        np.testing.assert_array_equal(saved.values, pysom.weights.values)

        with pytest.warns(UserWarning, match='Pixel SOM already trained on specified markers'):
            pysom.train_som()
        first = pysom.weights.values.copy()
        with pytest.warns(UserWarning, match='Overwrite flag set, retraining SOM'):
            pysom.train_som(overwrite=True)
        assert np.allclose(pysom.weights.values, first)  # same seed -> same weights

        # a fresh object picks the weights file up; different markers -> retrain
        pysom2 = cluster_helpers.PixelSOMCluster(
            os.path.join(str(tmp_path), "pixel_mat_subsetted"),
            os.path.join(str(tmp_path), "post_rowsum_chan_norm.feather"),
            pysom.weights_path, self.fovs, self.chans[:5], xdim=20, ydim=10)
        assert pysom2.weights is not None
        with pytest.warns(UserWarning, match='New markers specified, retraining'):
            pysom2.train_som()
        assert pysom2.weights.shape == (200, 5)

    def test_training_matches_oracle(self, tmp_path):
        pysom = self.make(str(tmp_path), seed=7)
        with warnings.catch_warnings():
            warnings.simplefilter("error")
            pysom.train_som()
        X = np.ascontiguousarray(pysom.train_data[self.chans].to_numpy(), np.float32)
        # a table this small trains with the reference's own online rule (som.som "auto"): the
        # weights equal the C restatement of pyFlowSOM's C_SOM bit for bit
        ref = oracle.som_online(X.astype(np.float64), 20, 10, rlen=1, seed=7)
        np.testing.assert_array_equal(pysom.weights.values, ref)
        # and the batch SOM (what a large table gets) agrees with ITS oracle
        from ark_analysis_b200 import som as S
        Wb = S.som(X, xdim=20, ydim=10, rlen=1, seed=7, algorithm="batch")
        refb = oracle.som_batch(X, 20, 10, rlen=1, seed=7)
        assert np.abs(Wb - refb).max() / np.abs(refb).max() < 1e-4

    @pytest.mark.parametrize("num_parallel_pixels", [10, 10000])
    def test_assign_som_clusters(self, tmp_path, num_parallel_pixels):
        pysom = self.make(str(tmp_path))
        pysom.train_som()
        ext = pd.DataFrame(np.random.default_rng(3).random((1000, 6)), columns=self.chans)
        ext["fov"] = "fov0"
        out = pysom.assign_som_clusters(ext, num_parallel_pixels=num_parallel_pixels)
        assert "pixel_som_cluster" in out.columns
        labels = out["pixel_som_cluster"].to_numpy()
        assert labels.min() >= 1 and labels.max() <= 200
        assert pysom.som_clusters_seen == set(np.unique(labels).tolist())
        np.testing.assert_allclose(out[self.chans].values, ext[self.chans].values / 0.5)
        # parity with the oracle on the normalised fp32 values
        Xn = np.ascontiguousarray(ext[self.chans].values / 0.5, np.float32)
        ref, _ = oracle.map_data_to_nodes_f32(pysom.weights.values.astype(np.float32), Xn)
        np.testing.assert_array_equal(labels, ref)
        # shuffled columns give the same labels (columns follow the weights' order)
        shuffled = ext[self.chans[::-1] + ["fov"]]
        np.testing.assert_array_equal(
            pysom.assign_som_clusters(shuffled, num_parallel_pixels=num_parallel_pixels)
            ["pixel_som_cluster"].to_numpy(), labels)
        # already-normalised data, normalize_data=False: identical labels, values untouched
        again = pysom.assign_som_clusters(out.drop(columns="pixel_som_cluster"),
                                          normalize_data=False)
        np.testing.assert_array_equal(again["pixel_som_cluster"].to_numpy(), labels)
        np.testing.assert_array_equal(again[self.chans].values, out[self.chans].values)

    def test_assign_bad(self, tmp_path):
        pysom = self.make(str(tmp_path))
        pysom.train_som()
        with pytest.raises(ValueError):
            pysom.assign_som_clusters(pysom.train_data, num_parallel_pixels=0)


# ------------------------------------------------------------------------------------------------
# pixel_som_clustering (reference pixel_som_clustering_test.py)
# ------------------------------------------------------------------------------------------------
CHANS = ['Marker1', 'Marker2', 'Marker3', 'Marker4']
FOVS = ['fov0', 'fov1', 'fov2']


def test_train_pixel_som(tmp_path, capsys):
    base = str(tmp_path)
    make_pixel_data(base, FOVS, CHANS)
    pysom = pixel_som_clustering.train_pixel_som(FOVS, CHANS, base)
    assert "Training SOM" in capsys.readouterr().out
    assert os.path.exists(os.path.join(base, 'pixel_som_weights.feather'))
    assert pysom.weights.shape == (100, 4)
    io_utils.verify_same_elements(enforce_order=True, cols=pysom.weights.columns.values,
                                  chans=CHANS)


def test_run_pixel_som_assignment(tmp_path):
    base = str(tmp_path)
    norm_path, weights_path = make_pixel_data(base, FOVS, CHANS, with_weights=100)
    pysom = cluster_helpers.PixelSOMCluster(os.path.join(base, 'pixel_mat_subsetted'), norm_path,
                                            weights_path, FOVS, CHANS)
    data_path = os.path.join(base, 'pixel_mat_data')
    os.mkdir(data_path + '_temp')
    assert pixel_som_clustering.run_pixel_som_assignment(
        data_path, pysom, False, 1000000, 'fov0') == ('fov0', 0)
    out = io_utils.read_dataframe(os.path.join(data_path + '_temp', 'fov0.feather'))
    assert np.all(out['pixel_som_cluster'] <= 100) and np.all(out['pixel_som_cluster'] >= 1)
    with open(os.path.join(data_path, 'fov1.feather'), 'w') as f:
        f.write('baddatabaddatabaddata')
    assert pixel_som_clustering.run_pixel_som_assignment(
        data_path, pysom, False, 1000000, 'fov1') == ('fov1', 1)


@pytest.mark.parametrize('multiprocess', [True, False])
def test_cluster_pixels_base(tmp_path, capsys, multiprocess):
    base = str(tmp_path)
    norm_path, weights_path = make_pixel_data(base, FOVS, CHANS, with_weights=100)
    pysom = cluster_helpers.PixelSOMCluster(os.path.join(base, 'pixel_mat_data'), norm_path,
                                            weights_path, FOVS, CHANS)
    raw = {f: io_utils.read_dataframe(os.path.join(base, 'pixel_mat_data', f + '.feather'))
           for f in FOVS}
    pixel_som_clustering.cluster_pixels(FOVS, base, pysom, 'pixel_mat_data',
                                        multiprocess=multiprocess)
    out = capsys.readouterr().out
    assert "Mapping pixel data to SOM cluster labels" in out and "Processed 3 fovs" in out
    assert not os.path.exists(os.path.join(base, 'pixel_mat_data_temp'))
    W32 = np.ascontiguousarray(pysom.weights.values, np.float32)
    seen = set()
    for fov in FOVS:
        df = io_utils.read_dataframe(os.path.join(base, 'pixel_mat_data', fov + '.feather'))
        labels = df['pixel_som_cluster'].to_numpy()
        assert np.all(labels <= 100) and np.all(labels >= 1)
        # the stored table holds the NORMALISED channels plus the labels (reference :289-301)
        np.testing.assert_allclose(df[CHANS].values, raw[fov][CHANS].values / 0.5)
        ref, _ = oracle.map_data_to_nodes_f32(
            W32, np.ascontiguousarray(raw[fov][CHANS].values / 0.5, np.float32))
        np.testing.assert_array_equal(labels, ref)
        seen |= set(np.unique(labels).tolist())
    assert pysom.som_clusters_seen == seen  # kept even with multiprocess=True

    # everything labelled: nothing left to do
    pixel_som_clustering.cluster_pixels(FOVS, base, pysom, 'pixel_mat_data',
                                        multiprocess=multiprocess)
    assert "There are no more FOVs to assign SOM labels to, skipping" in capsys.readouterr().out

    # overwrite: all FOVs again, without normalising twice
    pixel_som_clustering.cluster_pixels(FOVS, base, pysom, 'pixel_mat_data',
                                        multiprocess=multiprocess, overwrite=True)
    out = capsys.readouterr().out
    assert "Overwrite flag set, reassigning SOM cluster labels to all FOVs\n" in out
    assert "There are no more FOVs to assign SOM labels to" not in out
    df = io_utils.read_dataframe(os.path.join(base, 'pixel_mat_data', 'fov0.feather'))
    np.testing.assert_allclose(df[CHANS].values, raw['fov0'][CHANS].values / 0.5)


def test_cluster_pixels_restart(tmp_path, capsys):
    base = str(tmp_path)
    norm_path, weights_path = make_pixel_data(base, FOVS, CHANS, with_weights=100)
    pysom = cluster_helpers.PixelSOMCluster(os.path.join(base, 'pixel_mat_data'), norm_path,
                                            weights_path, FOVS, CHANS)
    data_path = os.path.join(base, 'pixel_mat_data')
    os.mkdir(data_path + '_temp')
    assert pixel_som_clustering.run_pixel_som_assignment(
        data_path, pysom, False, 1000000, 'fov0') == ('fov0', 0)
    pixel_som_clustering.cluster_pixels(FOVS, base, pysom, 'pixel_mat_data')
    out = capsys.readouterr().out
    assert "Restarting SOM label assignment from fov fov1, 2 fovs left to process" in out
    assert io_utils.list_files(data_path) == ['fov0.feather', 'fov1.feather', 'fov2.feather']


@pytest.mark.parametrize('multiprocess', [True, False])
def test_cluster_pixels_corrupt(tmp_path, capsys, multiprocess):
    base = str(tmp_path)
    norm_path, weights_path = make_pixel_data(base, FOVS, CHANS, with_weights=100)
    pysom = cluster_helpers.PixelSOMCluster(os.path.join(base, 'pixel_mat_data'), norm_path,
                                            weights_path, FOVS, CHANS)
    with open(os.path.join(base, 'pixel_mat_data', 'fov1.feather'), 'w') as f:
        f.write('baddatabaddatabaddata')
    pixel_som_clustering.cluster_pixels(FOVS, base, pysom, data_dir='pixel_mat_data',
                                        multiprocess=multiprocess)
    assert not os.path.exists(os.path.join(base, 'pixel_mat_data_temp'))
    assert "The data for FOV fov1 has been corrupted, skipping\n" in capsys.readouterr().out
    assert io_utils.list_files(os.path.join(base, 'pixel_mat_data')) == \
        ['fov0.feather', 'fov2.feather']


def test_cluster_pixels_column_order_is_enforced(tmp_path):
    base = str(tmp_path)
    norm_path, weights_path = make_pixel_data(base, FOVS, CHANS, with_weights=100)
    pysom = cluster_helpers.PixelSOMCluster(os.path.join(base, 'pixel_mat_data'), norm_path,
                                            weights_path, FOVS, CHANS)
    pysom.weights = pysom.weights[CHANS[::-1]]
    with pytest.raises(ValueError):
        pixel_som_clustering.cluster_pixels(FOVS, base, pysom, 'pixel_mat_data')


def test_generate_som_avg_files(tmp_path, capsys):
    base = str(tmp_path)
    r = np.random.default_rng(0)
    colnames = CHANS + ['fov', 'row_index', 'column_index', 'label']
    os.mkdir(os.path.join(base, 'pixel_data_dir'))
    frames = []
    for i, fov in enumerate(FOVS):
        df = pd.DataFrame(r.random((100, len(colnames))), columns=colnames)
        df['pixel_som_cluster'] = i + 1
        frames.append(df)
        io_utils.write_dataframe(df, os.path.join(base, 'pixel_data_dir', fov + '.feather'))
    norm_path = os.path.join(base, 'norm_vals.feather')
    io_utils.write_dataframe(pd.DataFrame(r.random((1, 4)), columns=CHANS), norm_path)
    weights_path = os.path.join(base, 'pixel_weights.feather')
    io_utils.write_dataframe(pd.DataFrame(r.random((3, 4)), columns=CHANS), weights_path)
    pysom = cluster_helpers.PixelSOMCluster(os.path.join(base, 'pixel_data_dir'), norm_path,
                                            weights_path, FOVS, CHANS)
    pysom.som_clusters_seen = set(range(3))
    pixel_som_clustering.generate_som_avg_files(FOVS, CHANS, base, pysom, 'pixel_data_dir',
                                                num_fovs_subset=3)
    avg_file = os.path.join(base, 'pixel_channel_avg_som_cluster.csv')
    avg = pd.read_csv(avg_file)
    assert list(avg['pixel_som_cluster']) == [1, 2, 3]
    assert np.all(avg['count'] == 100)
    want = np.stack([f[CHANS].values.mean(0) for f in frames])
    np.testing.assert_allclose(avg[CHANS].values, want, rtol=1e-5)  # fp32 rows, fp64 totals
    capsys.readouterr()
    pixel_som_clustering.generate_som_avg_files(FOVS, CHANS, base, pysom, 'pixel_data_dir',
                                                num_fovs_subset=1)
    assert capsys.readouterr().out == "Already generated SOM cluster channel average file, skipping\n"
    pixel_som_clustering.generate_som_avg_files(FOVS, CHANS, base, pysom, 'pixel_data_dir',
                                                num_fovs_subset=3, overwrite=True)
    assert "Overwrite flag set, regenerating SOM cluster channel average file\n" in \
        capsys.readouterr().out
    os.remove(avg_file)
    pysom.som_clusters_seen = set(range(200))
    with pytest.raises(ValueError):
        pixel_som_clustering.generate_som_avg_files(FOVS, CHANS, base, pysom, 'pixel_data_dir',
                                                    num_fovs_subset=1)
    pixel_som_clustering.generate_som_avg_files(FOVS, CHANS, base, pysom, 'pixel_data_dir',
                                                num_fovs_subset=1, require_all_som_clusters=False)
    assert os.path.exists(avg_file)


def test_compute_pixel_cluster_channel_avg_known_answer(tmp_path):
    """reference pixel_cluster_utils_test.py:355-487: constant rows [0.1, 0.2, 0.3] per cluster."""
    base = str(tmp_path)
    chans = ['chan0', 'chan1', 'chan2']
    os.mkdir(os.path.join(base, 'pixel_mat_data'))
    for fov in ['fov0', 'fov1']:
        df = pd.DataFrame(np.tile([0.1, 0.2, 0.3], (1000, 1)), columns=chans)
        df['fov'] = fov
        df['pixel_som_cluster'] = np.repeat(np.arange(1, 101), 10)
        df['pixel_meta_cluster'] = np.repeat(np.arange(1, 11), 100)
        io_utils.write_dataframe(df, os.path.join(base, 'pixel_mat_data', fov + '.feather'))
    for col, k, cnt in (('pixel_som_cluster', 100, 20), ('pixel_meta_cluster', 10, 200)):
        with pytest.warns(UserWarning, match='Provided num_fovs_subset'):
            out = pixel_cluster_utils.compute_pixel_cluster_channel_avg(
                ['fov0', 'fov1'], chans, base, col, k, 'pixel_mat_data', keep_count=True)
        assert list(out[col]) == list(range(1, k + 1))
        assert np.all(out['count'] == cnt)
        assert np.all(np.round(out[chans].values, 1) == [0.1, 0.2, 0.3])
    with pytest.raises(ValueError):
        pixel_cluster_utils.compute_pixel_cluster_channel_avg(
            ['fov0', 'fov1'], chans, base, 'bad_cluster_col', 100, 'pixel_mat_data')
    with pytest.raises(ValueError, match='Averaged data contains just'):
        pixel_cluster_utils.compute_pixel_cluster_channel_avg(
            ['fov0', 'fov1'], chans, base, 'pixel_som_cluster', 1000, 'pixel_mat_data',
            num_fovs_subset=1)


# ------------------------------------------------------------------------------------------------
# cell twin (reference cell_som_clustering_test.py, cluster_helpers_test.py:423-517)
# ------------------------------------------------------------------------------------------------
def make_cell_table(n_per_fov=500, ncols=15, seed=0):
    r = np.random.default_rng(seed)
    cols = [f'pixel_som_cluster_{i}' for i in range(ncols)]
    df = pd.DataFrame(r.random((3 * n_per_fov, ncols)), columns=cols)
    df.iloc[::7, 3] = 0.0  # zeros are ignored by the 99.9 % normalisation
    df['fov'] = np.repeat(['fov0', 'fov1', 'fov2'], n_per_fov)
    df['segmentation_label'] = np.tile(np.arange(1, n_per_fov + 1), 3)
    df['cell_size'] = r.integers(50, 500, 3 * n_per_fov)
    return df, cols


@pytest.mark.parametrize("normalize", [True, False])
def test_train_cell_som_and_cluster_cells(tmp_path, capsys, normalize):
    base = str(tmp_path)
    df, cols = make_cell_table()
    table = os.path.join(base, 'cell_table.csv')
    df.to_csv(table, index=False)
    pysom = cell_som_clustering.train_cell_som(['fov0', 'fov1'], base, table, cols, df.copy(),
                                               normalize=normalize)
    assert "Training SOM" in capsys.readouterr().out
    assert os.path.exists(os.path.join(base, 'cell_som_weights.feather'))
    assert pysom.weights.shape == (100, 15)
    io_utils.verify_same_elements(enforce_order=True, w=pysom.weights.columns.values, c=cols)
    assert pysom.cell_data.shape[0] == 1000  # fov2 filtered out
    with pytest.warns(UserWarning, match='Cell SOM already trained on specified columns'):
        pysom.train_som()

    out = cell_som_clustering.cluster_cells(base, pysom, cols)
    assert "Mapping cell data to SOM cluster labels" in capsys.readouterr().out
    labels = out['cell_som_cluster'].to_numpy()
    assert labels.min() >= 1 and labels.max() <= 100
    X = np.ascontiguousarray(pysom.cell_data[cols].to_numpy(), np.float32)
    ref, _ = oracle.map_data_to_nodes_f32(pysom.weights.values.astype(np.float32), X)
    np.testing.assert_array_equal(labels, ref)
    # idempotent without overwrite, reassigns with it
    again = cell_som_clustering.cluster_cells(base, pysom, cols)
    assert "SOM clusters already assigned to each cell" in capsys.readouterr().out
    np.testing.assert_array_equal(again['cell_som_cluster'].to_numpy(), labels)
    over = cell_som_clustering.cluster_cells(base, pysom, cols, overwrite=True)
    assert "Overwrite flag set, reassigning SOM cluster labels" in capsys.readouterr().out
    np.testing.assert_array_equal(over['cell_som_cluster'].to_numpy(), labels)

    cell_som_clustering.generate_som_avg_files(base, out, cols, 'cell_som_avg.csv')
    avg = pd.read_csv(os.path.join(base, 'cell_som_avg.csv'))
    assert avg['count'].sum() == 1000
    k = int(avg['cell_som_cluster'].iloc[0])
    np.testing.assert_allclose(avg[cols].iloc[0].values,
                               out.loc[out['cell_som_cluster'] == k, cols].mean().values)
    cell_som_clustering.generate_som_avg_files(base, out, cols, 'cell_som_avg.csv')
    assert "Already generated average expression file" in capsys.readouterr().out
