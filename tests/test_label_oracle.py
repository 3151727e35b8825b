"""CPU tests of the N4 oracle (oracle/label_oracle.py): pinned against the known answers the
reference's own tests hold, plus the numpy histogram / mask restatements against brute force."""
import tempfile

import numpy as np
import pandas as pd
import pytest

from oracle import label_oracle as LO
import label_fixtures as LF


def test_create_c2pc_data_known_answers_of_the_reference_test(rng):
    with tempfile.TemporaryDirectory() as d:
        fovs, pix, cells = LF.c2pc_case(d, rng)
        with pytest.raises(ValueError):
            LO.create_c2pc_data(fovs, 'consensus', 'cell_table', pixel_cluster_col='bad_col')
        counts, norm = LO.create_c2pc_data(fovs, pix, cells, pixel_cluster_col='pixel_som_cluster')
        som_cols = ['pixel_som_cluster_%d' % i for i in range(3)]
        assert set(som_cols) <= set(counts.columns)
        np.testing.assert_array_equal(counts[som_cols].values, np.array(LF.C2PC_SOM))
        np.testing.assert_array_equal(norm[som_cols].values, np.array(LF.C2PC_SOM) / 5)
        counts, norm = LO.create_c2pc_data(fovs, pix, cells,
                                           pixel_cluster_col='pixel_meta_cluster_rename')
        meta_cols = ['pixel_meta_cluster_rename_%d' % i for i in range(2)]
        np.testing.assert_array_equal(counts[meta_cols].values, np.array(LF.C2PC_META))
        np.testing.assert_array_equal(norm[meta_cols].values, np.array(LF.C2PC_META) / 5)
        # the cell of each FOV without any assigned pixel is dropped (12 cells -> 10 rows)
        assert len(counts) == 10 and list(counts['label']) == [0, 1, 2, 3, 4] * 2


def test_histogram_restatement_matches_brute_force_and_the_pandas_route(rng):
    n, n_seg, n_clu = 5000, 37, 11
    seg = rng.integers(-1, n_seg + 1, n)
    clu = rng.integers(-1, n_clu + 1, n)
    counts, bad = LO.label_histogram(seg, clu, n_seg, n_clu)
    ref = np.zeros((n_seg, n_clu), np.int32)
    nbad = 0
    for s, c in zip(seg, clu):
        if 0 <= s < n_seg and 0 <= c < n_clu:
            ref[s, c] += 1
        else:
            nbad += 1
    np.testing.assert_array_equal(counts, ref)
    assert bad == nbad
    # the pandas route of the reference on the same pixels (NaN = no cluster)
    ok = (seg >= 0) & (seg < n_seg)
    df = pd.DataFrame({'label': seg[ok], 'pixel_som_cluster':
                       np.where((clu[ok] >= 0) & (clu[ok] < n_clu), clu[ok], np.nan)})
    table = LO.fov_cluster_counts(df, 'pixel_som_cluster')
    rows, cols = np.flatnonzero(ref.sum(1) > 0), np.flatnonzero(ref.sum(0) > 0)
    np.testing.assert_array_equal(table.values, ref[np.ix_(rows, cols)])
    assert list(table.index) == list(rows)
    assert list(table.columns) == ['pixel_som_cluster_%d' % c for c in cols]


def test_mask_restatement_last_duplicate_wins(rng):
    H, W = 40, 40
    r, c = rng.integers(0, H, 100), rng.integers(0, W, 100)
    k = np.tile(np.arange(1, 11), 10)
    id_map = {i: (i + 1) // 2 for i in range(1, 11)}
    img = LO.pixel_cluster_mask(r, c, k, id_map, H, W)
    ref = np.zeros((H, W), np.int16)
    for i in range(100):
        ref[r[i], c[i]] = id_map[k[i]]
    np.testing.assert_array_equal(img, ref)
    assert img.dtype == np.int16 and img.max() <= 5
