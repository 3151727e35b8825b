"""Builders of the small on-disk cases the N4 tests share (the layout of the reference's
tests/phenotyping/cell_cluster_utils_test.py:103-176 and tests/utils/data_utils_test.py:391-443)."""
import os

import numpy as np
import pandas as pd
import pyarrow.feather as paf


def c2pc_case(temp_dir, rng):
    """The reference's create_c2pc_data test case: 2 FOVs x 6 cells x 10 pixels, one cell per FOV
    with no cluster assigned (NaN), SOM clusters 0-1 in fov1 and 1-2 in fov2."""
    chans = ['chan1', 'chan2', 'chan3']
    cell_table = pd.DataFrame(rng.random((12, 3)), columns=chans)
    cell_table.loc[0:5, 'fov'] = 'fov1'
    cell_table.loc[6:11, 'fov'] = 'fov2'
    cell_table.loc[0:5, 'label'] = np.arange(6)
    cell_table.loc[6:11, 'label'] = np.arange(6)
    cell_table['cell_size'] = 5
    cell_table_path = os.path.join(temp_dir, 'cell_table_size_normalized.csv')
    cell_table.to_csv(cell_table_path, index=False)
    pixel_data_path = os.path.join(temp_dir, 'pixel_data_path')
    os.mkdir(pixel_data_path)
    for fov in ['fov1', 'fov2']:
        t = pd.DataFrame(rng.random((60, 3)), columns=chans)
        t['fov'] = fov
        t['label'] = np.repeat(np.arange(6), 10)
        lo = 0 if fov == 'fov1' else 1
        t['pixel_som_cluster'] = np.concatenate((np.repeat(np.arange(lo, lo + 2), 25),
                                                 np.repeat(np.nan, 10)))
        t['pixel_meta_cluster_rename'] = np.concatenate((np.repeat(np.arange(2), 25),
                                                         np.repeat(np.nan, 10)))
        paf.write_feather(t, os.path.join(pixel_data_path, fov + '.feather'),
                          compression='uncompressed')
    return ['fov1', 'fov2'], pixel_data_path, cell_table_path


# known answers held by the reference's test (cell_cluster_utils_test.py:192-201, :226-235)
C2PC_SOM = [[10, 0, 0], [10, 0, 0], [5, 5, 0], [0, 10, 0], [0, 10, 0],
            [0, 10, 0], [0, 10, 0], [0, 5, 5], [0, 0, 10], [0, 0, 10]]
C2PC_META = [[10, 0], [10, 0], [5, 5], [0, 10], [0, 10],
             [10, 0], [10, 0], [5, 5], [0, 10], [0, 10]]


def big_c2pc_case(temp_dir, rng, nfov=2, hw=256, ncell=300, ncluster=100, missing=0.05):
    """FOV-sized case: hw x hw pixels per FOV in image order, blob-like cells, SOM clusters 1..K,
    a fraction of pixels without a cluster, segmentation label 0 = background, cells absent from
    the cell table and cell-table cells without pixels."""
    fovs = ['fov%d' % i for i in range(nfov)]
    pixel_data_path = os.path.join(temp_dir, 'pixel_mat_data')
    os.mkdir(pixel_data_path)
    rows = []
    for fov in fovs:
        yy, xx = np.mgrid[0:hw, 0:hw]
        cy, cx = rng.integers(0, hw, ncell), rng.integers(0, hw, ncell)
        d = (yy[..., None] - cy) ** 2 + (xx[..., None] - cx) ** 2
        seg = np.where(d.min(-1) < 36, d.argmin(-1) + 1, 0).ravel()
        clu = rng.integers(1, ncluster + 1, hw * hw).astype(np.float64)
        clu[rng.random(hw * hw) < missing] = np.nan
        t = pd.DataFrame({'chan0': rng.random(hw * hw), 'fov': fov,
                          'row_index': yy.ravel(), 'column_index': xx.ravel(),
                          'segmentation_label': seg, 'pixel_som_cluster': clu})
        paf.write_feather(t, os.path.join(pixel_data_path, fov + '.feather'),
                          compression='uncompressed')
        labels = np.unique(seg[seg > 0])
        keep = labels[rng.random(labels.size) < 0.9]                 # some cells not in the table
        extra = np.arange(ncell + 5, ncell + 9)                      # table cells without pixels
        for lab in np.concatenate((keep, extra)):
            rows.append({'fov': fov, 'label': float(lab), 'cell_size': int(rng.integers(20, 200))})
    cell_table_path = os.path.join(temp_dir, 'cell_table.csv')
    pd.DataFrame(rows).to_csv(cell_table_path, index=False)
    return fovs, pixel_data_path, cell_table_path
