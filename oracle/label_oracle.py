"""CPU oracle for the consumers of the label array -- TEST
INFRASTRUCTURE ONLY.  numpy / pandas restatements of

* the (segmentation label, pixel cluster) count table of ``create_c2pc_data``
  (/root/reference/src/ark/phenotyping/cell_cluster_utils.py:93-190), and
* the cluster mask of ``generate_pixel_cluster_mask``
  (/root/reference/src/ark/utils/data_utils.py:523-551).

PINNED: ``tests/test_label_oracle.py`` checks these against the known answers the reference's own
tests hold (tests/phenotyping/cell_cluster_utils_test.py:103-300).  The arithmetic is integer
counting, so everything downstream is compared bit-exactly."""
import os
import warnings

import numpy as np
import pandas as pd
import pyarrow.feather as paf


def label_histogram(seg, clu, n_seg, n_clu):
    """counts[s, c] = #pixels with label s and cluster c; pairs outside the table are skipped."""
    seg = np.asarray(seg).astype(np.int64)
    clu = np.asarray(clu).astype(np.int64)
    ok = (seg >= 0) & (seg < n_seg) & (clu >= 0) & (clu < n_clu)
    flat = np.bincount(seg[ok] * n_clu + clu[ok], minlength=n_seg * n_clu)
    return flat.reshape(n_seg, n_clu).astype(np.int32), int((~ok).sum())


def pixel_cluster_mask(row_index, column_index, clusters, id_map, H, W):
    """data_utils.py:523-551: int16 zeros, flat fancy assignment (the last duplicate wins)."""
    img = np.zeros((H, W), dtype='int16')
    flat = img.ravel()
    where = np.asarray(row_index) * W + np.asarray(column_index)
    flat[where] = [id_map[int(k)] for k in np.asarray(clusters).astype(int)]
    return flat.reshape(H, W)


def fov_cluster_counts(fov_pixel_data, pixel_cluster_col):
    """cell_cluster_utils.py:114-139: groupby size + pivot, columns renamed ``<col>_<cluster>``."""
    if "segmentation_label" in fov_pixel_data.columns:
        fov_pixel_data = fov_pixel_data.rename(columns={"segmentation_label": "label"})
    sizes = fov_pixel_data.groupby(['label', pixel_cluster_col]).size().reset_index(name='count')
    if sizes[pixel_cluster_col].dtype == float:
        sizes[pixel_cluster_col] = sizes[pixel_cluster_col].astype(int)
    table = sizes.pivot(index='label', columns=pixel_cluster_col, values='count').fillna(0).astype(int)
    table.columns = ['%s_' % pixel_cluster_col + str(c) for c in table.columns]
    return table


def create_c2pc_data(fovs, pixel_data_path, cell_table_path,
                     pixel_cluster_col='pixel_meta_cluster_rename'):
    """cell_cluster_utils.py:63-192 with pandas throughout (the reference's own route)."""
    if pixel_cluster_col not in ('pixel_som_cluster', 'pixel_meta_cluster_rename'):
        raise ValueError(pixel_cluster_col)
    cells = pd.read_csv(cell_table_path)
    for need in ('fov', 'label', 'cell_size'):
        if need not in cells.columns:
            raise ValueError(need)
    cells = cells[['fov', 'label', 'cell_size']]
    cells['label'] = cells['label'].astype(int)
    cells = cells[cells['fov'].isin(fovs)]
    for fov in fovs:
        per_label = fov_cluster_counts(paf.read_feather(os.path.join(pixel_data_path, fov + '.feather')),
                                       pixel_cluster_col)
        mine = cells['fov'] == fov
        both = list(set(list(cells[mine]['label'])).intersection(list(per_label.index.values)))
        per_label = per_label.loc[both]
        rows = pd.Index(cells[mine & cells['label'].isin(both)].index.values)
        cells = cells.combine_first(per_label.set_index(rows))
    cells = cells.fillna(0)
    ccols = [c for c in cells.columns if '%s_' % pixel_cluster_col in c]
    cells = cells[cells[ccols].sum(axis=1) != 0]
    norm = cells.copy()
    norm[ccols] = norm[ccols].div(norm['cell_size'], axis=0)
    cells, norm = cells.reset_index(drop=True), norm.reset_index(drop=True)
    dead = list(norm[ccols].columns[(norm[ccols] == 0).all()].values)
    if dead:
        warnings.warn('Pixel clusters %s do not appear in any cells, removed from analysis' %
                      ','.join(dead))
        cells, norm = cells.drop(columns=dead), norm.drop(columns=dead)
    return cells, norm
