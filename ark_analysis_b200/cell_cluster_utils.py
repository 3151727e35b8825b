"""``compute_cell_som_cluster_cols_avg`` -- the one helper of ``cell_cluster_utils`` the cell SOM
driver calls (reference cell_cluster_utils.py:10-60).  cells x features tables are small; this
stays a pandas group-by."""
import numpy as np

from . import io_utils


def compute_cell_som_cluster_cols_avg(cell_cluster_data, cell_som_cluster_cols,
                                      cell_cluster_col, keep_count=False):
    """Mean of every training column per cell SOM (or meta) cluster, optionally with counts."""
    io_utils.verify_in_list(provided_cluster_col=cell_cluster_col,
                            valid_cluster_cols=['cell_som_cluster', 'cell_meta_cluster'])
    io_utils.verify_in_list(provided_cluster_col=cell_som_cluster_cols,
                            cluster_data_valid_cols=cell_cluster_data.columns.values)
    sub = cell_cluster_data.loc[:, list(cell_som_cluster_cols) + [cell_cluster_col]]
    grouped = sub.groupby(cell_cluster_col)
    out = grouped.mean().reset_index()
    out[cell_cluster_col] = out[cell_cluster_col].astype(np.int64)
    if keep_count:
        out['count'] = grouped.size().to_numpy()
    return out
