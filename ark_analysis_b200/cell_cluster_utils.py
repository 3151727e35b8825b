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


# ------------------------------------------------------------------------------------------------
# N4: per-cell counts of pixel clusters (reference cell_cluster_utils.py:63-192)
# ------------------------------------------------------------------------------------------------
def _device_int_column(column, torch, device):
    """One Arrow column (int or float, NaN/null = "no value") as a CUDA int32 tensor, -1 for
    missing values (the reference's groupby drops them)."""
    import pyarrow as pa
    import pyarrow.compute as pc
    if isinstance(column, pa.ChunkedArray):
        column = column.combine_chunks()
    if column.null_count:
        column = pc.fill_null(column, float('nan') if pa.types.is_floating(column.type) else -1)
    t = torch.from_numpy(np.array(column.to_numpy(zero_copy_only=False))).to(device)
    if t.is_floating_point():
        t = torch.where(torch.isnan(t), torch.full_like(t, -1.0), t)
    return t.to(torch.int32)


def fov_cluster_counts(fov_path, pixel_cluster_col, device=None):
    """``num_cluster_per_seg_label`` of the reference (cell_cluster_utils.py:119-139) for one FOV
    file: a DataFrame indexed by segmentation label (ascending), one ``<col>_<cluster>`` column per
    cluster that occurs (ascending), values = pixel counts.  The (label, cluster) histogram over
    the FOV's pixels is one pass of ``pixie_label_histogram_i32`` on the GPU; only the small
    [labels x clusters] table comes back."""
    import pandas as pd
    import pyarrow as pa
    import torch

    from . import som
    dev = torch.device(device) if device is not None else som._default_device()
    with pa.memory_map(str(fov_path)) as src:
        schema = pa.ipc.open_file(src).schema
    label_col = 'segmentation_label' if 'segmentation_label' in schema.names else 'label'
    table = io_utils.read_table(fov_path, columns=[label_col, pixel_cluster_col])
    seg = _device_int_column(table.column(label_col), torch, dev)
    clu = _device_int_column(table.column(pixel_cluster_col), torch, dev)
    name = '%s_' % pixel_cluster_col
    if seg.numel() == 0 or int(seg.max()) < 0 or int(clu.max()) < 0:
        return pd.DataFrame(index=pd.Index([], dtype=np.int64, name='label'))
    n_seg, n_clu = int(seg.max()) + 1, int(clu.max()) + 1
    counts, _ = som.label_histogram(seg, clu, n_seg, n_clu)
    counts = counts.cpu().numpy().astype(np.int64)
    rows = np.flatnonzero(counts.sum(axis=1) > 0)
    cols = np.flatnonzero(counts.sum(axis=0) > 0)
    out = pd.DataFrame(counts[np.ix_(rows, cols)], index=pd.Index(rows, name='label'),
                       columns=[name + str(c) for c in cols])
    return out


def create_c2pc_data(fovs, pixel_data_path, cell_table_path,
                     pixel_cluster_col='pixel_meta_cluster_rename', device=None):
    """cell x pixel-cluster count table and its ``cell_size``-normalised twin, as the reference's
    ``create_c2pc_data`` returns them (same rows, columns, order and dtypes).  Per FOV the pixel
    table never becomes a DataFrame: its (label, cluster) columns go to the GPU histogram
    (``fov_cluster_counts``); what follows operates on cells x clusters tables with the pandas
    calls whose alignment rules define the reference's output (combine_first)."""
    import os
    import warnings

    import pandas as pd
    io_utils.verify_in_list(provided_cluster_col=[pixel_cluster_col],
                            valid_cluster_cols=['pixel_som_cluster', 'pixel_meta_cluster_rename'])
    cell_table = pd.read_csv(cell_table_path)
    io_utils.verify_in_list(required_cell_table_cols=['fov', 'label', 'cell_size'],
                            provided_cell_table_cols=cell_table.columns.values)
    cell_table = cell_table[['fov', 'label', 'cell_size']]
    cell_table['label'] = cell_table['label'].astype(int)
    cell_table = cell_table[cell_table['fov'].isin(fovs)]

    for fov in fovs:
        fov_counts = fov_cluster_counts(os.path.join(pixel_data_path, fov + '.feather'),
                                        pixel_cluster_col, device=device)
        in_fov = cell_table['fov'] == fov
        # cells of this FOV that own pixels: the reference subsets the count table by the SET
        # intersection of the labels and re-indexes it, positionally, with the cell table's index
        shared = list(set(cell_table[in_fov]['label']).intersection(list(fov_counts.index.values)))
        fov_counts = fov_counts.loc[shared]
        target = pd.Index(cell_table[in_fov & cell_table['label'].isin(shared)].index.values)
        cell_table = cell_table.combine_first(fov_counts.set_index(target))

    cell_table = cell_table.fillna(0)
    count_cols = [c for c in cell_table.columns if '%s_' % pixel_cluster_col in c]
    cell_table = cell_table[cell_table[count_cols].sum(axis=1) != 0]
    cell_table_norm = cell_table.copy()
    cell_table_norm[count_cols] = cell_table_norm[count_cols].div(cell_table_norm['cell_size'],
                                                                  axis=0)
    cell_table = cell_table.reset_index(drop=True)
    cell_table_norm = cell_table_norm.reset_index(drop=True)
    empty = list(cell_table_norm[count_cols].columns[(cell_table_norm[count_cols] == 0).all()].values)
    if len(empty) > 0:
        warnings.warn('Pixel clusters %s do not appear in any cells, removed from analysis' %
                      ','.join(empty))
        cell_table = cell_table.drop(columns=empty)
        cell_table_norm = cell_table_norm.drop(columns=empty)
    return cell_table, cell_table_norm
