"""Cell SOM drivers for the notebook-3 path (``train_cell_som``, ``cluster_cells``,
``generate_som_avg_files``).  Interface -- argument names, defaults, printed messages, errors and
output files -- follows ``/root/reference/src/ark/phenotyping/cell_som_clustering.py`` (:8-75,
:78-139, :142-191) so the cell-clustering notebook cells run unchanged; the work itself is done by
``cluster_helpers.CellSOMCluster`` on the B200 kernels."""
import os

from . import cell_cluster_utils, io_utils
from .cluster_helpers import CellSOMCluster

# columns of the cell table that are never SOM inputs
_META_COLS = ('fov', 'label', 'cell_size', 'cell_som_cluster')


def train_cell_som(fovs, base_dir, cell_table_path, cell_som_cluster_cols,
                   cell_som_input_data, som_weights_name='cell_som_weights.feather',
                   xdim=10, ydim=10, lr_start=0.05, lr_end=0.01, num_passes=1, seed=42,
                   overwrite=False, normalize=True):
    """Train the cell SOM on ``cell_som_cluster_cols`` of ``cell_som_input_data`` and save the
    weights to ``base_dir/som_weights_name``.  Returns the ``CellSOMCluster``."""
    io_utils.validate_paths([cell_table_path])
    io_utils.verify_in_list(provided_cluster_cols=cell_som_cluster_cols,
                            som_input_cluster_cols=cell_som_input_data.columns.values)
    hyper = dict(num_passes=num_passes, xdim=xdim, ydim=ydim, lr_start=lr_start, lr_end=lr_end,
                 seed=seed, normalize=normalize)
    cell_pysom = CellSOMCluster(cell_som_input_data, os.path.join(base_dir, som_weights_name), fovs,
                                cell_som_cluster_cols, **hyper)
    print("Training SOM")
    cell_pysom.train_som(overwrite=overwrite)
    return cell_pysom


def cluster_cells(base_dir, cell_pysom, cell_som_cluster_cols, num_parallel_cells=1000000,
                  overwrite=False):
    """Assign SOM labels to every cell of ``cell_pysom.cell_data``; returns the labelled table."""
    if cell_pysom.weights is None:
        raise ValueError("Using untrained cell_pysom object, please invoke train_cell_som first")

    table = cell_pysom.cell_data
    if "segmentation_label" in table.columns:
        table.rename(columns={"segmentation_label": "label"}, inplace=True)

    if 'cell_som_cluster' in table.columns:
        if not overwrite:
            print("SOM clusters already assigned to each cell")
            return table
        print("Overwrite flag set, reassigning SOM cluster labels")

    # every weights column must be one of the table's input columns
    inputs = [c for c in table.columns if c not in _META_COLS]
    io_utils.verify_in_list(cell_weights_columns=cell_pysom.weights.columns.values,
                            cell_som_input_data_columns=inputs)

    print("Mapping cell data to SOM cluster labels")
    return cell_pysom.assign_som_clusters(num_parallel_cells)


def generate_som_avg_files(base_dir, cell_som_input_data, cell_som_cluster_cols,
                           cell_som_expr_col_avg_name, overwrite=False):
    """Write the per-SOM-cluster averages of the training columns (with counts) to CSV."""
    if 'cell_som_cluster' not in cell_som_input_data.columns:
        raise ValueError('cell_som_input_data does not have SOM labels assigned')

    target = os.path.join(base_dir, cell_som_expr_col_avg_name)
    if os.path.exists(target) and not overwrite:
        print("Already generated average expression file for each cell SOM column, skipping")
        return
    if os.path.exists(target):
        print("Overwrite flag set, regenerating average expression file for cell SOM clusters")

    print("Computing the average value of each training column specified per cell SOM cluster")
    cell_cluster_utils.compute_cell_som_cluster_cols_avg(
        cell_som_input_data, cell_som_cluster_cols, 'cell_som_cluster', keep_count=True
    ).to_csv(target, index=False)
