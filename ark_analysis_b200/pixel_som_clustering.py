"""Pixel SOM drivers -- mirror of ``/root/reference/src/ark/phenotyping/pixel_som_clustering.py``
(``train_pixel_som`` :16-90, ``run_pixel_som_assignment`` :93-136, ``cluster_pixels`` :139-289,
``generate_som_avg_files`` :308-371): same signatures, defaults, printed messages, errors, files
and the ``_temp`` restart protocol, so ``templates/2_Pixie_Cluster_Pixels.ipynb`` cells 32 and 35
run on it unchanged.  The arithmetic runs on the B200 kernels.

One deliberate difference: ``multiprocess=True`` does not spawn processes (a process pool would
build one CUDA context per worker and pickle the training table to each).  FOVs are handed to a
thread pool of ``batch_size`` workers instead: Feather reads and writes overlap, GPU calls are
serialised by the library, and ``som_clusters_seen`` is updated in this process (the reference
loses those updates in its child processes).
"""
import os
from concurrent.futures import ThreadPoolExecutor
from functools import partial
from shutil import move, rmtree
from typing import Any, Callable, Tuple

from . import cluster_helpers, io_utils, pixel_cluster_utils
from .io_utils import ArrowInvalid


def train_pixel_som(fovs, channels, base_dir,
                    subset_dir='pixel_mat_subsetted',
                    norm_vals_name='post_rowsum_chan_norm.feather',
                    som_weights_name='pixel_som_weights.feather', xdim=10, ydim=10,
                    lr_start=0.05, lr_end=0.01, num_passes=1, seed=42,
                    overwrite=False):
    """Train the pixel SOM on the subsetted pixel data and save the weights to
    ``base_dir/som_weights_name``.  Returns the ``PixelSOMCluster``."""
    subsetted_path = os.path.join(base_dir, subset_dir)
    norm_vals_path = os.path.join(base_dir, norm_vals_name)
    som_weights_path = os.path.join(base_dir, som_weights_name)

    # the weights may or may not exist yet; PixelSOMCluster deals with that
    io_utils.validate_paths([subsetted_path, norm_vals_path])

    files = io_utils.list_files(subsetted_path, substrs='.feather')
    io_utils.verify_in_list(provided_fovs=fovs,
                            subsetted_fovs=io_utils.remove_file_extensions(files))

    sample_cols = io_utils.read_table(os.path.join(subsetted_path, files[0])).column_names
    io_utils.verify_in_list(provided_channels=channels, subsetted_channels=sample_cols)

    pixel_pysom = cluster_helpers.PixelSOMCluster(
        subsetted_path, norm_vals_path, som_weights_path, fovs, channels,
        num_passes=num_passes, xdim=xdim, ydim=ydim, lr_start=lr_start, lr_end=lr_end,
        seed=seed)

    print("Training SOM")
    pixel_pysom.train_som(overwrite=overwrite)
    return pixel_pysom


def run_pixel_som_assignment(pixel_data_path, pixel_pysom_obj, overwrite, num_parallel_pixels, fov):
    """Label one FOV and write it to ``pixel_data_path + '_temp'``.  Returns ``(fov, status)``,
    status 1 meaning the FOV's file could not be read (it is then skipped and dropped)."""
    fov_path = os.path.join(pixel_data_path, fov + '.feather')
    temp_path = os.path.join(pixel_data_path + '_temp', fov + '.feather')

    # Arrow-native path (N2): column buffers straight to the device, one normalise + cast +
    # transpose kernel, the labelled table written back without a DataFrame round trip
    fast = getattr(pixel_pysom_obj, 'assign_som_clusters_table', None)
    if fast is not None and num_parallel_pixels > 0:
        try:
            table = io_utils.read_table(fov_path)
        except (ArrowInvalid, OSError, IOError):
            return fov, 1
        if overwrite and 'pixel_som_cluster' in table.column_names:
            table = table.drop_columns(['pixel_som_cluster'])
        labelled = fast(table, normalize_data=not overwrite)
        if labelled is not None:
            io_utils.write_table(labelled, temp_path, compression='uncompressed')
            return fov, 0

    try:
        fov_data = io_utils.read_dataframe(fov_path)
    except (ArrowInvalid, OSError, IOError):
        return fov, 1

    # stored data is already normalised once labels were written: do not normalise twice
    if overwrite:
        fov_data = fov_data.drop(columns="pixel_som_cluster", errors="ignore")

    fov_data = pixel_pysom_obj.assign_som_clusters(
        fov_data, normalize_data=not overwrite, num_parallel_pixels=num_parallel_pixels)

    io_utils.write_dataframe(fov_data, temp_path, compression='uncompressed')
    return fov, 0


def _sample_fov_columns(base_dir, data_dir, data_files):
    """Channel columns of the first readable FOV file (metadata and label columns removed)."""
    sample_cols = None
    for f in data_files:
        try:
            sample_cols = io_utils.read_table(os.path.join(base_dir, data_dir, f)).column_names
            break
        except (ArrowInvalid, OSError, IOError):
            continue
    if sample_cols is None:
        raise ValueError("No readable FOV file found in %s" % os.path.join(base_dir, data_dir))
    meta = {'fov', 'row_index', 'column_index', 'label', 'segmentation_label',
            'pixel_som_cluster', 'pixel_meta_cluster', 'pixel_meta_cluster_rename'}
    return [c for c in sample_cols if c not in meta]


def cluster_pixels(fovs, base_dir, pixel_pysom, data_dir='pixel_mat_data',
                   multiprocess=False, batch_size=5, num_parallel_pixels=1000000,
                   overwrite=False):
    """Assign SOM cluster labels to the full pixel data of every FOV; the labelled (and
    normalised) tables replace the files in ``data_dir``."""
    data_path = os.path.join(base_dir, data_dir)
    io_utils.validate_paths([data_path])

    if pixel_pysom.weights is None:
        raise ValueError("Using untrained pixel_pysom object, please invoke train_pixel_som first")

    data_files = io_utils.list_files(data_path, substrs='.feather')
    io_utils.verify_in_list(provided_fovs=fovs,
                            subsetted_fovs=io_utils.remove_file_extensions(data_files))

    # norm values, weights and data must agree on the channel columns AND their order
    channel_cols = _sample_fov_columns(base_dir, data_dir, data_files)
    io_utils.verify_same_elements(
        enforce_order=True,
        norm_vals_columns=pixel_pysom.norm_data.columns.values,
        pixel_data_columns=channel_cols)
    io_utils.verify_same_elements(
        enforce_order=True,
        pixel_som_weights_columns=pixel_pysom.weights.columns.values,
        pixel_data_columns=channel_cols)

    if overwrite:
        print('Overwrite flag set, reassigning SOM cluster labels to all FOVs')
        pixel_pysom.som_clusters_seen = set()
        os.mkdir(data_path + '_temp')
        fovs_list = io_utils.remove_file_extensions(
            io_utils.list_files(data_path, substrs='.feather'))
    else:
        fovs_list = pixel_cluster_utils.find_fovs_missing_col(
            base_dir, data_dir, 'pixel_som_cluster')

    # only FOVs of the master list; keep a deterministic order
    wanted = set(fovs)
    fovs_list = sorted(f for f in set(fovs_list) if f in wanted)

    if len(fovs_list) == 0:
        print("There are no more FOVs to assign SOM labels to, skipping")
        return

    if len(fovs_list) < len(fovs):
        print("Restarting SOM label assignment from fov %s, "
              "%d fovs left to process" % (fovs_list[0], len(fovs_list)))

    fovs_processed = 0
    fov_data_func = partial(
        run_pixel_som_assignment, data_path, pixel_pysom, overwrite, num_parallel_pixels)

    print("Mapping pixel data to SOM cluster labels")

    if multiprocess:
        with ThreadPoolExecutor(max_workers=max(1, int(batch_size))) as pool:
            for start in range(0, len(fovs_list), batch_size):
                fov_batch = fovs_list[start:start + batch_size]
                for fs in pool.map(fov_data_func, fov_batch):
                    if fs[1] == 1:
                        print("The data for FOV %s has been corrupted, skipping" % fs[0])
                        fovs_processed -= 1
                fovs_processed += len(fov_batch)
                print("Processed %d fovs" % fovs_processed)
    else:
        for fov in fovs_list:
            fov_status = fov_data_func(fov)
            if fov_status[1] == 1:
                print("The data for FOV %s has been corrupted, skipping" % fov_status[0])
                fovs_processed -= 1
            fovs_processed += 1
            if fovs_processed % 10 == 0 or fovs_processed == len(fovs_list):
                print("Processed %d fovs" % fovs_processed)

    # the temp directory becomes the data directory
    rmtree(data_path, onerror=_ignore_extended_attributes)
    move(data_path + '_temp', data_path)


def _ignore_extended_attributes(func: Callable, filename: str, exc_info: Tuple[Any, Any, Any]):
    """rmtree error handler: tolerate macOS extended-attribute files ("._*") that vanish."""
    is_meta_file = os.path.basename(filename).startswith("._")
    if not (func is os.unlink and is_meta_file):
        raise


def generate_som_avg_files(fovs, channels, base_dir, pixel_pysom, data_dir='pixel_data_dir',
                           pc_chan_avg_som_cluster_name='pixel_channel_avg_som_cluster.csv',
                           num_fovs_subset=100, require_all_som_clusters=True, seed=42,
                           overwrite=False):
    """Write the average channel expression (and pixel count) per pixel SOM cluster to CSV."""
    som_cluster_avg_path = os.path.join(base_dir, pc_chan_avg_som_cluster_name)

    if pixel_pysom.weights is None:
        raise ValueError("Using untrained pixel_pysom object, please invoke train_som first")

    if os.path.exists(som_cluster_avg_path):
        if not overwrite:
            print("Already generated SOM cluster channel average file, skipping")
            return
        print("Overwrite flag set, regenerating SOM cluster channel average file")

    print("Computing average channel expression across pixel SOM clusters")
    avgs = pixel_cluster_utils.compute_pixel_cluster_channel_avg(
        fovs, channels, base_dir, 'pixel_som_cluster',
        len(pixel_pysom.som_clusters_seen) if require_all_som_clusters else None,
        data_dir, num_fovs_subset=num_fovs_subset, seed=seed, keep_count=True)
    avgs.to_csv(som_cluster_avg_path, index=False)
