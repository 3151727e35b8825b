"""Pixel SOM drivers -- mirror of ``/root/reference/src/ark/phenotyping/pixel_som_clustering.py``
(``train_pixel_som`` :16-90, ``run_pixel_som_assignment`` :93-136, ``cluster_pixels`` :139-289,
``generate_som_avg_files`` :308-371): same signatures, defaults, printed messages, errors, files
and the ``_temp`` restart protocol, so ``templates/2_Pixie_Cluster_Pixels.ipynb`` cells 32 and 35
run on it unchanged.  The arithmetic runs on the B200 kernels.

One deliberate difference: ``multiprocess=True`` does not spawn processes (a process pool would
build one CUDA context per worker and pickle the training table to each).  FOVs are handed to a
thread pool of ``batch_size`` workers instead: Feather reads and writes overlap, every worker thread
enqueues its kernels on its own workspace (``som._workspace``), and ``som_clusters_seen`` is
updated in this process (the reference loses those updates in its child processes).
"""
import os
from concurrent.futures import ThreadPoolExecutor
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


class _LabelRun:
    """One pass of SOM label assignment over the FOV files of ``data_path``.

    The labelled tables are written to ``data_path + '_temp'``, which replaces ``data_path`` when
    the pass is over: an interrupted pass leaves both directories behind, and the next call only
    labels the FOVs that have no ``pixel_som_cluster`` column yet (the reference's restart
    protocol, pixel_som_clustering.py:220-289, pixel_cluster_utils.py:419-478)."""

    def __init__(self, base_dir, data_dir, pixel_pysom, overwrite, num_parallel_pixels):
        self.base_dir, self.data_dir = base_dir, data_dir
        self.data_path = os.path.join(base_dir, data_dir)
        self.temp_path = self.data_path + '_temp'
        self.pysom = pixel_pysom
        self.overwrite = overwrite
        self.num_parallel_pixels = num_parallel_pixels
        self.done = 0

    def pending(self, fovs):
        """FOVs of the master list still to label, sorted (every run sees the same order)."""
        if self.overwrite:
            print('Overwrite flag set, reassigning SOM cluster labels to all FOVs')
            self.pysom.som_clusters_seen = set()
            os.mkdir(self.temp_path)
            candidates = io_utils.remove_file_extensions(
                io_utils.list_files(self.data_path, substrs='.feather'))
        else:
            candidates = pixel_cluster_utils.find_fovs_missing_col(
                self.base_dir, self.data_dir, 'pixel_som_cluster')
        return sorted(set(candidates) & set(fovs))

    def label(self, fov):
        return run_pixel_som_assignment(self.data_path, self.pysom, self.overwrite,
                                        self.num_parallel_pixels, fov)

    def account(self, results, total, every=None):
        """Book a batch of ``(fov, status)`` results: corrupted FOVs are reported and not counted;
        progress is printed per batch, or every ``every`` FOVs and at the end."""
        for fov, status in results:
            if status == 1:
                print("The data for FOV %s has been corrupted, skipping" % fov)
            else:
                self.done += 1
        if every is None or self.done % every == 0 or self.done == total:
            print("Processed %d fovs" % self.done)

    def commit(self):
        rmtree(self.data_path, onerror=_ignore_extended_attributes)
        move(self.temp_path, self.data_path)


def cluster_pixels(fovs, base_dir, pixel_pysom, data_dir='pixel_mat_data',
                   multiprocess=False, batch_size=5, num_parallel_pixels=1000000,
                   overwrite=False):
    """Assign SOM cluster labels to the full pixel data of every FOV; the labelled (and
    normalised) tables replace the files in ``data_dir``."""
    run = _LabelRun(base_dir, data_dir, pixel_pysom, overwrite, num_parallel_pixels)
    io_utils.validate_paths([run.data_path])
    if pixel_pysom.weights is None:
        raise ValueError("Using untrained pixel_pysom object, please invoke train_pixel_som first")

    data_files = io_utils.list_files(run.data_path, substrs='.feather')
    io_utils.verify_in_list(provided_fovs=fovs,
                            subsetted_fovs=io_utils.remove_file_extensions(data_files))

    # norm values, weights and data must agree on the channel columns AND their order
    channel_cols = _sample_fov_columns(base_dir, data_dir, data_files)
    for name, frame in (("norm_vals_columns", pixel_pysom.norm_data),
                        ("pixel_som_weights_columns", pixel_pysom.weights)):
        io_utils.verify_same_elements(enforce_order=True, **{name: frame.columns.values},
                                      pixel_data_columns=channel_cols)

    todo = run.pending(fovs)
    if not todo:
        print("There are no more FOVs to assign SOM labels to, skipping")
        return
    if len(todo) < len(fovs):
        print("Restarting SOM label assignment from fov %s, "
              "%d fovs left to process" % (todo[0], len(todo)))

    print("Mapping pixel data to SOM cluster labels")
    if multiprocess:
        # a thread pool, not processes (module docstring): one batch of FOVs in flight at a time
        width = max(1, int(batch_size))
        with ThreadPoolExecutor(max_workers=width) as pool:
            for lo in range(0, len(todo), width):
                run.account(pool.map(run.label, todo[lo:lo + width]), len(todo))
    else:
        for fov in todo:
            run.account([run.label(fov)], len(todo), every=10)
    run.commit()


def _ignore_extended_attributes(func: Callable, filename: str, exc_info: Tuple[Any, Any, Any]):
    """rmtree error handler: tolerate macOS extended-attribute files ("._*") that vanish."""
    if func is os.unlink and os.path.basename(filename).startswith("._"):
        return
    if isinstance(exc_info[1], BaseException):
        raise exc_info[1]
    raise  # called outside an error (the reference's test does): RuntimeError, as there


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
