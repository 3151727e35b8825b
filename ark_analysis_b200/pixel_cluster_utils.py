"""The two helpers of ``pixel_cluster_utils`` that sit on the Pixie SOM path:
``find_fovs_missing_col`` (restart protocol, reference pixel_cluster_utils.py:419-478) and
``compute_pixel_cluster_channel_avg`` (SOM-cluster channel averages, reference :294-416).  The rest
of that module (image preprocessing) is out of scope."""
import os
import random
import warnings

import numpy as np
import pandas as pd
import torch

from . import io_utils, som
from .io_utils import ArrowInvalid


def compute_pixel_cluster_channel_avg(fovs, channels, base_dir, pixel_cluster_col,
                                      num_pixel_clusters, pixel_data_dir='pixel_mat_data',
                                      num_fovs_subset=100, seed=42, keep_count=False):
    """Average channel value per pixel cluster over a seeded subsample of FOVs.

    Same contract as the reference (arguments, FOV subsampling with ``random.seed(seed)``, the
    ``count`` column, the error when clusters went missing).  The per-FOV ``groupby().sum()`` and
    ``.size()`` run on the GPU: rows go up as one fp32 matrix and come back as a K x (C+1) table
    of sums and counts (``pixie_cluster_sums_f32``); totals across FOVs are kept in float64."""
    io_utils.verify_in_list(provided_cluster_col=[pixel_cluster_col],
                            valid_cluster_cols=['pixel_som_cluster', 'pixel_meta_cluster'])
    if num_pixel_clusters is not None and num_pixel_clusters <= 0:
        raise ValueError("If set, number of pixel clusters desired must be a positive integer")
    if num_fovs_subset <= 0:
        raise ValueError("Number of fovs to subset must be a positive integer")
    if len(fovs) < num_fovs_subset:
        warnings.warn(
            'Provided num_fovs_subset=%d but only %d FOVs in dataset, '
            'subsetting just the %d FOVs' % (num_fovs_subset, len(fovs), len(fovs)))

    random.seed(seed)
    fovs_sub = random.sample(fovs, num_fovs_subset) if num_fovs_subset < len(fovs) else fovs

    channels = list(channels)
    totals = {}  # cluster id -> float64 [C + 1] (channel sums, count)
    for fov in fovs_sub:
        try:
            fov_data = io_utils.read_dataframe(
                os.path.join(base_dir, pixel_data_dir, fov + '.feather'),
                columns=channels + [pixel_cluster_col])
        except (ArrowInvalid, OSError, IOError):
            print("The data for FOV %s has been corrupted, skipping" % fov)
            continue
        if fov_data.shape[0] == 0:
            continue
        # dense 1..K codes for whatever ids the column holds
        ids, codes = np.unique(fov_data[pixel_cluster_col].to_numpy(), return_inverse=True)
        X = som.to_device_matrix(fov_data[channels].to_numpy())
        labels = torch.from_numpy((codes + 1).astype(np.int32)).to(X.device)
        SN = som.label_sums(X, labels, len(ids)).cpu().numpy()
        for row, cid in zip(SN, ids):
            if cid in totals:
                totals[cid] += row
            else:
                totals[cid] = row.copy()

    if not totals:
        # mirrors pandas.concat([]) in the reference: nothing could be read
        raise ValueError("No objects to concatenate")

    cluster_ids = sorted(totals)
    if num_pixel_clusters is not None and len(cluster_ids) < num_pixel_clusters:
        raise ValueError(
            'Averaged data contains just %d clusters out of %d. '
            'Average expression file not written. '
            'Consider increasing your num_fovs_subset value.' %
            (len(cluster_ids), num_pixel_clusters))

    table = np.stack([totals[c] for c in cluster_ids])
    counts = table[:, -1]
    out = pd.DataFrame(table[:, :-1] / counts[:, None], columns=channels)
    out.insert(0, pixel_cluster_col, np.asarray(cluster_ids).astype(int))
    if keep_count:
        out['count'] = counts.astype(np.int64)
    return out


def find_fovs_missing_col(base_dir, data_dir, missing_col):
    """FOV names in ``data_dir`` that still lack ``missing_col`` -- the restart protocol of
    ``cluster_pixels`` (reference :419-478): work is written to ``data_dir + '_temp'``; when that
    directory exists the FOVs left are those in ``data_dir`` but not yet in it."""
    data_path = os.path.join(base_dir, data_dir)
    temp_path = os.path.join(base_dir, data_dir + '_temp')
    io_utils.validate_paths(data_path)

    if os.path.exists(temp_path):
        done = set(io_utils.list_files(temp_path, substrs='.feather'))
        left = [f for f in io_utils.list_files(data_path, substrs='.feather') if f not in done]
        return io_utils.remove_file_extensions(left)

    fov_files = io_utils.list_files(data_path, substrs='.feather')
    # first readable file decides; corrupted files are skipped over
    sample_cols = None
    for f in fov_files:
        try:
            sample_cols = io_utils.read_table(os.path.join(data_path, f)).column_names
            break
        except (ArrowInvalid, OSError, IOError):
            continue
    if sample_cols is not None and missing_col in sample_cols:
        return []
    os.mkdir(temp_path)
    return io_utils.remove_file_extensions(fov_files)
