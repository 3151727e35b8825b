"""Pixie pixel preprocessing with the arithmetic on the GPU.

Mirrors ``create_fov_pixel_data`` of the reference (src/ark/phenotyping/pixie_preprocessing.py:18-80)
-- same arguments, same two DataFrames -- and adds the device-native form the SOM path wants:
``preprocess_fov_device`` leaves the normalised pixel matrix in HBM (fp32 rows for
``som.bmu`` / ``som.train_som``, fp64 rows for the Feather file), so a FOV goes image -> labels without
a DataFrame in between.  Blur, filter and row normalisation run in ``pixie_preprocess_fov_f64``
(csrc/preprocess_kernels.cu) in fp64 and in the reference's operation order: bit-identical values.
"""
import ctypes
import re

import numpy as np
import torch

from . import _native, io_utils
from ._native import PixieError
from . import som as _som

BLUR_ONLY, IMG_F64 = 1, 2
MAX_RADIUS = 32


def gaussian_taps(sigma, truncate=4.0):
    """Half of scipy.ndimage's 1-D gaussian kernel, centre first (``_gaussian_kernel1d`` with
    order 0: exp(-0.5 / sigma^2 * x^2) normalised by the sum over [-radius, radius])."""
    sd = float(sigma)
    radius = int(truncate * sd + 0.5)
    if sd <= 1e-15 or radius == 0:
        return np.ones(1, np.float64), 0   # scipy skips axes with sigma <= 1e-15
    if radius > MAX_RADIUS:
        raise PixieError(f"blur radius {radius} exceeds the kernel's limit of {MAX_RADIUS} taps")
    x = np.arange(-radius, radius + 1)
    phi = np.exp(-0.5 / (sd * sd) * x ** 2)
    phi = phi / phi.sum()
    return np.ascontiguousarray(phi[radius:]), radius


def natural_key(name):
    """Sort key of ``natsort.natsort_key`` for plain channel names: digit runs compare as numbers."""
    return [int(t) if t.isdigit() else t for t in re.split(r'(\d+)', str(name))]


def preprocess_fov_device(img, norm_vect=None, pixel_thresh_val=0.0, blur_factor=2,
                          seg_labels=None, want_x64=True, want_x32=True, blur_only=False,
                          device=None):
    """Run the preprocessing of one FOV on the GPU.

    ``img``: [H, W, C] image stack, float32 (divided by ``norm_vect`` in fp64, as preprocess_fov
    does) or float64 (already normalised, what create_fov_pixel_data receives); numpy or CUDA.
    Returns a dict of CUDA tensors: ``blurred`` [H, W, C] fp64, and unless ``blur_only``:
    ``X64`` [n, C] fp64, ``X32`` [n, C] fp32 (row pitch a multiple of 4 floats: feeds som.bmu
    directly), ``row_index`` / ``column_index`` / ``label`` int32 [n], ``n`` (int)."""
    dev = torch.device(device) if device is not None else _som._default_device()
    t = torch.as_tensor(img)
    if t.dim() != 3:
        raise PixieError("img must be [H, W, C]")
    if t.dtype not in (torch.float32, torch.float64):
        t = t.to(torch.float32)
    t = t.to(dev).contiguous()
    H, W, C = (int(v) for v in t.shape)
    flags = (IMG_F64 if t.dtype == torch.float64 else 0) | (BLUR_ONLY if blur_only else 0)
    taps, radius = gaussian_taps(blur_factor)
    norm = None
    if norm_vect is not None:
        norm = torch.as_tensor(np.asarray(norm_vect, np.float64).reshape(-1)).to(dev)
        if norm.numel() != C:
            raise PixieError("norm_vect must hold one value per channel")
    seg = None
    if seg_labels is not None and not blur_only:
        seg = torch.as_tensor(np.ascontiguousarray(seg_labels)).reshape(-1).to(dev).to(torch.int32)
        if seg.numel() != H * W:
            raise PixieError("seg_labels must have H * W elements")
    n = H * W
    lib = _native.lib()
    ws = torch.empty(lib.pixie_preprocess_workspace_bytes(H, W, C), dtype=torch.uint8, device=dev)
    out = {"blurred": torch.empty((H, W, C), dtype=torch.float64, device=dev)}
    X64 = X32 = rows = cols = labs = nk = None
    ld = (C + 3) // 4 * 4
    if not blur_only:
        X64 = torch.empty((n, C), dtype=torch.float64, device=dev) if want_x64 else None
        X32 = (torch.zeros if ld != C else torch.empty)((n, ld), dtype=torch.float32, device=dev) \
            if want_x32 else None
        rows = torch.empty(n, dtype=torch.int32, device=dev)
        cols = torch.empty(n, dtype=torch.int32, device=dev)
        labs = torch.empty(n, dtype=torch.int32, device=dev) if seg is not None else None
        nk = torch.zeros(1, dtype=torch.int64, device=dev)
    p = _som._ptr
    with torch.cuda.device(dev):
        rc = lib.pixie_preprocess_fov_f64(
            p(t), H, W, C, p(norm), taps.ctypes.data_as(ctypes.c_void_p), radius,
            float(pixel_thresh_val), p(seg), p(out["blurred"]), p(X64), p(X32), ld, p(rows),
            p(cols), p(labs), p(nk), p(ws), ws.numel(), flags, _som._stream(dev))
    _native.check(rc, "pixie_preprocess_fov_f64")
    if blur_only:
        return out
    k = int(nk.item())   # also orders the host's view after the kernels
    out.update(n=k, X64=None if X64 is None else X64[:k],
               X32=None if X32 is None else X32[:k, :C],
               row_index=rows[:k], column_index=cols[:k],
               label=None if labs is None else labs[:k])
    return out


def _pixel_tables(res, fov, channels, seg_labels, subset_proportion):
    """The two DataFrames of create_fov_pixel_data from the device results."""
    import pandas as pd
    pixel_mat = pd.DataFrame(res["X64"].cpu().numpy(), columns=channels)
    pixel_mat['fov'] = fov
    pixel_mat['row_index'] = res["row_index"].cpu().numpy().astype(np.int64)
    pixel_mat['column_index'] = res["column_index"].cpu().numpy().astype(np.int64)
    if seg_labels is not None:
        pixel_mat['label'] = res["label"].cpu().numpy().astype(np.asarray(seg_labels).dtype)
    pixel_mat_subset = pixel_mat.sample(frac=subset_proportion)
    return pixel_mat, pixel_mat_subset


def create_fov_pixel_data(fov, channels, img_data, seg_labels, pixel_thresh_val,
                          blur_factor=2, subset_proportion=0.1):
    """The reference's ``create_fov_pixel_data``: (pixel_mat, pixel_mat_subset) for one FOV --
    blurred, thresholded, row-normalised channel columns plus ``fov``, ``row_index``,
    ``column_index`` (and ``label`` when ``seg_labels`` is given); the subset is
    ``pixel_mat.sample(frac=subset_proportion)`` under numpy's global seed, as there.  Like the
    reference, ``channels`` is sorted in place and ``img_data`` receives the blurred planes."""
    channels.sort(key=natural_key)
    res = preprocess_fov_device(img_data, None, pixel_thresh_val, blur_factor, seg_labels,
                                want_x64=True, want_x32=False)
    if isinstance(img_data, np.ndarray) and img_data.dtype == np.float64 and img_data.flags.writeable:
        img_data[...] = res["blurred"].cpu().numpy()
    return _pixel_tables(res, fov, channels, seg_labels, subset_proportion)


def _read_image(path):
    from PIL import Image
    with Image.open(path) as im:
        return np.asarray(im)


def load_fov_channels(tiff_dir, fov, channels, img_sub_folder=None):
    """[H, W, C] stack of ``tiff_dir/fov/[img_sub_folder/]<channel>.tiff`` in the order of
    ``channels`` -- the slice ``img_xr.loc[fov, :, :, channels]`` the reference takes from
    alpineer's ``load_imgs_from_tree`` (pixie_preprocessing.py:131-154).  A channel without a file
    is a ValueError, as the reference's verify_in_list makes it."""
    import os
    base = os.path.join(tiff_dir, fov, img_sub_folder) if img_sub_folder else os.path.join(tiff_dir, fov)
    io_utils.validate_paths(base)
    planes, missing = [], []
    for ch in channels:
        for ext in ('.tiff', '.tif'):
            path = os.path.join(base, ch + ext)
            if os.path.exists(path):
                planes.append(_read_image(path))
                break
        else:
            missing.append(ch)
    if missing:
        raise ValueError("Not all values given in list provided_chans were found in list "
                         "pixel_mat_chans.\n Invalid values (first 10): " + ", ".join(missing[:10]))
    return np.stack(planes, axis=-1)


def preprocess_fov(base_dir, tiff_dir, data_dir, subset_dir, seg_dir, seg_suffix,
                   img_sub_folder, is_mibitiff, channels, blur_factor,
                   subset_proportion, pixel_thresh_val, seed, channel_norm_df, fov,
                   _device_result=None):
    """The reference's ``preprocess_fov`` (pixie_preprocessing.py:83-185): load one FOV's channel
    images (and its segmentation mask), normalise by ``channel_norm_df``, blur / filter / row-
    normalise, write the full table to ``base_dir/data_dir/<fov>.feather`` and the sampled subset to
    ``base_dir/subset_dir/<fov>.feather``, return the full table.  The float32 image and the
    normalisation row go to the device as they are (the kernel divides in fp64, as numpy does for
    float32 / float64); single-channel TIFFs only (MIBItiff containers are image IO, out of scope)."""
    import os
    if is_mibitiff:
        raise NotImplementedError("MIBItiff containers are not read here: extract the channels to "
                                  "single TIFFs (image IO is outside the Pixie SOM path)")
    img = load_fov_channels(tiff_dir, fov, channels, img_sub_folder)
    seg_labels = _read_image(os.path.join(seg_dir, fov + seg_suffix)) if seg_dir is not None else None
    img_data = img.astype(np.float32)
    norm_vect = np.array(channel_norm_df.iloc[0].values)
    np.random.seed(seed)                       # the subset is drawn from numpy's global stream
    channels.sort(key=natural_key)             # as create_fov_pixel_data does (names only)
    res = preprocess_fov_device(img_data, norm_vect, pixel_thresh_val, blur_factor, seg_labels,
                                want_x64=True, want_x32=False)
    pixel_mat, pixel_mat_subset = _pixel_tables(res, fov, channels, seg_labels, subset_proportion)
    io_utils.write_dataframe(pixel_mat, os.path.join(base_dir, data_dir, fov + ".feather"),
                             compression='uncompressed')
    io_utils.write_dataframe(pixel_mat_subset, os.path.join(base_dir, subset_dir, fov + ".feather"),
                             compression='uncompressed')
    if _device_result is not None:  # create_pixel_matrix takes the quantiles from the device rows
        _device_result.update(res)
    return pixel_mat


# ------------------------------------------------------------------------------------------------
# the cohort driver (reference pixie_preprocessing.py:188-456) and its raw-image statistics
# (pixel_cluster_utils.py:16-108)
# ------------------------------------------------------------------------------------------------
def calculate_channel_percentiles(tiff_dir, fovs, channels, img_sub_folder, percentile):
    """1 x C DataFrame: per channel, the mean over FOVs of ``np.quantile(img[img > 0], percentile)``
    (FOVs without a non-zero pixel are left out), columns in natural order -- the reference's
    ``calculate_channel_percentiles``.  Host numpy on the raw images: image IO bound, one scalar per
    (FOV, channel)."""
    import pandas as pd
    means = []
    for channel in channels:
        per_fov = []
        for fov in fovs:
            img = load_fov_channels(tiff_dir, fov, [channel], img_sub_folder)[:, :, 0]
            nz = img[img > 0]
            if len(nz) > 0:
                per_fov.append(np.quantile(nz, percentile))
        means.append(np.mean(per_fov))
    order = sorted(range(len(channels)), key=lambda i: natural_key(channels[i]))
    return pd.DataFrame([[means[i] for i in order]], columns=[channels[i] for i in order])


def calculate_pixel_intensity_percentile(tiff_dir, fovs, channels, img_sub_folder,
                                         channel_percentiles, percentile=0.05):
    """Mean over FOVs of the ``percentile`` quantile of the per-pixel total signal after every
    channel was divided by its ``channel_percentiles`` value (the reference's function of the
    same name): the pixel threshold of the preprocessing."""
    norm_vect = channel_percentiles.iloc[0].values.reshape([1, 1, -1])
    per_fov = []
    for fov in fovs:
        img = load_fov_channels(tiff_dir, fov, channels, img_sub_folder)
        per_fov.append(np.quantile(np.sum(img / norm_vect, axis=-1), percentile))
    return np.mean(per_fov)


def check_for_modified_channels(tiff_dir, test_fov, img_sub_folder, channels):
    """Warn when a selected channel also exists in a modified version (``_smoothed``,
    ``_nuc_include``, ``_nuc_exclude``) in the example FOV's folder."""
    import os
    import warnings
    present = set(io_utils.remove_file_extensions(
        io_utils.list_files(os.path.join(tiff_dir, test_fov, img_sub_folder or ''))))
    for channel in channels:
        for mod in ('_smoothed', '_nuc_include', '_nuc_exclude'):
            if channel + mod in present:
                warnings.warn('You selected {} as the channel to analyze, but there were potential'
                              ' modified channels found: {}. Make sure you selected the correct '
                              'version of the channel for inclusion in '
                              'clustering'.format(channel, channel + mod))


class _Cohort:
    """Files and restart state of one preprocessing run under ``base_dir``."""

    def __init__(self, base_dir, pixel_output_dir, data_dir, subset_dir, pre_name, thresh_name):
        import os
        self.data = os.path.join(base_dir, data_dir)
        self.subset = os.path.join(base_dir, subset_dir)
        self.pre_norm = os.path.join(base_dir, pixel_output_dir, pre_name)
        self.thresh = os.path.join(base_dir, pixel_output_dir, thresh_name)
        self.quantiles = os.path.join(self.data, "channel_norm_post_rownorm_perfov.csv")
        for d in (self.data, self.subset):
            if not os.path.exists(d):
                os.mkdir(d)

    def reset_if_channels_changed(self, channels):
        """A different channel set invalidates everything written so far."""
        import os
        from shutil import rmtree
        if not os.path.exists(self.pre_norm):
            return
        if set(io_utils.read_dataframe(self.pre_norm).columns.values) == set(channels):
            return
        print("New channels provided: overwriting whole cohort")
        for d in (self.data, self.subset):
            rmtree(d)
            os.mkdir(d)
        os.remove(self.pre_norm)
        os.remove(self.thresh)

    def finished_fovs(self):
        """FOVs with BOTH their full and their subset file (a lone file is regenerated)."""
        both = set(io_utils.list_files(self.subset, substrs='.feather')) & \
            set(io_utils.list_files(self.data, substrs='.feather'))
        return set(io_utils.remove_file_extensions(sorted(both)))


def create_pixel_matrix(fovs, channels, base_dir, tiff_dir, seg_dir,
                        img_sub_folder="TIFs", seg_suffix='_whole_cell.tiff',
                        pixel_output_dir='pixel_output_dir',
                        data_dir='pixel_mat_data',
                        subset_dir='pixel_mat_subsetted',
                        norm_vals_name_pre_rownorm='channel_norm_pre_rownorm.feather',
                        norm_vals_name_post_rownorm='channel_norm_post_rownorm.feather',
                        pixel_thresh_name='pixel_thresh.feather',
                        channel_percentile_pre_rownorm=0.99, channel_percentile_post_rownorm=0.999,
                        is_mibitiff=False, blur_factor=2, subset_proportion=0.1, seed=42,
                        multiprocess=False, batch_size=5):
    """The notebook-facing driver of the preprocessing (reference pixie_preprocessing.py:188-456):
    per FOV, blur / threshold / row-normalise the channel images on the device
    (``preprocess_fov``), write the full and the subsetted pixel tables, and collect the per-FOV
    99.9 % quantiles whose mean becomes ``base_dir/norm_vals_name_post_rownorm`` -- the row
    ``PixelSOMCluster`` divides by.  Same arguments, files, restart behaviour and messages as the
    reference.  ``multiprocess`` / ``batch_size`` are accepted for compatibility: the FOVs go
    through the one GPU in turn (a FOV takes ~1 ms of kernels; the loop is image-IO bound)."""
    import os
    import pandas as pd

    channels.sort(key=natural_key)
    if subset_proportion <= 0 or subset_proportion > 1:
        raise ValueError('Invalid subset percentage entered: must be in (0, 1]')
    io_utils.validate_paths([base_dir, tiff_dir, os.path.join(base_dir, pixel_output_dir)])

    cohort = _Cohort(base_dir, pixel_output_dir, data_dir, subset_dir,
                     norm_vals_name_pre_rownorm, pixel_thresh_name)
    cohort.reset_if_channels_changed(channels)

    todo = set(fovs) - cohort.finished_fovs()
    if not todo:
        print("There are no more FOVs to preprocess, skipping")
        return
    # FOVs whose quantiles never reached the per-FOV file are redone as well
    quant_all = pd.read_csv(cohort.quantiles, index_col="channel") \
        if os.path.exists(cohort.quantiles) else pd.DataFrame()
    todo |= set(fovs) - set(quant_all.columns)
    todo = sorted(todo, key=natural_key)
    if len(todo) < len(fovs):
        print("Restarting preprocessing from FOV %s, "
              "%d fovs left to process" % (todo[0], len(todo)))

    check_for_modified_channels(tiff_dir=tiff_dir, test_fov=fovs[0],
                                img_sub_folder=img_sub_folder, channels=channels)

    # cohort-wide statistics of the raw images: computed once, then read back on a restart
    if os.path.exists(cohort.pre_norm):
        pre_norm = io_utils.read_dataframe(cohort.pre_norm)
    else:
        pre_norm = calculate_channel_percentiles(tiff_dir, fovs, channels, img_sub_folder,
                                                 channel_percentile_pre_rownorm)
        io_utils.write_dataframe(pre_norm, cohort.pre_norm, compression='uncompressed')
    if os.path.exists(cohort.thresh):
        pixel_thresh_val = io_utils.read_dataframe(cohort.thresh)['pixel_thresh_val'].values[0]
    else:
        pixel_thresh_val = calculate_pixel_intensity_percentile(
            tiff_dir, fovs, channels, img_sub_folder, pre_norm)
        io_utils.write_dataframe(pd.DataFrame({'pixel_thresh_val': [pixel_thresh_val]}),
                                 cohort.thresh, compression='uncompressed')

    for done, fov in enumerate(todo, start=1):
        on_device = {}
        preprocess_fov(base_dir, tiff_dir, data_dir, subset_dir, seg_dir, seg_suffix,
                       img_sub_folder, is_mibitiff, channels, blur_factor, subset_proportion,
                       pixel_thresh_val, seed, pre_norm, fov, _device_result=on_device)
        # the FOV's quantiles of the non-zero entries, from the rows still on the device
        quant_fov = fov_channel_quantiles(on_device["X64"], channels,
                                          channel_percentile_post_rownorm, name=fov)
        quant_fov.index.name = "channel"
        quant_all = quant_all.merge(quant_fov, how="outer", left_index=True, right_index=True)
        quant_all.to_csv(cohort.quantiles)  # after every FOV: the restart point
        if multiprocess:
            if done % batch_size == 0 or done == len(todo):
                print("Processed %d fovs" % done)
        elif done % 10 == 0 or done == len(todo):
            print("Processed %d fovs" % done)

    mean_quant = pd.DataFrame(quant_all.mean(axis=1))
    mean_quant = mean_quant.loc[sorted(mean_quant.index, key=natural_key)]
    io_utils.write_dataframe(mean_quant.T, os.path.join(base_dir, norm_vals_name_post_rownorm),
                             compression='uncompressed')
    os.remove(cohort.quantiles)


# ------------------------------------------------------------------------------------------------
# per-channel quantiles of the non-zero entries (reference pixie_preprocessing.py:405-410, :424-427)
# ------------------------------------------------------------------------------------------------
def _lerp(a, b, t):
    """numpy's ``_lerp`` (the interpolation step of np.quantile / np.percentile, method 'linear')."""
    diff = b - a
    out = a + diff * t
    hi = t >= 0.5
    out[hi] = (b - diff * (1 - t))[hi]
    return out


def column_order_stats(X64, q):
    """(lo, hi, m) per column of a CUDA float64 matrix [n, C]: the number m of valid entries
    (non-zero, non-NaN) and the valid entries of rank floor((m-1) q) and that rank + 1 -- exact,
    by radix select on the device (``pixie_column_quantile_f64``).  numpy arrays of length C."""
    if not isinstance(X64, torch.Tensor) or not X64.is_cuda or X64.dtype != torch.float64 \
            or X64.dim() != 2 or (X64.shape[1] > 1 and X64.stride(1) != 1):
        raise PixieError("X64 must be a CUDA float64 matrix [n, C] with contiguous rows")
    n, C = (int(v) for v in X64.shape)
    dev = X64.device
    lib = _native.lib()
    lo = torch.empty(C, dtype=torch.float64, device=dev)
    hi = torch.empty(C, dtype=torch.float64, device=dev)
    m = torch.empty(C, dtype=torch.int64, device=dev)
    ws = torch.empty(lib.pixie_column_quantile_workspace_bytes(C), dtype=torch.uint8, device=dev)
    p = _som._ptr
    with torch.cuda.device(dev):
        rc = lib.pixie_column_quantile_f64(p(X64), n, C, X64.stride(0) if n > 1 else max(C, 1),
                                           float(q), p(lo), p(hi), p(m), p(ws), ws.numel(),
                                           _som._stream(dev))
    _native.check(rc, "pixie_column_quantile_f64")
    return lo.cpu().numpy(), hi.cpu().numpy(), m.cpu().numpy()


def column_quantile(X64, q):
    """``np.quantile(column[valid], q)`` (method 'linear') for every column, valid = non-zero and
    non-NaN; NaN for a column without valid entries.  Bit-identical to numpy: the two order
    statistics come from the device, the interpolation is numpy's own formula."""
    q = np.float64(q)
    lo, hi, m = column_order_stats(X64, q)
    v = (m - 1).astype(np.float64) * q            # virtual index (n - 1) * q
    gamma = v - np.floor(v)
    with np.errstate(invalid="ignore"):
        out = _lerp(lo, hi, gamma)
    out[m == 0] = np.nan
    return out


def fov_channel_quantiles(X64, channels, q=0.999, name=None):
    """The reference's per-FOV normalisation statistic
    ``pixel_mat[channels].replace(0, np.nan).quantile(q=q, axis=0).rename(fov)``
    (pixie_preprocessing.py:405-410) on the device-resident pixel matrix: a float64 Series indexed
    by channel.  pandas < 3 hands ``q * 100`` to np.percentile, which divides by 100 again (that
    can move q by one ulp, and the quantile by a few); pandas >= 3 calls np.quantile with q itself.
    The installed pandas decides which of the two is reproduced."""
    import pandas as pd
    import pandas.core.array_algos.quantile as _pq
    q_eff = np.float64(q)
    if not hasattr(_pq, "_nanquantile"):      # the np.percentile route of pandas 1.x / 2.x
        q_eff = np.true_divide(q_eff * 100.0, np.float64(100))
    out = pd.Series(column_quantile(X64, q_eff), index=pd.Index(list(channels)), name=q if name is None else name)
    return out
