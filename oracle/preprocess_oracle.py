"""CPU oracle for the pixel preprocessing row -- TEST INFRASTRUCTURE ONLY.

Two independent statements of /root/reference/src/ark/phenotyping/pixie_preprocessing.py:18-80
(+ normalize_rows, pixel_cluster_utils.py:109-142):

* ``create_fov_pixel_data``: the reference's own route -- ``scipy.ndimage.gaussian_filter`` per
  channel and the pandas filter / normalise / sample calls.  scipy and pandas ARE the reference's
  dependencies for this path and are installed, so this half is the reference's arithmetic itself:
  PINNED.  (alpineer / natsort are absent: the natural channel sort is restated.)
* ``preprocess_explicit``: numpy with every operation spelled out in the order the CUDA kernels
  use (scipy's symmetric correlate1d order, sequential channel sums, IEEE division).
  tests/test_preprocess_oracle.py checks the two agree BIT FOR BIT and against committed golden
  vectors (tests/golden/preprocess_*.npz, made by tests/golden/make_golden.py with scipy 1.18 /
  pandas 3.0)."""
import re

import numpy as np
import pandas as pd
from scipy import ndimage


def _natural_key(name):
    return [int(t) if t.isdigit() else t for t in re.split(r'(\d+)', str(name))]


def create_fov_pixel_data(fov, channels, img_data, seg_labels, pixel_thresh_val,
                          blur_factor=2, subset_proportion=0.1):
    """pixie_preprocessing.py:45-80 with scipy + pandas."""
    channels.sort(key=_natural_key)
    for m in range(len(channels)):
        img_data[:, :, m] = ndimage.gaussian_filter(img_data[:, :, m], sigma=blur_factor)
    mat = pd.DataFrame(img_data.reshape(-1, len(channels)), columns=channels)
    mat['fov'] = fov
    mat['row_index'] = np.repeat(range(img_data.shape[0]), img_data.shape[1])
    mat['column_index'] = np.tile(range(img_data.shape[1]), img_data.shape[0])
    if seg_labels is not None:
        mat['label'] = seg_labels.flatten()
    rowsums = mat[channels].sum(axis=1)
    mat = mat.loc[rowsums > pixel_thresh_val, :].reset_index(drop=True)
    mat = mat.loc[(mat[channels] != 0).any(axis=1), :].reset_index(drop=True)
    # normalize_rows (pixel_cluster_utils.py:109-142)
    sub = mat[channels]
    sub = sub.div(sub.sum(axis=1), axis=0)
    meta = ['fov', 'row_index', 'column_index'] + (['label'] if seg_labels is not None else [])
    sub[meta] = mat.loc[sub.index.values, meta]
    return sub, sub.sample(frac=subset_proportion)


def gaussian_taps(sigma, truncate=4.0):
    radius = int(truncate * float(sigma) + 0.5)
    x = np.arange(-radius, radius + 1)
    phi = np.exp(-0.5 / (float(sigma) ** 2) * x ** 2)
    phi = phi / phi.sum()
    return phi[radius:], radius


def _correlate_axis(a, axis, taps, radius):
    """scipy NI_Correlate1D, symmetric branch, mode 'reflect': centre product first, then the tap
    pairs from the farthest to the nearest, each (left + right) * w added to the running sum."""
    a = np.moveaxis(a, axis, 0)
    n = a.shape[0]
    idx = np.arange(-radius, n + radius)
    if n == 1:
        idx[:] = 0
    else:
        while ((idx < 0) | (idx >= n)).any():
            idx = np.where(idx < 0, -idx - 1, idx)
            idx = np.where(idx >= n, 2 * n - 1 - idx, idx)
    p = a[idx]
    acc = p[radius:radius + n] * taps[0]
    for j in range(radius, 0, -1):
        acc = acc + (p[radius - j:radius - j + n] + p[radius + j:radius + j + n]) * taps[j]
    return np.moveaxis(acc, 0, axis)


def gaussian_blur_explicit(x, sigma):
    """[H, W, C] fp64 -> per-channel 2-D blur, axis 0 then axis 1."""
    taps, radius = gaussian_taps(sigma)
    if radius == 0:
        return x.copy()
    return _correlate_axis(_correlate_axis(x, 0, taps, radius), 1, taps, radius)


def preprocess_explicit(img, norm_vect, pixel_thresh_val, blur_factor, seg_labels=None):
    """img fp32/fp64 [H, W, C] -> dict(blurred, X64, X32, row_index, column_index, label)."""
    x = np.asarray(img).astype(np.float64)
    if norm_vect is not None:
        x = x / np.asarray(norm_vect, np.float64).reshape(1, 1, -1)
    b = gaussian_blur_explicit(x, blur_factor)
    H, W, C = b.shape
    flat = b.reshape(-1, C)
    s = np.zeros(H * W)
    for c in range(C):
        s = s + flat[:, c]
    keep = (s > pixel_thresh_val) & (flat != 0).any(axis=1)
    ii = np.flatnonzero(keep)
    X64 = flat[ii] / s[ii, None]
    return {"blurred": b, "X64": X64, "X32": X64.astype(np.float32),
            "row_index": (ii // W).astype(np.int32), "column_index": (ii % W).astype(np.int32),
            "label": None if seg_labels is None else np.asarray(seg_labels).reshape(-1)[ii]}


def fov_channel_quantiles(pixel_mat, channels, q=0.999):
    """pixie_preprocessing.py:405-410: per-channel quantile of the non-zero entries, pandas' route."""
    return pixel_mat[channels].replace(0, np.nan).quantile(q=q, axis=0)


def column_quantile_explicit(X, q):
    """numpy restatement per column: valid = non-zero, non-NaN; np.quantile(valid, q)."""
    X = np.asarray(X, np.float64)
    out = np.full(X.shape[1], np.nan)
    for c in range(X.shape[1]):
        v = X[:, c]
        v = v[(v != 0) & ~np.isnan(v)]
        if v.size:
            out[c] = np.quantile(v, q)
    return out


def preprocess_fov(base_dir, tiff_dir, data_dir, subset_dir, seg_dir, seg_suffix, img_sub_folder,
                   channels, blur_factor, subset_proportion, pixel_thresh_val, seed,
                   channel_norm_df, fov):
    """pixie_preprocessing.py:83-185 for single-channel TIFF trees (PIL stands in for alpineer's
    loader and skimage's imread): float32 image / float64 normalisation row, seeded subset, the two
    uncompressed Feather files."""
    import os

    import pyarrow.feather as paf
    from PIL import Image
    sub = os.path.join(tiff_dir, fov, img_sub_folder) if img_sub_folder else os.path.join(tiff_dir, fov)
    planes = [np.asarray(Image.open(os.path.join(sub, ch + '.tiff'))) for ch in channels]
    seg = None if seg_dir is None else np.asarray(Image.open(os.path.join(seg_dir, fov + seg_suffix)))
    img = np.stack(planes, axis=-1).astype(np.float32)
    norm = np.array(channel_norm_df.iloc[0].values).reshape([1, 1, -1])
    img = img / norm
    np.random.seed(seed)
    mat, sub_mat = create_fov_pixel_data(fov, channels, img, seg, pixel_thresh_val,
                                         blur_factor=blur_factor,
                                         subset_proportion=subset_proportion)
    paf.write_feather(mat, os.path.join(base_dir, data_dir, fov + ".feather"), compression='uncompressed')
    paf.write_feather(sub_mat, os.path.join(base_dir, subset_dir, fov + ".feather"),
                      compression='uncompressed')
    return mat


def create_pixel_matrix(fovs, channels, base_dir, tiff_dir, seg_dir, img_sub_folder="TIFs",
                        seg_suffix='_whole_cell.tiff', pixel_output_dir='pixel_output_dir',
                        data_dir='pixel_mat_data', subset_dir='pixel_mat_subsetted',
                        norm_vals_name_post_rownorm='channel_norm_post_rownorm.feather',
                        channel_percentile_pre_rownorm=0.99, channel_percentile_post_rownorm=0.999,
                        blur_factor=2, subset_proportion=0.1, seed=42):
    """The arithmetic of pixie_preprocessing.py:188-456 for a fresh cohort (no restart logic),
    with numpy / pandas / scipy only: raw-image channel percentiles (pixel_cluster_utils.py:16-57),
    the pixel threshold (:60-108), preprocess_fov per FOV, the per-FOV 99.9 % quantiles of the
    non-zero entries and their mean.  Returns (pre_norm_df, pixel_thresh_val, post_norm_df)."""
    import os

    import pandas as pd
    import pyarrow.feather as paf
    from PIL import Image

    def load(fov, chans):
        sub = os.path.join(tiff_dir, fov, img_sub_folder) if img_sub_folder else os.path.join(tiff_dir, fov)
        return np.stack([np.asarray(Image.open(os.path.join(sub, c + '.tiff'))) for c in chans], -1)

    channels = sorted(channels)
    means = []
    for ch in channels:
        vals = []
        for fov in fovs:
            img = load(fov, [ch])[:, :, 0]
            img = img[img > 0]
            if len(img) > 0:
                vals.append(np.quantile(img, channel_percentile_pre_rownorm))
        means.append(np.mean(vals))
    pre = pd.DataFrame(np.expand_dims(means, axis=0), columns=channels)
    norm_vect = pre.iloc[0].values.reshape([1, 1, -1])
    thresh = np.mean([np.quantile(np.sum(load(f, channels) / norm_vect, axis=-1), 0.05) for f in fovs])
    for d in (data_dir, subset_dir):
        os.makedirs(os.path.join(base_dir, d), exist_ok=True)
    quant = pd.DataFrame()
    drop = ['fov', 'row_index', 'column_index'] + (['label'] if seg_dir else [])
    for fov in fovs:
        mat = preprocess_fov(base_dir, tiff_dir, data_dir, subset_dir, seg_dir, seg_suffix,
                             img_sub_folder, list(channels), blur_factor, subset_proportion, thresh,
                             seed, pre, fov)
        q = mat.drop(columns=drop).replace(0, np.nan).quantile(
            q=channel_percentile_post_rownorm, axis=0).rename(fov)
        q.index.name = "channel"
        quant = quant.merge(q, how="outer", left_index=True, right_index=True)
    post = pd.DataFrame(quant.mean(axis=1)).sort_index().T
    paf.write_feather(post, os.path.join(base_dir, norm_vals_name_post_rownorm), compression='uncompressed')
    return pre, thresh, post
