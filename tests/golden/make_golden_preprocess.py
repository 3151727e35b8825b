"""Generates tests/golden/preprocess_*.npz -- outputs of the REFERENCE's own arithmetic for the
preprocessing row (N3): scipy.ndimage.gaussian_filter + the pandas filter / normalise calls of
pixie_preprocessing.py:45-78, run through oracle/preprocess_oracle.create_fov_pixel_data with the
scipy / pandas installed in this image (their versions are stored in the file).  These ARE
reference outputs (the two libraries are the reference's dependencies for this path): they pin the
explicit-order restatement and the CUDA kernels.  Run: python tests/golden/make_golden_preprocess.py
"""
import os
import sys

import numpy as np
import pandas as pd
import scipy

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import preprocess_oracle as PO  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def case(name, H, W, C, seed, thresh, sigma=2, empty_from=None):
    r = np.random.default_rng(seed)
    img = r.gamma(0.6, 1.0, (H, W, C)).astype(np.float32)
    if empty_from is not None:
        img[:, empty_from:, :] = 0       # blurred columns >= empty_from + radius are exactly zero
    norm = r.uniform(0.2, 3.0, C)
    if thresh is None:                   # about half of the pixels pass
        from scipy import ndimage
        b = np.stack([ndimage.gaussian_filter((img / norm)[:, :, c], sigma) for c in range(C)], -1)
        thresh = float(np.round(np.median(b.sum(-1)), 3))
    seg = r.integers(0, 50, (H, W)).astype(np.int32)
    channels = ['chan%d' % i for i in range(C)]
    x = img / norm.reshape(1, 1, C)                       # pixie_preprocessing.py:154-161
    mat, _ = PO.create_fov_pixel_data('fov0', channels, x, seg, thresh, blur_factor=sigma)
    np.savez_compressed(
        os.path.join(HERE, name), img=img, norm=norm, seg=seg, thresh=thresh, sigma=sigma,
        blurred=x, X64=mat[channels].values, row_index=mat['row_index'].values,
        column_index=mat['column_index'].values, label=mat['label'].values,
        versions=np.array([scipy.__version__, pd.__version__, np.__version__]))
    print(name, img.shape, "kept", len(mat), "of", H * W)


if __name__ == "__main__":
    case("preprocess_41x37x5.npz", 41, 37, 5, 1, thresh=None)
    case("preprocess_6x50x3_sparse.npz", 6, 50, 3, 2, thresh=0.0, empty_from=20)   # H < blur radius
    case("preprocess_48x40x8.npz", 48, 40, 8, 3, thresh=None)
