"""On-disk FOV trees for the preprocess_fov tests (the layout of the reference's
tests/phenotyping/pixie_preprocessing_test.py:136-170: tiff_dir/<fov>/TIFs/<chan>.tiff + a
segmentation directory) and the run-both-and-compare helper shared by the CPU and GPU tests."""
import os

import numpy as np
import pandas as pd
import pyarrow.feather as paf
from PIL import Image


def make_tree(temp_dir, rng, fovs=('fov0', 'fov1'), chans=('chan0', 'chan1', 'chan2'),
              shape=(40, 36), sub_dir='TIFs', dtype=np.float32):
    tiff_dir = os.path.join(temp_dir, 'sample_image_data')
    seg_dir = os.path.join(temp_dir, 'segmentation')
    os.mkdir(tiff_dir)
    os.mkdir(seg_dir)
    for fov in fovs:
        d = os.path.join(tiff_dir, fov, sub_dir) if sub_dir else os.path.join(tiff_dir, fov)
        os.makedirs(d)
        for ch in chans:
            plane = rng.gamma(0.8, 20.0, shape).astype(dtype)
            Image.fromarray(plane).save(os.path.join(d, ch + '.tiff'))
        Image.fromarray(rng.integers(0, 16, shape).astype(np.int32)).save(
            os.path.join(seg_dir, fov + '_whole_cell.tiff'))
    norm = pd.DataFrame(np.expand_dims(rng.uniform(5, 15, len(chans)), 0), columns=list(chans))
    return tiff_dir, seg_dir, norm


def run_both(temp_dir, rng, mirror, oracle_fn, with_seg=True, sub_dir='TIFs'):
    """Runs the mirror and the oracle on the same tree (separate output directories) and returns
    the four tables read back from the Feather files plus the two return values."""
    chans = ['chan0', 'chan1', 'chan2']
    tiff_dir, seg_dir, norm = make_tree(temp_dir, rng, sub_dir=sub_dir)
    out = {}
    for tag, fn in (('mirror', mirror), ('oracle', oracle_fn)):
        base = os.path.join(temp_dir, tag)
        os.makedirs(os.path.join(base, 'pixel_mat_data'))
        os.makedirs(os.path.join(base, 'pixel_mat_subsetted'))
        args = [base, tiff_dir, 'pixel_mat_data', 'pixel_mat_subsetted',
                seg_dir if with_seg else None, '_whole_cell.tiff', sub_dir]
        tail = [list(chans), 2, 0.1, 4.8, 42, norm, "fov0"]
        ret = fn(*args, False, *tail) if tag == 'mirror' else fn(*args, *tail)
        out[tag] = (ret,
                    paf.read_feather(os.path.join(base, 'pixel_mat_data', 'fov0.feather')),
                    paf.read_feather(os.path.join(base, 'pixel_mat_subsetted', 'fov0.feather')))
    return out, chans
