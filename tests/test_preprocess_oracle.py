"""CPU tests of the N3 oracle: the explicit-order numpy restatement against the reference's own
scipy + pandas route, bit for bit, on seeded cases and on the committed golden vectors."""
import os

import numpy as np
import pytest
from scipy import ndimage

from oracle import preprocess_oracle as PO

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ["preprocess_41x37x5.npz", "preprocess_6x50x3_sparse.npz", "preprocess_48x40x8.npz"]


@pytest.mark.parametrize("name", CASES)
def test_explicit_restatement_reproduces_the_golden_reference_outputs(name):
    g = np.load(os.path.join(GOLD, name))
    out = PO.preprocess_explicit(g["img"], g["norm"], float(g["thresh"]), float(g["sigma"]), g["seg"])
    np.testing.assert_array_equal(out["blurred"], g["blurred"])
    np.testing.assert_array_equal(out["X64"], g["X64"])
    np.testing.assert_array_equal(out["row_index"], g["row_index"])
    np.testing.assert_array_equal(out["column_index"], g["column_index"])
    np.testing.assert_array_equal(out["label"], g["label"])
    assert 0 < len(out["X64"]) < g["img"].shape[0] * g["img"].shape[1]   # the filter did filter


@pytest.mark.parametrize("shape,sigma", [((33, 29, 4), 2), ((5, 70, 2), 2), ((70, 3, 2), 2),
                                         ((40, 40, 3), 1), ((20, 20, 2), 3.5), ((1, 9, 1), 2)])
def test_blur_order_matches_scipy_bit_for_bit(shape, sigma, rng):
    x = rng.random(shape) * 10
    ref = np.stack([ndimage.gaussian_filter(x[:, :, c], sigma=sigma) for c in range(shape[2])], -1)
    np.testing.assert_array_equal(PO.gaussian_blur_explicit(x, sigma), ref)


def test_pandas_route_against_explicit_on_a_fresh_case(rng):
    H, W, C = 48, 52, 7
    img = rng.gamma(0.5, 1.0, (H, W, C)).astype(np.float32)
    norm = rng.uniform(0.5, 2.0, C)
    seg = rng.integers(0, 9, (H, W))
    channels = ['c%d' % i for i in range(C)]
    x = img / norm.reshape(1, 1, C)
    np.random.seed(42)
    mat, sub = PO.create_fov_pixel_data('f', channels, x, seg, 2.0)
    out = PO.preprocess_explicit(img, norm, 2.0, 2, seg)
    np.testing.assert_array_equal(mat[channels].values, out["X64"])
    np.testing.assert_array_equal(mat['row_index'].values, out["row_index"])
    np.testing.assert_array_equal(mat['label'].values, out["label"])
    assert list(mat.columns) == channels + ['fov', 'row_index', 'column_index', 'label']
    assert len(sub) == round(0.1 * len(mat))
    # rows sum to 1 (pixie_preprocessing_test.py:79 of the reference asserts the same)
    np.testing.assert_allclose(mat[channels].sum(axis=1).values, 1.0, rtol=0, atol=1e-12)


def test_channel_names_sort_naturally():
    names = ['chan10', 'chan2', 'chan1', 'CD45', 'CD4']
    names.sort(key=PO._natural_key)
    assert names == ['CD4', 'CD45', 'chan1', 'chan2', 'chan10']


def test_quantile_routes_agree_and_host_interpolation_is_numpys(rng):
    """pandas' replace(0, nan).quantile against per-column np.quantile of the valid entries, and
    the host-side interpolation the GPU path uses (order statistics + numpy's lerp formula) against
    both -- bit for bit."""
    import pandas as pd
    from ark_analysis_b200.pixie_preprocessing import _lerp
    X = rng.random((500, 6)) * np.array([1, 1e-3, 50, 1, 1, 1])
    X[rng.random(X.shape) < 0.3] = 0
    X[:, 4] = np.round(X[:, 4], 1)       # heavy ties
    X[:, 5] = 0                          # no valid entry
    chans = ['c%d' % i for i in range(6)]
    df = pd.DataFrame(X, columns=chans)
    for q in (0.999, 0.5, 0.05, 0.0, 1.0, 0.3141):
        ref = PO.fov_channel_quantiles(df, chans, q).values
        np.testing.assert_array_equal(PO.column_quantile_explicit(X, q), ref)
        mine = np.full(6, np.nan)
        for c in range(6):
            v = np.sort(X[:, c][X[:, c] != 0])
            if v.size:
                vi = (v.size - 1) * np.float64(q)
                r = min(int(np.floor(vi)), v.size - 1)
                mine[c] = _lerp(v[r:r + 1], v[min(r + 1, v.size - 1):][:1], np.array([vi - np.floor(vi)]))[0]
        np.testing.assert_array_equal(mine, ref)


def _explicit_as_device_result(img, norm_vect=None, pixel_thresh_val=0.0, blur_factor=2,
                               seg_labels=None, **_):
    """TEST-ONLY stand-in for preprocess_fov_device: the explicit-order numpy restatement wrapped in
    the result format of the device call (CPU tensors), so the HOST logic of the mirror -- file
    layout, column order, seeding, Feather writes -- can be checked without a GPU.  The product
    has no such route (test_host_logic.py checks it raises without a device)."""
    import torch
    r = PO.preprocess_explicit(img, norm_vect, pixel_thresh_val, blur_factor, seg_labels)
    out = {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in r.items() if v is not None}
    out["label"] = None if seg_labels is None else torch.from_numpy(np.ascontiguousarray(r["label"]))
    out["n"] = len(r["X64"])
    return out


@pytest.mark.parametrize("with_seg,sub_dir", [(True, 'TIFs'), (False, None)])
def test_preprocess_fov_host_logic_against_the_pandas_route(with_seg, sub_dir, rng, monkeypatch, tmp_path):
    import pandas as pd
    import preprocess_fixtures as PF
    from ark_analysis_b200 import pixie_preprocessing as PP
    monkeypatch.setattr(PP, "preprocess_fov_device", _explicit_as_device_result)
    out, chans = PF.run_both(str(tmp_path), rng, PP.preprocess_fov, PO.preprocess_fov,
                             with_seg=with_seg, sub_dir=sub_dir)
    (m_ret, m_full, m_sub), (o_ret, o_full, o_sub) = out['mirror'], out['oracle']
    pd.testing.assert_frame_equal(m_full, o_full)
    pd.testing.assert_frame_equal(m_sub, o_sub)
    pd.testing.assert_frame_equal(m_ret, o_ret)
    assert list(m_full.columns[:3]) == chans and ('label' in m_full.columns) == with_seg
    assert 0 < len(m_full) < 40 * 36 and len(m_sub) == round(0.1 * len(m_full))
    assert np.all(m_full[chans].sum(axis=1) != 0)       # the reference test's own assertion
    # the reference's error behaviour for this function
    with pytest.raises(ValueError):
        PP.load_fov_channels(str(tmp_path / 'sample_image_data'), 'fov0', ['chan0', 'nope'], sub_dir)
    with pytest.raises(FileNotFoundError):
        PP.load_fov_channels(str(tmp_path / 'sample_image_data'), 'fov9', ['chan0'], sub_dir)
    with pytest.raises(NotImplementedError):
        PP.preprocess_fov(str(tmp_path), str(tmp_path), 'a', 'b', None, '', None, True, chans, 2,
                          0.1, 1.0, 42, None, 'fov0')
