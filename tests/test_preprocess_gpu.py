"""GPU parity tests of the N3 row: pixie_preprocess_fov_f64 through the C ABI
and the create_fov_pixel_data mirror, against the reference's scipy + pandas route (golden vectors
and live, oracle/preprocess_oracle.py).  fp64 in the reference's operation order: BIT-EXACT."""
import os

import numpy as np
import pandas as pd
import pytest
import torch
import pyarrow.feather as paf
from scipy import ndimage

from oracle import preprocess_oracle as PO
from ark_analysis_b200 import pixie_preprocessing as PP, som as S

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ["preprocess_41x37x5.npz", "preprocess_6x50x3_sparse.npz", "preprocess_48x40x8.npz"]


@pytest.mark.parametrize("name", CASES)
def test_golden_reference_outputs(name):
    g = np.load(os.path.join(GOLD, name))
    out = PP.preprocess_fov_device(g["img"], g["norm"], float(g["thresh"]), float(g["sigma"]), g["seg"])
    np.testing.assert_array_equal(out["blurred"].cpu().numpy(), g["blurred"])
    assert out["n"] == len(g["X64"])
    np.testing.assert_array_equal(out["X64"].cpu().numpy(), g["X64"])
    np.testing.assert_array_equal(out["X32"].cpu().numpy(), g["X64"].astype(np.float32))
    np.testing.assert_array_equal(out["row_index"].cpu().numpy(), g["row_index"])
    np.testing.assert_array_equal(out["column_index"].cpu().numpy(), g["column_index"])
    np.testing.assert_array_equal(out["label"].cpu().numpy(), g["label"])


@pytest.mark.parametrize("shape,sigma", [((33, 29, 4), 2), ((5, 70, 2), 2), ((70, 3, 2), 2),
                                         ((40, 40, 3), 1), ((20, 20, 2), 3.5), ((1, 9, 1), 2),
                                         ((16, 16, 1), 0)])
def test_blur_matches_scipy_bit_for_bit(shape, sigma, rng):
    x = rng.random(shape) * 10
    ref = np.stack([ndimage.gaussian_filter(x[:, :, c], sigma=sigma) for c in range(shape[2])], -1)
    out = PP.preprocess_fov_device(x, blur_factor=sigma, blur_only=True)      # fp64 input path
    np.testing.assert_array_equal(out["blurred"].cpu().numpy(), ref)


def test_everything_filtered_and_nothing_filtered(rng):
    img = rng.random((20, 24, 3)).astype(np.float32)
    none = PP.preprocess_fov_device(img, None, 1e9, 2)
    assert none["n"] == 0 and none["X64"].shape == (0, 3)
    every = PP.preprocess_fov_device(img, None, -1.0, 2)
    assert every["n"] == 20 * 24
    zeros = PP.preprocess_fov_device(np.zeros((20, 24, 3), np.float32), None, -1.0, 2)
    assert zeros["n"] == 0      # rows whose channels are all zero go even when the sum passes


def test_create_fov_pixel_data_mirror_equals_the_pandas_route(rng):
    H, W, C = 48, 52, 7
    img = rng.gamma(0.5, 1.0, (H, W, C)).astype(np.float32)
    norm = rng.uniform(0.5, 2.0, C)
    seg = rng.integers(0, 9, (H, W))
    for seg_arg in (seg, None):
        ch_a = ['chan10', 'chan2', 'chan1', 'CD45', 'CD4', 'HH3', 'Ki67']
        ch_b = list(ch_a)
        xa = img / norm.reshape(1, 1, C)
        xb = xa.copy()
        np.random.seed(7)
        ref_mat, ref_sub = PO.create_fov_pixel_data('fov3', ch_a, xa, seg_arg, 2.8)
        np.random.seed(7)
        mat, sub = PP.create_fov_pixel_data('fov3', ch_b, xb, seg_arg, 2.8)
        assert ch_a == ch_b                           # sorted in place, naturally
        np.testing.assert_array_equal(xa, xb)         # img_data receives the blurred planes
        pd.testing.assert_frame_equal(mat, ref_mat)
        pd.testing.assert_frame_equal(sub, ref_sub)   # same rows sampled under the same seed
        assert 0 < len(mat) < H * W


def test_full_fov_size_properties_and_device_handoff():
    """cfg2-sized FOV (1024 x 1024 x 32): two channels of the blur against scipy bit for bit, rows
    sum to one, indices in image order, and the fp32 matrix feeds the BMU kernel as it is."""
    H = W = 1024
    C = 32
    g = torch.Generator(device="cuda").manual_seed(3)
    img = torch.empty((H, W, C), device="cuda").exponential_(1.0, generator=g)
    img *= (torch.rand((H, W, 1), device="cuda", generator=g) > 0.3)
    norm = np.linspace(0.5, 2.0, C)
    out = PP.preprocess_fov_device(img, norm, 20.0, 2)
    host = img.cpu().numpy()
    for c in (0, 17):
        ref = ndimage.gaussian_filter(host[:, :, c].astype(np.float64) / norm[c], sigma=2)
        np.testing.assert_array_equal(out["blurred"][:, :, c].cpu().numpy(), ref)
    n = out["n"]
    assert 0 < n < H * W
    sums = out["X64"].sum(1)
    assert float((sums - 1).abs().max()) < 1e-13
    lin = out["row_index"].long() * W + out["column_index"].long()
    assert bool((lin[1:] > lin[:-1]).all())
    b = out["blurred"].reshape(-1, C)
    keep = (b.sum(1) > 20.0)
    assert abs(int(keep.sum()) - n) <= 2          # torch's sum order may flip a boundary pixel
    # hand-off: the fp32 rows go straight into the assignment kernel
    W0 = out["X32"][:100].contiguous()
    lab = S.bmu(out["X32"], W0)
    torch.cuda.synchronize()
    assert lab.shape == (n,) and int(lab.min()) >= 1 and int(lab.max()) <= 100


@pytest.mark.parametrize("q", [0.999, 0.5, 0.05, 0.0, 1.0, 0.3141])
def test_column_quantile_matches_pandas_bit_for_bit(q, rng):
    X = rng.random((5000, 9)) * np.array([1, 1e-3, 50, 1, 1, 1, 1, 1e-300, 1])
    X[rng.random(X.shape) < 0.3] = 0
    X[:, 3] = np.round(X[:, 3], 1)                 # heavy ties
    X[:, 4] = 0                                    # no valid entry -> NaN
    X[:, 5] = np.where(rng.random(5000) < 0.5, -X[:, 5], X[:, 5])   # negative values
    X[:7, 6] = np.nan                              # NaN entries are skipped like zeros
    X[1:, 8] = 0                                   # a single valid entry
    chans = ['c%d' % i for i in range(9)]
    ref = PO.fov_channel_quantiles(pd.DataFrame(X, columns=chans), chans, q)
    got = PP.fov_channel_quantiles(torch.from_numpy(X).cuda(), chans, q)
    np.testing.assert_array_equal(got.values, ref.values)
    assert list(got.index) == chans
    lo, hi, m = PP.column_order_stats(torch.from_numpy(X).cuda(), q)
    np.testing.assert_array_equal(m, ((X != 0) & ~np.isnan(X)).sum(0))


def test_quantiles_of_a_preprocessed_fov_and_edge_shapes(rng):
    g = np.load(os.path.join(GOLD, "preprocess_48x40x8.npz"))
    out = PP.preprocess_fov_device(g["img"], g["norm"], float(g["thresh"]), float(g["sigma"]))
    chans = ['chan%d' % i for i in range(8)]
    ref = PO.fov_channel_quantiles(pd.DataFrame(g["X64"], columns=chans), chans, 0.999)
    got = PP.fov_channel_quantiles(out["X64"], chans, 0.999, name=0.999)
    pd.testing.assert_series_equal(got, ref)
    # empty matrix, one row, one column
    assert np.isnan(PP.column_quantile(torch.empty((0, 3), dtype=torch.float64, device="cuda"), 0.5)).all()
    one = torch.tensor([[0.25, 0.0, 3.0]], dtype=torch.float64, device="cuda")
    np.testing.assert_array_equal(PP.column_quantile(one, 0.999), np.array([0.25, np.nan, 3.0]))
    # full-size column set: 2^20 rows x 32 channels against np.quantile per column
    X = torch.rand((1 << 20, 32), dtype=torch.float64, device="cuda")
    X[X < 0.2] = 0
    got = PP.column_quantile(X, 0.999)
    np.testing.assert_array_equal(got, PO.column_quantile_explicit(X.cpu().numpy(), 0.999))


@pytest.mark.parametrize("with_seg,sub_dir", [(True, 'TIFs'), (False, None)])
def test_preprocess_fov_writes_the_reference_files(with_seg, sub_dir, rng, tmp_path):
    """preprocess_fov end to end on the GPU: both Feather files and the return value equal the
    scipy + pandas route's (same rows sampled under the same seed)."""
    import preprocess_fixtures as PF
    out, chans = PF.run_both(str(tmp_path), rng, PP.preprocess_fov, PO.preprocess_fov,
                             with_seg=with_seg, sub_dir=sub_dir)
    (m_ret, m_full, m_sub), (o_ret, o_full, o_sub) = out['mirror'], out['oracle']
    pd.testing.assert_frame_equal(m_full, o_full)
    pd.testing.assert_frame_equal(m_sub, o_sub)
    pd.testing.assert_frame_equal(m_ret, o_ret)
    assert 0 < len(m_full) < 40 * 36


def test_create_pixel_matrix_matches_the_reference_route(tmp_path, rng, capsys):
    """The cohort driver end to end: raw-image statistics, per-FOV tables, the post-row-norm
    normalisation row -- every file equal to the numpy / scipy / pandas route's; then the restart
    and the channel-change behaviours of the reference (pixie_preprocessing.py:260-300)."""
    import preprocess_fixtures as PF
    fovs, chans = ['fov0', 'fov1'], ['chan0', 'chan1', 'chan2']
    # integer-typed images (the usual TIFFs): the raw-image quantiles, hence the normalisation row,
    # are float64 and the reference's arithmetic is float64 throughout -- the case mirrored bit for
    # bit.  (float32 images make numpy keep the whole reference pipeline in float32; the mirror
    # stays in fp64 and then agrees to ~1e-6 relative: DESIGN.md section 5.2.)
    tiff_dir, seg_dir, _ = PF.make_tree(str(tmp_path), rng, dtype=np.uint16)
    bases = {}
    for tag in ('mirror', 'oracle'):
        base = os.path.join(str(tmp_path), tag)
        os.makedirs(os.path.join(base, 'pixel_output_dir'))
        bases[tag] = base
    PP.create_pixel_matrix(list(fovs), list(chans), bases['mirror'], tiff_dir, seg_dir)
    assert "Processed 2 fovs" in capsys.readouterr().out
    pre, thresh, post = PO.create_pixel_matrix(list(fovs), list(chans), bases['oracle'], tiff_dir, seg_dir)
    m, o = bases['mirror'], bases['oracle']
    pd.testing.assert_frame_equal(
        paf.read_feather(os.path.join(m, 'pixel_output_dir', 'channel_norm_pre_rownorm.feather')), pre)
    got_thresh = paf.read_feather(os.path.join(m, 'pixel_output_dir', 'pixel_thresh.feather'))
    assert got_thresh['pixel_thresh_val'].values[0] == thresh
    pd.testing.assert_frame_equal(
        paf.read_feather(os.path.join(m, 'channel_norm_post_rownorm.feather')),
        paf.read_feather(os.path.join(o, 'channel_norm_post_rownorm.feather')))
    for fov in fovs:
        for d in ('pixel_mat_data', 'pixel_mat_subsetted'):
            pd.testing.assert_frame_equal(paf.read_feather(os.path.join(m, d, fov + '.feather')),
                                          paf.read_feather(os.path.join(o, d, fov + '.feather')))
    assert not os.path.exists(os.path.join(m, 'pixel_mat_data', 'channel_norm_post_rownorm_perfov.csv'))
    # everything is there: a second call does nothing
    PP.create_pixel_matrix(list(fovs), list(chans), m, tiff_dir, seg_dir)
    assert "There are no more FOVs to preprocess, skipping" in capsys.readouterr().out
    # a lost subset file: only that FOV is redone
    os.remove(os.path.join(m, 'pixel_mat_subsetted', 'fov1.feather'))
    PP.create_pixel_matrix(list(fovs), list(chans), m, tiff_dir, seg_dir)
    assert os.path.exists(os.path.join(m, 'pixel_mat_subsetted', 'fov1.feather'))
    # other channels: the cohort starts over
    PP.create_pixel_matrix(list(fovs), ['chan0', 'chan1'], m, tiff_dir, seg_dir)
    assert "New channels provided: overwriting whole cohort" in capsys.readouterr().out
    assert list(paf.read_feather(os.path.join(m, 'channel_norm_post_rownorm.feather')).columns) == \
        ['chan0', 'chan1']
    with pytest.raises(ValueError):
        PP.create_pixel_matrix(list(fovs), list(chans), m, tiff_dir, seg_dir, subset_proportion=0)
