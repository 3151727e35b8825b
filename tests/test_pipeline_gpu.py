"""The rows of SURVEY.md section 8 composed on the device: image stack -> preprocessing (N3) ->
normalisation quantiles (N3) -> SOM training and assignment (a1-a4) -> per-cell cluster counts and
cluster mask (N4), every hand-off device-resident, each stage checked against the oracle chain
(scipy + pandas preprocessing, C restatement of the SOM, pandas / numpy label consumers)."""
import numpy as np
import pandas as pd
import pytest
import torch

import oracle
from oracle import label_oracle as LO, preprocess_oracle as PO
from ark_analysis_b200 import pixie_preprocessing as PP, som as S

pytestmark = pytest.mark.gpu


def test_image_to_masks_device_resident(rng):
    H, W, C, xd, yd = 96, 80, 6, 4, 3
    K = xd * yd
    img = rng.gamma(0.7, 1.0, (H, W, C)).astype(np.float32)
    img[:, 60:, :] = 0                                  # an empty strip: pixels that get dropped
    norm = rng.uniform(0.5, 2.0, C)
    yy, xx = np.mgrid[0:H, 0:W]
    seg = ((yy // 12) * 7 + xx // 12 + 1).astype(np.int32)   # 12 x 12 "cells"
    chans = ['chan%d' % i for i in range(C)]
    thresh = 1.0

    # ---- oracle chain
    x = img / norm.reshape(1, 1, C)
    mat, _ = PO.create_fov_pixel_data('fov0', list(chans), x, seg, thresh)
    qref = PO.fov_channel_quantiles(mat, chans, 0.999).values
    Xo = (mat[chans].values / qref).astype(np.float32)         # normalize_data, then the fp32 matrix
    idx = oracle.init_codebook_indices(Xo.shape[0], K, 42)
    Wo = oracle.som_batch(Xo, xd, yd, rlen=2, init_idx=idx)
    lab_o, _ = oracle.map_data_to_nodes_f32(Wo.astype(np.float32), Xo)

    # ---- device chain
    out = PP.preprocess_fov_device(img, norm, thresh, 2, seg)
    assert out["n"] == len(mat)
    q = PP.fov_channel_quantiles(out["X64"], chans, 0.999).values
    np.testing.assert_array_equal(q, qref)
    Xd = S.to_device_matrix((out["X64"] / torch.from_numpy(q).cuda()).float())
    np.testing.assert_array_equal(Xd.cpu().numpy(), Xo)
    Wd = S.train_som(Xd, torch.from_numpy(Xo[idx].astype(np.float64)).cuda(), xd, yd, rlen=2)
    rel = float(np.abs(Wd.cpu().numpy() - Wo).max() / np.abs(Wo).max())
    assert rel < 1e-4                                       # north-star tolerance for the weights
    # assignment against the ORACLE's codebook (labels are bit-exact for a given codebook)
    lab = S.bmu(Xd, torch.from_numpy(Wo.astype(np.float32)).cuda())
    np.testing.assert_array_equal(lab.cpu().numpy(), lab_o)

    # ---- consumers of the labels, still on the device
    counts, bad = S.label_histogram(out["label"], lab, int(seg.max()) + 1, K + 1)
    ref_counts, _ = LO.label_histogram(mat['label'].values, lab_o, int(seg.max()) + 1, K + 1)
    np.testing.assert_array_equal(counts.cpu().numpy(), ref_counts)
    assert int(bad) == 0 and int(counts.sum()) == out["n"]
    table = LO.fov_cluster_counts(pd.DataFrame({'label': mat['label'].values,
                                                'pixel_som_cluster': lab_o}), 'pixel_som_cluster')
    rows = np.flatnonzero(ref_counts.sum(1) > 0)
    np.testing.assert_array_equal(table.values, ref_counts[np.ix_(rows, np.flatnonzero(ref_counts.sum(0) > 0))])
    mask, _ = S.scatter_labels(out["row_index"], out["column_index"], lab, H, W, unique=True)
    ref_mask = LO.pixel_cluster_mask(mat['row_index'].values, mat['column_index'].values, lab_o,
                                     {k: k for k in range(K + 1)}, H, W)
    np.testing.assert_array_equal(mask.cpu().numpy(), ref_mask)
    assert (mask.cpu().numpy()[:, 70:] == 0).all()          # dropped pixels stay unlabelled
