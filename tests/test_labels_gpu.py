"""GPU parity tests of the N4 row: the label-array consumers
(pixie_label_histogram_i32, pixie_scatter_labels_i16 through the C ABI) and the host mirrors of
create_c2pc_data / generate_pixel_cluster_mask against the oracle (oracle/label_oracle.py, itself
pinned on the reference's known answers).  Integer work: everything is compared bit-exactly."""
import os
import tempfile
import warnings

import numpy as np
import pandas as pd
import pyarrow.feather as paf
import pytest
import torch

from oracle import label_oracle as LO
import label_fixtures as LF
from ark_analysis_b200 import cell_cluster_utils, data_utils, som as S

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n", [0, 1, 3, 4, 5, 127, 4096, 100003, 1 << 20])
def test_histogram_kernel_bit_exact(n, rng):
    n_seg, n_clu = 301, 100
    # image-order-like runs of equal cells, a few out-of-range values either side
    seg = np.repeat(rng.integers(-1, n_seg + 1, n // 7 + 1), 7)[:n].astype(np.int32)
    clu = np.repeat(rng.integers(-1, n_clu + 1, n // 3 + 1), 3)[:n].astype(np.int32)
    ref, nbad = LO.label_histogram(seg, clu, n_seg, n_clu)
    counts, bad = S.label_histogram(torch.from_numpy(seg).cuda(), torch.from_numpy(clu).cuda(),
                                    n_seg, n_clu)
    np.testing.assert_array_equal(counts.cpu().numpy(), ref)
    assert int(bad) == nbad
    # accumulation over chunks gives the same table
    if n > 8:
        h = n // 2 // 4 * 4 + 1   # second chunk starts unaligned: the wrapper re-aligns it
        c2, _ = S.label_histogram(torch.from_numpy(seg[:h]).cuda(), torch.from_numpy(clu[:h]).cuda(),
                                  n_seg, n_clu)
        S.label_histogram(torch.from_numpy(seg).cuda()[h:], torch.from_numpy(clu).cuda()[h:],
                          n_seg, n_clu, counts=c2)
        np.testing.assert_array_equal(c2.cpu().numpy(), ref)


def test_histogram_full_fov_size_properties(rng):
    """2048 x 2048 pixels: the table sums to the number of in-range pixels and its marginals equal
    the bincounts of the two inputs."""
    n, n_seg, n_clu = 2048 * 2048, 5000, 400
    g = torch.Generator(device="cuda").manual_seed(7)
    seg = torch.randint(0, n_seg, (n // 64,), device="cuda", generator=g,
                        dtype=torch.int32).repeat_interleave(64)
    clu = torch.randint(0, n_clu, (n,), device="cuda", generator=g, dtype=torch.int32)
    counts, bad = S.label_histogram(seg, clu, n_seg, n_clu)
    assert int(bad) == 0 and int(counts.sum()) == n
    assert torch.equal(counts.sum(1), torch.bincount(seg, minlength=n_seg).to(torch.int64))
    assert torch.equal(counts.sum(0), torch.bincount(clu, minlength=n_clu).to(torch.int64))


@pytest.mark.parametrize("with_map", [True, False])
def test_scatter_kernel_matches_numpy_including_duplicates(with_map, rng):
    H, W, n = 97, 131, 30000                      # n > H * W: many duplicates
    r = rng.integers(0, H, n).astype(np.int32)
    c = rng.integers(0, W, n).astype(np.int32)
    k = rng.integers(1, 101, n).astype(np.int32)
    id_map = {i: (i * 7) % 50 + 1 for i in range(1, 101)} if with_map else {i: i for i in range(101)}
    ref = LO.pixel_cluster_mask(r, c, k, id_map, H, W)
    lut = None
    if with_map:
        lut = np.zeros(101, np.int16)
        for a, b in id_map.items():
            lut[a] = b
    img, bad = S.scatter_labels(torch.from_numpy(r).cuda(), torch.from_numpy(c).cuda(),
                                torch.from_numpy(k).cuda(), H, W, id_map=lut)
    np.testing.assert_array_equal(img.cpu().numpy(), ref)
    assert int(bad) == 0
    # unique coordinates: the one-pass form
    perm = rng.permutation(H * W)[:5000]
    r2, c2 = (perm // W).astype(np.int32), (perm % W).astype(np.int32)
    img2, _ = S.scatter_labels(torch.from_numpy(r2).cuda(), torch.from_numpy(c2).cuda(),
                               torch.from_numpy(k[:5000]).cuda(), H, W, id_map=lut, unique=True)
    np.testing.assert_array_equal(img2.cpu().numpy(),
                                  LO.pixel_cluster_mask(r2, c2, k[:5000], id_map, H, W))
    # out-of-range coordinates are skipped and counted
    r3 = r.copy()
    r3[:10] = H + 3
    _, bad3 = S.scatter_labels(torch.from_numpy(r3).cuda(), torch.from_numpy(c).cuda(),
                               torch.from_numpy(k).cuda(), H, W, id_map=lut)
    assert int(bad3) == 10


def test_create_c2pc_data_known_answers_and_oracle_equality(rng):
    with tempfile.TemporaryDirectory() as d:
        fovs, pix, cells = LF.c2pc_case(d, rng)
        with pytest.raises(ValueError):
            cell_cluster_utils.create_c2pc_data(fovs, 'consensus', 'cell_table',
                                                pixel_cluster_col='bad_col')
        bad = pd.read_csv(cells).rename({'cell_size': 'bad_col'}, axis=1)
        bad.to_csv(os.path.join(d, 'bad.csv'), index=False)
        with pytest.raises(ValueError):
            cell_cluster_utils.create_c2pc_data(fovs, pix, os.path.join(d, 'bad.csv'),
                                                pixel_cluster_col='pixel_som_cluster')
        for col, answer in (('pixel_som_cluster', LF.C2PC_SOM),
                            ('pixel_meta_cluster_rename', LF.C2PC_META)):
            counts, norm = cell_cluster_utils.create_c2pc_data(fovs, pix, cells, pixel_cluster_col=col)
            cols = ['%s_%d' % (col, i) for i in range(len(answer[0]))]
            np.testing.assert_array_equal(counts[cols].values, np.array(answer))
            np.testing.assert_array_equal(norm[cols].values, np.array(answer) / 5)
            o_counts, o_norm = LO.create_c2pc_data(fovs, pix, cells, pixel_cluster_col=col)
            pd.testing.assert_frame_equal(counts, o_counts)
            pd.testing.assert_frame_equal(norm, o_norm)


def test_create_c2pc_data_fov_sized_case_equals_the_pandas_route(rng):
    with tempfile.TemporaryDirectory() as d:
        fovs, pix, cells = LF.big_c2pc_case(d, rng)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            counts, norm = cell_cluster_utils.create_c2pc_data(fovs, pix, cells,
                                                               pixel_cluster_col='pixel_som_cluster')
            o_counts, o_norm = LO.create_c2pc_data(fovs, pix, cells,
                                                   pixel_cluster_col='pixel_som_cluster')
        assert len(counts) > 100
        pd.testing.assert_frame_equal(counts, o_counts)
        pd.testing.assert_frame_equal(norm, o_norm)


def test_generate_pixel_cluster_mask_mirror(rng):
    from PIL import Image
    fov = 'fov0'
    chans = ['chan0', 'chan1', 'chan2', 'chan3']
    with tempfile.TemporaryDirectory() as d:
        with pytest.raises(FileNotFoundError):
            data_utils.generate_pixel_cluster_mask(fov, d, 'bad_tiff_dir', 'bad_chan_file',
                                                   'bad_consensus_path', {})
        with pytest.raises(FileNotFoundError):
            data_utils.generate_pixel_cluster_mask(fov, d, d, 'bad_chan_file',
                                                   'bad_consensus_path', {})
        os.mkdir(os.path.join(d, 'fov0'))
        Image.fromarray(rng.integers(0, 5, (40, 50)).astype(np.int16)).save(
            os.path.join(d, 'fov0', 'chan0.tiff'))
        with pytest.raises(FileNotFoundError):
            data_utils.generate_pixel_cluster_mask(fov, d, d, os.path.join('fov0', 'chan0.tiff'),
                                                   'bad_consensus_path', {})
        os.mkdir(os.path.join(d, 'pixel_mat_consensus'))
        t = pd.DataFrame(rng.random((100, 4)), columns=chans)
        t['pixel_som_cluster'] = np.tile(np.arange(1, 11), 10)
        t['pixel_meta_cluster'] = np.tile(np.arange(2, 7), 20)
        t['row_index'] = rng.integers(0, 40, 100)
        t['column_index'] = rng.integers(0, 50, 100)
        paf.write_feather(t, os.path.join(d, 'pixel_mat_consensus', fov + '.feather'),
                          compression='uncompressed')
        mapping = pd.DataFrame.from_dict({
            "pixel_som_cluster": np.arange(1, 11),
            "pixel_meta_cluster": np.repeat(np.arange(2, 7), 2),
            "pixel_meta_cluster_rename": ["meta" + str(i) for i in np.repeat(np.arange(2, 7), 2)],
            "cluster_id": np.repeat(np.arange(1, 6), 2)})
        args = (fov, d, d, os.path.join('fov0', 'chan0.tiff'), 'pixel_mat_consensus', mapping)
        with pytest.raises(ValueError):
            data_utils.generate_pixel_cluster_mask(*args, 'bad_cluster')
        with pytest.raises(ValueError):
            data_utils.generate_pixel_cluster_mask('fov1', *args[1:], 'pixel_som_cluster')
        for col, top in (('pixel_som_cluster', 10), ('pixel_meta_cluster', 5)):
            mask = data_utils.generate_pixel_cluster_mask(*args, col)
            assert mask.shape == (40, 50) and mask.dtype == np.int16 and np.all(mask <= top)
            pairs = mapping.drop_duplicates()[[col, 'cluster_id']]
            id_map = dict(zip(pairs[col], pairs['cluster_id']))
            ref = LO.pixel_cluster_mask(t['row_index'].values, t['column_index'].values,
                                        t[col].values, id_map, 40, 50)
            np.testing.assert_array_equal(mask, ref)
        # a cluster the mapping does not know: KeyError, like the reference's dict lookup
        with pytest.raises(KeyError):
            data_utils.generate_pixel_cluster_mask(*args[:5], mapping[mapping.cluster_id < 5],
                                                   'pixel_som_cluster')
