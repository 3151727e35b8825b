"""``generate_pixel_cluster_mask`` -- the consumer of the label array that paints cluster ids back
onto the image grid (reference src/ark/utils/data_utils.py:476-555).  SURVEY.md section 8f, N4."""
import os

import numpy as np

from . import io_utils


def _image_shape(path):
    """(rows, cols) of the sample channel image (the reference reads it with skimage only to size
    the mask, data_utils.py:519-523)."""
    from PIL import Image
    with Image.open(path) as im:
        w, h = im.size
    return h, w


def generate_pixel_cluster_mask(fov, base_dir, tiff_dir, chan_file_path, pixel_data_dir,
                                cluster_mapping, pixel_cluster_col='pixel_meta_cluster',
                                device=None):
    """int16 [H, W] image with every pixel of ``fov`` labelled by the ``cluster_id`` its SOM / meta
    cluster maps to; pixels absent from the pixel table stay 0.  Same arguments, checks and errors
    as the reference; the scatter runs on the GPU (``pixie_scatter_labels_i16``, duplicate
    coordinates resolved like numpy: the last row wins)."""
    import torch

    from . import som
    io_utils.validate_paths([tiff_dir, os.path.join(tiff_dir, chan_file_path),
                             os.path.join(base_dir, pixel_data_dir)])
    io_utils.verify_in_list(provided_cluster_col=[pixel_cluster_col],
                            valid_cluster_cols=['pixel_som_cluster', 'pixel_meta_cluster'])
    io_utils.verify_in_list(provided_fov_file=[fov + '.feather'],
                            consensus_fov_files=os.listdir(os.path.join(base_dir, pixel_data_dir)))
    H, W = _image_shape(os.path.join(tiff_dir, chan_file_path))
    table = io_utils.read_table(os.path.join(base_dir, pixel_data_dir, fov + '.feather'),
                                columns=[pixel_cluster_col, 'row_index', 'column_index'])
    dev = torch.device(device) if device is not None else som._default_device()

    def col(name):
        a = table.column(name).combine_chunks().to_numpy(zero_copy_only=False)
        return torch.from_numpy(np.ascontiguousarray(a)).to(dev).to(torch.int32)  # astype(int)

    clusters = col(pixel_cluster_col)
    pairs = cluster_mapping.drop_duplicates()[[pixel_cluster_col, 'cluster_id']]
    id_of = dict(zip(pairs[pixel_cluster_col], pairs['cluster_id']))
    present = torch.unique(clusters).cpu().numpy() if clusters.numel() else np.empty(0, np.int32)
    for k in present:
        if int(k) not in id_of:
            raise KeyError(int(k))  # the reference's dict lookup fails the same way
    top = max([int(k) for k in id_of] + [0])
    if present.size and int(present.min()) < 0:
        raise KeyError(int(present.min()))
    lut = np.zeros(top + 1, np.int16)
    for k, v in id_of.items():
        if int(k) >= 0:
            lut[int(k)] = v
    img, _ = som.scatter_labels(col('row_index'), col('column_index'), clusters, H, W, id_map=lut)
    return img.cpu().numpy()
