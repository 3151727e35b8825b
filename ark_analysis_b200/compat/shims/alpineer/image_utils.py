"""Only imported (pixel_cluster_utils.py:10); its one call site on this path (`save_image` in the
smoothing helper) is outside the SOM train / assign path."""


def save_image(fname, data, compression_level=6):
    raise NotImplementedError("alpineer.image_utils.save_image is outside the Pixie SOM path")
