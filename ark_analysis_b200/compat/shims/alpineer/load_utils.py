"""Only imported (pixel_cluster_utils.py:10); image loading is upstream of the SOM path."""


def load_imgs_from_tree(*args, **kwargs):
    raise NotImplementedError("alpineer.load_utils.load_imgs_from_tree is outside the Pixie SOM path")
