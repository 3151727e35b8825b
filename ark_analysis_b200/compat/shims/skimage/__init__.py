"""`skimage.io.imread` is imported by pixel_cluster_utils.py:12 but not called on the SOM path."""
