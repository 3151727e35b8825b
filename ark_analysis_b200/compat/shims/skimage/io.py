def imread(fname, *args, **kwargs):
    raise NotImplementedError("skimage.io.imread is outside the Pixie SOM path")
