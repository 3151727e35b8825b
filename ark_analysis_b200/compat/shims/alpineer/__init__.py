"""The `alpineer` helpers the Pixie SOM path calls."""
from . import image_utils, io_utils, load_utils, misc_utils, test_utils  # noqa: F401
